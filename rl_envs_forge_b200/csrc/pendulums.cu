// envforge_b200 -- CartPole / PendulumDisk batched step, fused rollout and reset kernels (sm_100a, fp32).
//
// Replaces, for N independent instances, the reference's per-instance
//   CartPole.step      rl_envs_forge/envs/inverted_pendulum/cart_pole/cart_pole.py:107-180   (+ reset :91-105)
//   PendulumDisk.step  rl_envs_forge/envs/inverted_pendulum/pendulum_disk/pendulum_disk.py:108-172 (+ reset :94-106)
// State is one float4 (CartPole: x, x_dot, theta, theta_dot) or float2 (PendulumDisk: alpha, alpha_dot) per env, so a
// warp moves it with one fully coalesced 128-bit (64-bit) access per lane.  The reference integrates in float64; the
// device integrates in fp32 with accurate sincosf/division (no fast-math) and is held to 1e-5 relative per step.
#include "ef_common.cuh"

namespace ef {

struct PendArgs {
    ef_pendulum_params p;
    float* state;
    int32_t* ep_len;
    float* ep_ret;
    const float* action;
    uint64_t seed, t, env_offset;
    PhiloxKeys rk;  // round keys of `seed` (rollout policy draws)
    int64_t per_cta;  // step kernel: contiguous envs per CTA (a multiple of the CTA size), divided on the host
    float* obs;
    float* reward;
    uint8_t* terminated;
    uint8_t* truncated;
    float* final_obs;
    float* acc;
    double* stats;
    uint32_t* err;  // rejected env-steps: an action outside [-3, 3] or NaN (the reference asserts, cart_pole.py:108)
    int64_t n;
    int32_t autoreset;
    int32_t K;
    uint8_t* flags;
    const uint8_t* mask;
    unsigned long long* clock;
};

// (a + pi) floored-mod 2pi - pi with -pi -> +pi (cart_pole.py:84-89).  On (-pi, pi] the map is the identity, which is
// what the fast path returns; fp32 has no value strictly between pi and 3.14159274f, so |a| >= 3.14159274f is exactly
// "outside (-pi, pi]".  The rare wrap is evaluated in double like the reference.
__device__ __forceinline__ float normalize_angle(float a) {
    if (fabsf(a) < 3.14159274f) return a;
    constexpr double kPi = 3.141592653589793, kTwoPi = 6.283185307179586;
    double m = fmod(static_cast<double>(a) + kPi, kTwoPi);
    if (m < 0.0) m += kTwoPi;
    double r = m - kPi;
    if (r == -kPi) r = kPi;
    return static_cast<float>(r);
}

__device__ __forceinline__ float angle_reward(const ef_pendulum_params& p, float angle) {
    if (p.reward_mode == 0) return (p.reward_lo <= angle && angle <= p.reward_hi) ? 1.0f : 0.0f;
    const float a = fabsf(angle);
    if (p.reward_mode == 1) return fmaf(-a, p.reward_inv_span, 1.0f);
    return fmaf(-(1.0f - expf(-p.reward_decay * a)), p.reward_inv_norm, 1.0f);
}

struct CartPole {
    static constexpr int DIM = 4;
    using Vec = float4;
    static __device__ __forceinline__ Vec init(const ef_pendulum_params& p, uint64_t seed, uint32_t g, uint64_t ctr,
                                               uint32_t stream) {
        if (p.fixed_init) return fixed_state(p);
        return from_words(p, draw_block(seed, g, ctr, stream));
    }
    static __device__ __forceinline__ Vec fixed_state(const ef_pendulum_params& p) {
        return make_float4(p.init_state[0], p.init_state[1], p.init_state[2], p.init_state[3]);
    }
    // initial state from the four words of the env's reset block (_initialize_state: U(init_lo, init_hi) per component)
    static __device__ __forceinline__ Vec from_words(const ef_pendulum_params& p, uint4 w) {
        const float span = p.init_hi - p.init_lo;
        return make_float4(uniform_f32(w.x, p.init_lo, span), uniform_f32(w.y, p.init_lo, span),
                           uniform_f32(w.z, p.init_lo, span), uniform_f32(w.w, p.init_lo, span));
    }
    // cart_pole.py:110-143.  Returns the "done" predicate before the truncation override; writes accelerations.
    static __device__ __forceinline__ bool advance(const ef_pendulum_params& p, Vec& s, float force, float& acc0,
                                                   float& acc1, float& angle) {
        float sn, cs;
        sincosf(s.z, &sn, &cs);
        // explicit fmaf + -fmad=false: the step and rollout kernels round identically (bit-equal trajectories)
        const float temp = fmaf(p.c[1] * (s.w * s.w), sn, force) * p.c[2];
        const float theta_acc = fmaf(p.c[0], sn, -(cs * temp)) / (p.c[3] * fmaf(-(p.c[5] * cs), cs, p.c[4]));
        const float x_acc = fmaf(-(p.c[6] * theta_acc), cs, temp);
        s.x = fmaf(p.tau, s.y, s.x);
        s.y = fmaf(p.tau, x_acc, s.y);
        s.z = normalize_angle(fmaf(p.tau, s.w, s.z));
        s.w = fmaf(p.tau, theta_acc, s.w);
        acc0 = x_acc;
        acc1 = theta_acc;
        angle = s.z;
        bool done = (s.x < -p.c[7]) | (s.x > p.c[7]) | (s.z < -p.angle_threshold) | (s.z > p.angle_threshold);
        if (p.angle_termination >= 0.0f) done |= fabsf(s.z) > p.angle_termination;
        return done;
    }
    static __device__ __forceinline__ void store_acc(float* acc, int64_t i, float a0, float a1) {
        reinterpret_cast<float2*>(acc)[i] = make_float2(a0, a1);
    }
};

struct PendulumDisk {
    static constexpr int DIM = 2;
    using Vec = float2;
    static __device__ __forceinline__ Vec init(const ef_pendulum_params& p, uint64_t seed, uint32_t g, uint64_t ctr,
                                               uint32_t stream) {
        if (p.fixed_init) return fixed_state(p);
        return from_words(p, draw_block(seed, g, ctr, stream));
    }
    static __device__ __forceinline__ Vec fixed_state(const ef_pendulum_params& p) {
        return make_float2(p.init_state[0], p.init_state[1]);
    }
    static __device__ __forceinline__ Vec from_words(const ef_pendulum_params& p, uint4 w) {
        const float span = p.init_hi - p.init_lo;
        return make_float2(uniform_f32(w.x, p.init_lo, span), uniform_f32(w.y, p.init_lo, span));
    }
    // pendulum_disk.py:111-136
    static __device__ __forceinline__ bool advance(const ef_pendulum_params& p, Vec& s, float u, float& acc0,
                                                   float& acc1, float& angle) {
        const float acc = fmaf(p.c[0], sinf(s.x), fmaf(-p.c[1], s.y, p.c[2] * u));
        s.x = normalize_angle(fmaf(p.tau, s.y, s.x));
        s.y = fmaf(p.tau, acc, s.y);
        acc0 = acc;
        acc1 = 0.0f;
        angle = s.x;
        // |alpha| > pi cannot fire after normalisation in the reference either (pendulum_disk.py:128-133)
        bool done = (s.y < -p.c[3]) | (s.y > p.c[3]);
        if (p.angle_termination >= 0.0f) done |= fabsf(s.x) > p.angle_termination;
        return done;
    }
    static __device__ __forceinline__ void store_acc(float* acc, int64_t i, float a0, float) { acc[i] = a0; }
};

// One reference step of one env: returns flags bit0 terminated, bit1 truncated.
template <class Env>
__device__ __forceinline__ uint32_t env_step(const ef_pendulum_params& p, typename Env::Vec& s, int32_t& len, float& ret,
                                             float action, float& reward, float& acc0, float& acc1) {
    float angle;
    const bool done = Env::advance(p, s, action, acc0, acc1, angle);
    reward = angle_reward(p, angle);
    ret += reward;
    len += 1;
    const bool trunc = len >= p.episode_limit;
    return trunc ? 2u : (done ? 1u : 0u);  // truncation overrides done (cart_pole.py:163-164)
}

__device__ __forceinline__ float policy_action(uint32_t w) { return uniform_f32(w, -3.0f, 6.0f); }

// `assert self.action_space.contains(action)` (cart_pole.py:108, pendulum_disk.py:109): Box(-3, 3) rejects |a| > 3 and NaN.
// A rejected env-step leaves the env untouched and is counted in `err` (detail = the top 24 bits of the fp32 action).
__device__ __forceinline__ bool action_rejected(float a) { return !(fabsf(a) <= 3.0f); }

#ifndef EF_PEND_GRID_PER_SM
#define EF_PEND_GRID_PER_SM 64  // measured on B200 at 4M envs: 64 (one pass, dynamic CTA scheduling) beats a persistent 3/SM grid by 20-25 %
#endif
#ifndef EF_PEND_ROLLOUT_GRID_PER_SM
#define EF_PEND_ROLLOUT_GRID_PER_SM 8
#endif
#ifndef EF_PEND_U
#define EF_PEND_U 4  // envs per thread whose loads are issued before any of them is stepped (memory-level parallelism)
#endif

// Step kernel (HBM-bound).  The batch is cut into one contiguous chunk per CTA; a thread issues the loads of EF_PEND_U
// envs (stride = CTA size, so every access of a warp is contiguous) before stepping them, which keeps ~4x more bytes in
// flight per SM than one env per thread.
// FAST != 0: the common call shape is known at compile time -- actions given, obs / reward / terminated / truncated all
// requested, no final_obs, no acc, auto-reset on -- so the launch-uniform null checks vanish.  FAST == 1: obs is a buffer
// of its own; FAST == 2: obs aliases the state buffer (the observation IS the new state, cart_pole.py:166 /
// pendulum_disk.py:158), so it is not written a second time: 16 (8) fewer bytes per env-step.
template <class Env, int FAST>
__global__ void __launch_bounds__(kThreads) pendulum_step_kernel(const PendArgs a) {
    using Vec = typename Env::Vec;
    pdl_launch_dependents();
    pdl_wait();
    EpisodeAcc acc;
    ErrorAcc errs;
    unsigned long long ticket;
    const uint64_t t = a.t + clock_begin(a.clock, ticket);
    const int64_t per_cta = a.per_cta;  // keeps every CTA's chunk 256-env aligned
    const int64_t begin = static_cast<int64_t>(blockIdx.x) * per_cta;
    const int64_t end = min(begin + per_cta, a.n);
    for (int64_t chunk = begin; chunk < end; chunk += static_cast<int64_t>(EF_PEND_U) * kThreads) {
        Vec s[EF_PEND_U];
        int32_t len[EF_PEND_U];
        float ret[EF_PEND_U], act[EF_PEND_U];
        // 32-bit bookkeeping inside the chunk; 64-bit only for the addresses
        const uint32_t left = static_cast<uint32_t>(min(end - chunk, static_cast<int64_t>(EF_PEND_U) * kThreads));
#pragma unroll
        for (int j = 0; j < EF_PEND_U; ++j) {
            const uint32_t local = static_cast<uint32_t>(j) * kThreads + threadIdx.x;
            if (local < left) {
                const int64_t i = chunk + local;
                s[j] = reinterpret_cast<const Vec*>(a.state)[i];
                len[j] = a.ep_len[i];
                ret[j] = a.ep_ret[i];
                act[j] = (FAST || a.action) ? __ldg(a.action + i) : 0.0f;
            }
        }
#pragma unroll
        for (int j = 0; j < EF_PEND_U; ++j) {
            const uint32_t local = static_cast<uint32_t>(j) * kThreads + threadIdx.x;
            if (local >= left) continue;
            const int64_t i = chunk + local;
            const uint32_t g = static_cast<uint32_t>(a.env_offset + i);
            if (!FAST && !a.action) act[j] = policy_action(group_word1(a.seed, g, t, kStreamPolicy));
            float rew = 0.0f, a0 = 0.0f, a1 = 0.0f;
            uint32_t f = 0u;
            if ((FAST || a.action) && action_rejected(act[j])) {
                errs.record(g, __float_as_uint(act[j]) >> 8, EF_STEP_BAD_ACTION);   // env untouched, outputs zero
            } else {
                f = env_step<Env>(a.p, s[j], len[j], ret[j], act[j], rew, a0, a1);
                acc.steps += 1u;
            }
            if (!FAST && a.final_obs) reinterpret_cast<Vec*>(a.final_obs)[i] = s[j];
            if ((FAST || a.autoreset) && f) {
                acc.finish(len[j], ret[j]);
                s[j] = Env::init(a.p, a.seed, g, t, kStreamResetAuto);
                len[j] = 0;
                ret[j] = 0.0f;
            }
            reinterpret_cast<Vec*>(a.state)[i] = s[j];
            a.ep_len[i] = len[j];
            a.ep_ret[i] = ret[j];
            if (FAST == 1 || (FAST == 0 && a.obs && a.obs != a.state)) reinterpret_cast<Vec*>(a.obs)[i] = s[j];
            if (FAST || a.reward) a.reward[i] = rew;
            if (FAST || a.terminated) a.terminated[i] = f & 1u;
            if (FAST || a.truncated) a.truncated[i] = (f >> 1) & 1u;
            if (!FAST && a.acc) Env::store_acc(a.acc, i, a0, a1);
        }
    }
    cta_epilogue(acc, errs, a.stats, a.err, a.clock, ticket, 1ull);
}

// Fused K-step rollout.  A thread keeps FOUR consecutive envs in registers for all K steps: one Philox call per step
// yields their four policy draws, external actions / emitted rewards move as one 128-bit access per step, and the four
// independent integrations give the FP32 pipes instruction-level parallelism.  A warp ballot skips the reset path in
// warps where no lane finished an episode this step.
template <class Env, bool EXT_ACTIONS, bool EMIT>
__global__ void __launch_bounds__(kThreads) pendulum_rollout_kernel(const PendArgs a) {
    using Vec = typename Env::Vec;
    pdl_launch_dependents();
    pdl_wait();
    EpisodeAcc acc;
    ErrorAcc errs;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t quads = (a.n + 3) / 4;
    const int64_t quads_round = (quads + 31) & ~static_cast<int64_t>(31);  // warp-uniform trip count
    unsigned long long ticket;
    const uint64_t t0 = a.t + clock_begin(a.clock, ticket);
    const bool io_vec = ((a.n & 3) == 0) && aligned(a.action, 16) && aligned(a.reward, 16) && aligned(a.flags, 4);
    for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < quads_round; v += stride) {
        const int64_t base = v * 4;
        const uint32_t g0 = static_cast<uint32_t>(a.env_offset + base);
        Vec s[4];
        int32_t len[4];
        float ret[4];
        bool valid[4];
        uint32_t rejected = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            valid[j] = base + j < a.n;
            const int64_t i = valid[j] ? base + j : 0;
            s[j] = reinterpret_cast<const Vec*>(a.state)[i];
            len[j] = a.ep_len[i];
            ret[j] = a.ep_ret[i];
        }
        const bool full = valid[3] && io_vec;
        for (int k = 0; k < a.K; ++k) {
            const uint64_t t = t0 + static_cast<uint64_t>(k);
            const int64_t o = static_cast<int64_t>(k) * a.n + base;
            float act[4];
            if constexpr (EXT_ACTIONS) {
                if (full) {
                    const float4 a4 = __ldg(reinterpret_cast<const float4*>(a.action + o));
                    act[0] = a4.x; act[1] = a4.y; act[2] = a4.z; act[3] = a4.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) act[j] = valid[j] ? __ldg(a.action + o + j) : 0.0f;
                }
            } else {
                uint32_t w[4];
                group_words4_rk(a.rk, a.seed, g0, t, kStreamPolicy, w);
#pragma unroll
                for (int j = 0; j < 4; ++j) act[j] = policy_action(w[j]);
            }
            float rew[4];
            uint32_t f[4];
            bool any_fin = false;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float a0, a1;
                if (EXT_ACTIONS && action_rejected(act[j])) {   // the reference asserts: env untouched, outputs zero
                    if (valid[j]) {
                        errs.record(g0 + j, __float_as_uint(act[j]) >> 8, EF_STEP_BAD_ACTION);
                        ++rejected;
                    }
                    rew[j] = 0.0f;
                    f[j] = 0u;
                } else {
                    f[j] = env_step<Env>(a.p, s[j], len[j], ret[j], act[j], rew[j], a0, a1);
                }
                any_fin |= valid[j] && f[j];
            }
            if constexpr (EMIT) {
                if (full) {
                    if (a.reward) *reinterpret_cast<float4*>(a.reward + o) = make_float4(rew[0], rew[1], rew[2], rew[3]);
                    if (a.flags) *reinterpret_cast<uint32_t*>(a.flags + o) = f[0] | (f[1] << 8) | (f[2] << 16) | (f[3] << 24);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (valid[j]) {
                            if (a.reward) a.reward[o + j] = rew[j];
                            if (a.flags) a.flags[o + j] = static_cast<uint8_t>(f[j]);
                        }
                    }
                }
            }
            if (__ballot_sync(0xffffffffu, any_fin)) {  // warp vote: reset path only where some lane finished
                bool fin[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    fin[j] = valid[j] && f[j];
                    if (fin[j]) {
                        acc.finish(len[j], ret[j]);
                        len[j] = 0;
                        ret[j] = 0.0f;
                    }
                }
                if (a.p.fixed_init) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (fin[j]) s[j] = Env::fixed_state(a.p);
                } else {
                    // Reset draws, one pass per finished slot OF A LANE rather than one pass per slot of the warp.  Each
                    // finished env needs its own Philox block (stream RESET_AUTO of (env, t), ~100 instructions).  Written
                    // as `for j: if (fin[j]) s[j] = init(...)` the warp executes the call once for every slot j in which
                    // ANY lane finished -- with random actions a CartPole episode lasts ~20 steps, so that is 3-4 calls
                    // per warp-step for a handful of active lanes (37 % of the executed instructions).  Here every lane
                    // walks its OWN pending slots; a lane rarely has more than one, so the warp makes ~1.4 calls.
                    uint32_t pending = (fin[0] ? 1u : 0u) | (fin[1] ? 2u : 0u) | (fin[2] ? 4u : 0u) | (fin[3] ? 8u : 0u);
                    while (__any_sync(0xffffffffu, pending != 0u)) {
                        if (pending != 0u) {
                            const uint32_t j = static_cast<uint32_t>(__ffs(pending)) - 1u;
                            const Vec fresh = Env::from_words(a.p, draw_block_rk(a.rk, g0 + j, t, kStreamResetAuto));
                            if (j == 0u) s[0] = fresh;
                            else if (j == 1u) s[1] = fresh;
                            else if (j == 2u) s[2] = fresh;
                            else s[3] = fresh;
                            pending &= pending - 1u;
                        }
                    }
                }
            }
            if constexpr (EMIT) {
                if (a.obs) {  // trajectory of observations [K, n, DIM]: the state after auto-reset, like step()
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (valid[j]) reinterpret_cast<Vec*>(a.obs)[o + j] = s[j];
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (valid[j]) {
                acc.steps += static_cast<uint32_t>(a.K);
                reinterpret_cast<Vec*>(a.state)[base + j] = s[j];
                a.ep_len[base + j] = len[j];
                a.ep_ret[base + j] = ret[j];
            }
        }
        acc.steps -= rejected;
    }
    cta_epilogue(acc, errs, a.stats, a.err, a.clock, ticket, static_cast<unsigned long long>(a.K));
}

template <class Env>
__global__ void __launch_bounds__(kThreads) pendulum_reset_kernel(const PendArgs a) {
    using Vec = typename Env::Vec;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        if (a.mask && !a.mask[i]) continue;
        reinterpret_cast<Vec*>(a.state)[i] =
            Env::init(a.p, a.seed, static_cast<uint32_t>(a.env_offset + i), a.t, kStreamResetUser);
        a.ep_len[i] = 0;
        a.ep_ret[i] = 0.0f;
    }
}

template <class Env>
static int check_common(const ef_pendulum_params* p, const float* state, const int32_t* ep_len, const float* ep_ret,
                        uint64_t env_offset, int64_t n) {
    if (!p || !state || !ep_len || !ep_ret || n < 0) return EF_ERR_INVALID_ARGUMENT;
    if (!aligned(state, sizeof(typename Env::Vec))) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    return EF_OK;
}

template <class Env>
static int step_impl(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret, const float* action,
                     uint64_t seed, uint64_t t, uint64_t env_offset, float* obs, float* reward, uint8_t* terminated,
                     uint8_t* truncated, float* final_obs, float* acc, double* stats, uint32_t* err, uint64_t* clock, int64_t n,
                     int32_t autoreset, void* stream) {
    if (int s = check_common<Env>(p, state, ep_len, ep_ret, env_offset, n)) return s;
    if (!aligned(err, 8)) return EF_ERR_INVALID_ARGUMENT;
    if (!aligned(obs, sizeof(typename Env::Vec)) || !aligned(final_obs, sizeof(typename Env::Vec)) ||
        !aligned(acc, Env::DIM == 4 ? 8 : 4))
        return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    PendArgs a{};
    a.p = *p; a.state = state; a.ep_len = ep_len; a.ep_ret = ep_ret; a.action = action;
    a.seed = seed; a.t = t; a.env_offset = env_offset; a.obs = obs; a.reward = reward;
    a.terminated = terminated; a.truncated = truncated; a.final_obs = final_obs; a.acc = acc;
    a.stats = stats; a.err = err; a.n = n; a.autoreset = autoreset; a.clock = reinterpret_cast<unsigned long long*>(clock);
    const int grid = grid_for((n + EF_PEND_U - 1) / EF_PEND_U, EF_PEND_GRID_PER_SM);
    a.per_cta = ((n + grid - 1) / grid + kThreads - 1) / kThreads * kThreads;
    const bool fast = action && obs && reward && terminated && truncated && !final_obs && !acc && autoreset != 0;
    if (fast && obs != state) return launch_kernel(pendulum_step_kernel<Env, 1>, a, grid, 0, static_cast<cudaStream_t>(stream));
    if (fast) return launch_kernel(pendulum_step_kernel<Env, 2>, a, grid, 0, static_cast<cudaStream_t>(stream));
    return launch_kernel(pendulum_step_kernel<Env, 0>, a, grid, 0, static_cast<cudaStream_t>(stream));
}

template <class Env>
static int rollout_impl(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret,
                        const float* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K,
                        float* rewards, uint8_t* flags, double* stats, uint32_t* err, uint64_t* clock, int64_t n,
                        void* stream, float* obs_traj = nullptr) {
    if (int s = check_common<Env>(p, state, ep_len, ep_ret, env_offset, n)) return s;
    if (!aligned(err, 8)) return EF_ERR_INVALID_ARGUMENT;
    if (K < 0 || !aligned(obs_traj, sizeof(typename Env::Vec))) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0 || K == 0) return EF_OK;
    PendArgs a{};
    a.p = *p; a.state = state; a.ep_len = ep_len; a.ep_ret = ep_ret; a.action = actions;
    a.seed = seed; a.t = t0; a.env_offset = env_offset; a.K = K; a.reward = rewards; a.flags = flags; a.obs = obs_traj;
    a.rk = philox_round_keys(seed);
    a.stats = stats; a.err = err; a.n = n; a.autoreset = 1; a.clock = reinterpret_cast<unsigned long long*>(clock);
    const int grid = grid_for((n + 3) / 4, EF_PEND_ROLLOUT_GRID_PER_SM);
    const cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool ext = actions != nullptr, emit = rewards != nullptr || flags != nullptr || obs_traj != nullptr;
    if (ext && emit) return launch_kernel(pendulum_rollout_kernel<Env, true, true>, a, grid, 0, st);
    if (ext) return launch_kernel(pendulum_rollout_kernel<Env, true, false>, a, grid, 0, st);
    if (emit) return launch_kernel(pendulum_rollout_kernel<Env, false, true>, a, grid, 0, st);
    return launch_kernel(pendulum_rollout_kernel<Env, false, false>, a, grid, 0, st);
}

template <class Env>
static int reset_impl(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret, uint64_t seed,
                      uint64_t epoch, uint64_t env_offset, const uint8_t* mask, int64_t n, void* stream) {
    if (int s = check_common<Env>(p, state, ep_len, ep_ret, env_offset, n)) return s;
    if (n == 0) return EF_OK;
    PendArgs a{};
    a.p = *p; a.state = state; a.ep_len = ep_len; a.ep_ret = ep_ret; a.seed = seed; a.t = epoch;
    a.env_offset = env_offset; a.mask = mask; a.n = n;
    pendulum_reset_kernel<Env><<<grid_for(n, 8), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return cuda_status(cudaGetLastError());
}

}  // namespace ef

using namespace ef;

extern "C" int ef_cartpole_step(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret,
                                const float* action, uint64_t seed, uint64_t t, uint64_t env_offset, float* obs,
                                float* reward, uint8_t* terminated, uint8_t* truncated, float* final_obs, float* acc,
                                double* stats, uint32_t* err, uint64_t* clock, int64_t n, int32_t autoreset, void* stream) {
    return step_impl<CartPole>(p, state, ep_len, ep_ret, action, seed, t, env_offset, obs, reward, terminated, truncated,
                               final_obs, acc, stats, err, clock, n, autoreset, stream);
}
extern "C" int ef_cartpole_rollout(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret,
                                   const float* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K,
                                   float* rewards, uint8_t* flags, double* stats, uint32_t* err, uint64_t* clock, int64_t n,
                                   void* stream) {
    return rollout_impl<CartPole>(p, state, ep_len, ep_ret, actions, seed, t0, env_offset, K, rewards, flags, stats, err, clock, n, stream);
}
extern "C" int ef_cartpole_rollout_traj(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret,
                                        const float* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K,
                                        float* obs, float* rewards, uint8_t* flags, double* stats, uint32_t* err, uint64_t* clock,
                                        int64_t n, void* stream) {
    return rollout_impl<CartPole>(p, state, ep_len, ep_ret, actions, seed, t0, env_offset, K, rewards, flags, stats, err, clock, n,
                                  stream, obs);
}

extern "C" int ef_cartpole_reset(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret,
                                 uint64_t seed, uint64_t epoch, uint64_t env_offset, const uint8_t* mask, int64_t n,
                                 void* stream) {
    return reset_impl<CartPole>(p, state, ep_len, ep_ret, seed, epoch, env_offset, mask, n, stream);
}
extern "C" int ef_pendulum_disk_step(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret,
                                     const float* action, uint64_t seed, uint64_t t, uint64_t env_offset, float* obs,
                                     float* reward, uint8_t* terminated, uint8_t* truncated, float* final_obs,
                                     float* acc, double* stats, uint32_t* err, uint64_t* clock, int64_t n,
                                     int32_t autoreset, void* stream) {
    return step_impl<PendulumDisk>(p, state, ep_len, ep_ret, action, seed, t, env_offset, obs, reward, terminated,
                                   truncated, final_obs, acc, stats, err, clock, n, autoreset, stream);
}
extern "C" int ef_pendulum_disk_rollout(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret,
                                        const float* actions, uint64_t seed, uint64_t t0, uint64_t env_offset,
                                        int32_t K, float* rewards, uint8_t* flags, double* stats, uint32_t* err,
                                        uint64_t* clock, int64_t n, void* stream) {
    return rollout_impl<PendulumDisk>(p, state, ep_len, ep_ret, actions, seed, t0, env_offset, K, rewards, flags, stats, err, clock, n, stream);
}
extern "C" int ef_pendulum_disk_rollout_traj(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret,
                                             const float* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K,
                                             float* obs, float* rewards, uint8_t* flags, double* stats, uint32_t* err,
                                             uint64_t* clock, int64_t n, void* stream) {
    return rollout_impl<PendulumDisk>(p, state, ep_len, ep_ret, actions, seed, t0, env_offset, K, rewards, flags, stats, err, clock,
                                      n, stream, obs);
}

extern "C" int ef_pendulum_disk_reset(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret,
                                      uint64_t seed, uint64_t epoch, uint64_t env_offset, const uint8_t* mask,
                                      int64_t n, void* stream) {
    return reset_impl<PendulumDisk>(p, state, ep_len, ep_ret, seed, epoch, env_offset, mask, n, stream);
}
