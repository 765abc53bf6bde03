// envforge_b200 -- ACML envs: GamblersProblem and JacksCarRental batched step / reset kernels (sm_100a).
//
// Replaces, for N independent instances, the reference's per-instance
//   GamblersProblem.step / reset   rl_envs_forge/envs/acml/gambler_problem/gambler_problem.py:36-66 / :68-76
//   JacksCarRental.step / reset    rl_envs_forge/envs/acml/car_rental/car_rental.py:106-121 / :123-143
//     (_move_cars_and_calculate_cost :54-74, _simulate_rentals_and_returns :76-104)
//
// GamblersProblem is integer state + one Bernoulli draw: `np.random.rand() < win_probability` becomes the exact integer
// test r < ceil(p * 2^32) on the 32-bit draw r (u = r * 2^-32), like the GridWorld outcome thresholds.  HBM-bound
// (42 B per env-step: action 4, capital/ep_len/ep_ret 12 read + 12 written, obs 4, reward 4, flags 2).
//
// JacksCarRental draws four Poisson variates per step through scipy.stats.poisson.rvs, i.e. numpy's legacy
// random_poisson.  For lam < 10 that is the multiplication method on doubles (lam >= 10: PTRS, see poisson4_general); it is restated here in FP64 over doubles
// built from two Philox words exactly as numpy builds one from two MT19937 outputs, so the integer results are
// bit-exact against the oracle fed the same words (no accept/reject flips).  Issue-bound (about eight Philox calls and
// sixteen FP64 multiplies per env-step), not HBM-bound.
#include "ef_common.cuh"

namespace ef {

constexpr uint32_t kStreamAcml = 6;  // per (env, t), block j: Poisson doubles of JacksCarRental (twin: oracle/acml.py)

// ---- GamblersProblem ------------------------------------------------------------------------------------------------
struct GamblerArgs {
    int32_t goal, start, limit;
    uint64_t win_thr;  // ceil(p * 2^32) in [0, 2^32]: win iff r < win_thr
    int32_t* state;
    int32_t* ep_len;
    float* ep_ret;
    const int32_t* action;
    const uint32_t* draw;
    uint64_t seed, t, env_offset;
    int32_t* obs;
    float* reward;
    uint8_t* terminated;
    uint8_t* truncated;
    int32_t* final_obs;
    double* stats;
    uint32_t* err;
    unsigned long long* clock;
    int64_t n;
    int32_t autoreset;
    // rollout only
    int32_t K;
    uint8_t* flags;
    PhiloxKeys rk;
};

// One env-step.  Returns bit0 terminated, bit1 truncated, bit2 rejected (state untouched: the reference asserts first).
__device__ __forceinline__ uint32_t gambler_transition(const GamblerArgs& a, uint32_t g, int32_t& s, int32_t& len, float& ret,
                                                       int32_t bet, uint32_t r, float& rew, ErrorAcc& errs) {
    rew = 0.0f;
    if (static_cast<uint32_t>(bet) > static_cast<uint32_t>(a.goal)) {  // assert self.action_space.contains(action)
        errs.record(g, static_cast<uint32_t>(bet), EF_STEP_BAD_ACTION);
        return 4u;
    }
    uint32_t f = 0u;
    if (static_cast<uint64_t>(r) < a.win_thr) {  // gambler_problem.py:51-56
        s += bet;
        if (s >= a.goal) {
            s = a.goal;
            rew = 1.0f;
            f = 1u;
        }
    } else {  // :57-62
        s -= bet;
        if (s <= 0) {
            s = 0;
            f = 1u;
        }
    }
    ret += rew;
    len += 1;
    if (a.limit > 0 && len >= a.limit) f |= 2u;
    return f;
}

// Random policy (action == NULL): a legal bet, uniform in 0 .. min(s, goal - s)  (the action range build_mdp enumerates).
__device__ __forceinline__ int32_t gambler_policy(const GamblerArgs& a, int32_t s, uint32_t w) {
    const int32_t hi = max(min(s, a.goal - s), 0);
    return static_cast<int32_t>(__umulhi(w, static_cast<uint32_t>(hi) + 1u));
}

constexpr int kGamblerQ = 2;  // quads per thread whose loads are in flight together

template <bool VEC>
__global__ void __launch_bounds__(kThreads) gambler_step_kernel(const GamblerArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    EpisodeAcc acc;
    ErrorAcc errs;
    unsigned long long ticket;
    const uint64_t t = a.t + clock_begin(a.clock, ticket);
    const bool policy = a.action == nullptr;
    if constexpr (VEC) {
        const int64_t quads = a.n >> 2;
        const int64_t first = static_cast<int64_t>(blockIdx.x) * (kThreads * kGamblerQ) + threadIdx.x;
        int4 s4[kGamblerQ], l4[kGamblerQ], a4[kGamblerQ];
        float4 r4[kGamblerQ];
#pragma unroll
        for (int j = 0; j < kGamblerQ; ++j) {
            const int64_t v = first + static_cast<int64_t>(j) * kThreads;
            if (v < quads) {
                s4[j] = *reinterpret_cast<const int4*>(a.state + v * 4);
                l4[j] = *reinterpret_cast<const int4*>(a.ep_len + v * 4);
                r4[j] = *reinterpret_cast<const float4*>(a.ep_ret + v * 4);
                a4[j] = policy ? make_int4(0, 0, 0, 0) : __ldg(reinterpret_cast<const int4*>(a.action + v * 4));
            }
        }
#pragma unroll
        for (int j = 0; j < kGamblerQ; ++j) {
            const int64_t v = first + static_cast<int64_t>(j) * kThreads;
            if (v >= quads) continue;
            const int64_t base = v * 4;
            const uint32_t g0 = static_cast<uint32_t>(a.env_offset + base);
            int32_t s[4] = {s4[j].x, s4[j].y, s4[j].z, s4[j].w};
            int32_t len[4] = {l4[j].x, l4[j].y, l4[j].z, l4[j].w};
            float ret[4] = {r4[j].x, r4[j].y, r4[j].z, r4[j].w};
            int32_t bet[4] = {a4[j].x, a4[j].y, a4[j].z, a4[j].w};
            uint32_t w[4], term = 0u, trunc = 0u;
            float rew[4];
            int32_t fo[4];
            group_words4(a.seed, g0, t, kStreamTransition, w);
            if (policy) {
                uint32_t wp[4];
                group_words4(a.seed, g0, t, kStreamPolicy, wp);
#pragma unroll
                for (int k = 0; k < 4; ++k) bet[k] = gambler_policy(a, s[k], wp[k]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t f = gambler_transition(a, g0 + k, s[k], len[k], ret[k], bet[k], w[k], rew[k], errs);
                if (!(f & 4u)) acc.steps += 1u;
                fo[k] = s[k];
                if (a.autoreset && (f & 3u)) {
                    acc.finish(len[k], ret[k]);
                    s[k] = a.start;
                    len[k] = 0;
                    ret[k] = 0.0f;
                }
                term |= (f & 1u) << (8 * k);
                trunc |= ((f >> 1) & 1u) << (8 * k);
            }
            *reinterpret_cast<int4*>(a.state + base) = make_int4(s[0], s[1], s[2], s[3]);
            *reinterpret_cast<int4*>(a.ep_len + base) = make_int4(len[0], len[1], len[2], len[3]);
            *reinterpret_cast<float4*>(a.ep_ret + base) = make_float4(ret[0], ret[1], ret[2], ret[3]);
            if (a.obs) *reinterpret_cast<int4*>(a.obs + base) = make_int4(s[0], s[1], s[2], s[3]);
            if (a.final_obs) *reinterpret_cast<int4*>(a.final_obs + base) = make_int4(fo[0], fo[1], fo[2], fo[3]);
            if (a.reward) *reinterpret_cast<float4*>(a.reward + base) = make_float4(rew[0], rew[1], rew[2], rew[3]);
            if (a.terminated) *reinterpret_cast<uint32_t*>(a.terminated + base) = term;
            if (a.truncated) *reinterpret_cast<uint32_t*>(a.truncated + base) = trunc;
        }
    }
    // scalar path: the whole batch (VEC == false: unaligned buffers or injected draws) or the ragged tail of < 4 envs
    const int64_t begin = VEC ? (a.n & ~static_cast<int64_t>(3)) : 0;
    for (int64_t i = begin + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint32_t g = static_cast<uint32_t>(a.env_offset + i);
        int32_t s = a.state[i], len = a.ep_len[i];
        float ret = a.ep_ret[i], rew;
        const int32_t bet = policy ? gambler_policy(a, s, group_word1(a.seed, g, t, kStreamPolicy)) : __ldg(a.action + i);
        const uint32_t r = a.draw ? __ldg(a.draw + i) : group_word1(a.seed, g, t, kStreamTransition);
        const uint32_t f = gambler_transition(a, g, s, len, ret, bet, r, rew, errs);
        if (!(f & 4u)) acc.steps += 1u;
        if (a.final_obs) a.final_obs[i] = s;
        if (a.autoreset && (f & 3u)) {
            acc.finish(len, ret);
            s = a.start;
            len = 0;
            ret = 0.0f;
        }
        a.state[i] = s;
        a.ep_len[i] = len;
        a.ep_ret[i] = ret;
        if (a.obs) a.obs[i] = s;
        if (a.reward) a.reward[i] = rew;
        if (a.terminated) a.terminated[i] = static_cast<uint8_t>(f & 1u);
        if (a.truncated) a.truncated[i] = static_cast<uint8_t>((f >> 1) & 1u);
    }
    cta_epilogue(acc, errs, a.stats, a.err, a.clock, ticket, 1ull);
}

// Fused K-step rollout: a thread keeps four consecutive gamblers in registers for all K steps (one Philox call per step
// for their four outcome draws, one more for their four random bets), auto-reset always on.  State is read and written
// once per launch (24 B / K per env-step); `reward` [K, n] / `flags` [K, n] (bit0 terminated, bit1 truncated) are optional.
__global__ void __launch_bounds__(kThreads) gambler_rollout_kernel(const GamblerArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    EpisodeAcc acc;
    ErrorAcc errs;
    unsigned long long ticket;
    const uint64_t t0 = a.t + clock_begin(a.clock, ticket);
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t quads = (a.n + 3) / 4;
    for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < quads; v += stride) {
        const int64_t base = v * 4;
        const uint32_t g0 = static_cast<uint32_t>(a.env_offset + base);
        int32_t s[4], len[4];
        float ret[4];
        bool valid[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            valid[j] = base + j < a.n;
            const int64_t i = valid[j] ? base + j : 0;
            s[j] = a.state[i];
            len[j] = a.ep_len[i];
            ret[j] = a.ep_ret[i];
        }
        for (int k = 0; k < a.K; ++k) {
            const uint64_t t = t0 + static_cast<uint64_t>(k);
            const int64_t o = static_cast<int64_t>(k) * a.n + base;
            uint32_t w[4], wp[4];
            group_words4_rk(a.rk, a.seed, g0, t, kStreamTransition, w);
            if (!a.action) group_words4_rk(a.rk, a.seed, g0, t, kStreamPolicy, wp);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!valid[j]) continue;
                const int32_t bet = a.action ? __ldg(a.action + o + j) : gambler_policy(a, s[j], wp[j]);
                float rew;
                const uint32_t f = gambler_transition(a, g0 + j, s[j], len[j], ret[j], bet, w[j], rew, errs);
                if (!(f & 4u)) acc.steps += 1u;
                if (a.reward) a.reward[o + j] = rew;
                if (a.flags) a.flags[o + j] = static_cast<uint8_t>(f & 3u);
                if (f & 3u) {
                    acc.finish(len[j], ret[j]);
                    s[j] = a.start;
                    len[j] = 0;
                    ret[j] = 0.0f;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (valid[j]) {
                a.state[base + j] = s[j];
                a.ep_len[base + j] = len[j];
                a.ep_ret[base + j] = ret[j];
            }
        }
    }
    cta_epilogue(acc, errs, a.stats, a.err, a.clock, ticket, static_cast<unsigned long long>(a.K));
}

// ---- JacksCarRental -------------------------------------------------------------------------------------------------
struct CarArgs {
    int32_t max_cars, max_move, credit, move_cost, limit;
    double lam[4], enlam[4];  // requests A, requests B, returns A, returns B; enlam = exp(-lam) from the host's libm
    // rates >= 10 (numpy's PTRS branch): per-rate constants from the host's libm -- log(lam), a, b, log(invalpha), vr
    double loglam[4], ptrs_a[4], ptrs_b[4], ptrs_log_invalpha[4], ptrs_vr[4];
    int2* state;
    int32_t* ep_len;
    float* ep_ret;
    const int32_t* action;
    uint64_t seed, t, env_offset;
    int2* obs;
    float* reward;
    uint8_t* terminated;
    uint8_t* truncated;
    int2* final_obs;
    double* stats;
    unsigned long long* clock;
    int64_t n;
    int32_t autoreset, reset_mode;  // reset_mode: 0 random (Philox), 1 equal (max_cars // 2 each)
    PhiloxKeys rk;                  // round keys of `seed` (step / rollout kernels)
    int32_t K;                      // rollout only
    uint8_t* flags;                 // rollout only: [K, n] bit1 = truncated
};

// numpy legacy next_double over the (env, step) Philox substream: ((a >> 5) * 2^26 + (b >> 6)) / 2^53 from two words.
struct DoubleStream {
    const PhiloxKeys& rk;
    uint64_t t;
    uint32_t g, consumed;
    uint4 blk;
    __device__ __forceinline__ DoubleStream(const PhiloxKeys& rk_, uint32_t g_, uint64_t t_) : rk(rk_), t(t_), g(g_), consumed(0) {
        blk = draw_block_rk(rk, g, t, kStreamAcml, 0);
    }
    __device__ __forceinline__ double next() {  // words are consumed in pairs, so a double never straddles two blocks
        if ((consumed & 3u) == 0u && consumed != 0u) blk = draw_block_rk(rk, g, t, kStreamAcml, consumed >> 2);
        const bool second = (consumed & 2u) != 0u;
        const uint32_t wa = second ? blk.z : blk.x, wb = second ? blk.w : blk.y;
        consumed += 2u;
        return (static_cast<double>(wa >> 5) * 67108864.0 + static_cast<double>(wb >> 6)) * (1.0 / 9007199254740992.0);
    }
};

constexpr int kMaxPoissonTrips = 4096;  // termination guard; lam < 10 makes more than ~100 trips astronomically unlikely

// The four Poisson draws of one step -- numpy legacy random_poisson, lam < 10: multiplication method (X = number of
// uniforms whose running product stays above exp(-lam); lam == 0 returns 0 without drawing) -- evaluated as ONE loop over
// the double stream.  The doubles are consumed in exactly the reference's order (requests A until the product drops, then
// requests B, returns A, returns B); what the single loop buys is that every lane of a warp draws in lock step, so the
// Philox block behind two consecutive doubles is generated by all lanes together instead of once per divergent branch,
// and a warp runs max(sum of trips) iterations instead of sum(max of trips).
__device__ __forceinline__ int4 poisson4(DoubleStream& ds, const CarArgs& a) {
    int32_t x0 = 0, x1 = 0, x2 = 0, x3 = 0;
    const uint32_t skip = (a.lam[0] == 0.0 ? 1u : 0u) | (a.lam[1] == 0.0 ? 2u : 0u) | (a.lam[2] == 0.0 ? 4u : 0u) |
                          (a.lam[3] == 0.0 ? 8u : 0u);  // stages that return 0 without drawing
    int stage = 0;
    while (stage < 4 && ((skip >> stage) & 1u)) ++stage;
    double prod = 1.0;
    for (int it = 0; it < kMaxPoissonTrips && stage < 4; ++it) {
        prod *= ds.next();
        const double e = stage == 0 ? a.enlam[0] : (stage == 1 ? a.enlam[1] : (stage == 2 ? a.enlam[2] : a.enlam[3]));
        if (prod > e) {
            x0 += stage == 0;
            x1 += stage == 1;
            x2 += stage == 2;
            x3 += stage == 3;
        } else {
            prod = 1.0;
            ++stage;
            while (stage < 4 && ((skip >> stage) & 1u)) ++stage;
        }
    }
    return make_int4(x0, x1, x2, x3);
}

// numpy's random_loggam (distributions.c): log-gamma by an asymptotic series after shifting x up to >= 7.
__device__ double loggam(double x) {
    if (x == 1.0 || x == 2.0) return 0.0;
    const int n = x < 7.0 ? static_cast<int>(7.0 - x) : 0;
    double x0 = x + n;
    const double x2 = (1.0 / x0) * (1.0 / x0);
    const double c[10] = {8.333333333333333e-02, -2.777777777777778e-03, 7.936507936507937e-04, -5.952380952380952e-04,
                          8.417508417508418e-04, -1.917526917526918e-03, 6.410256410256410e-03, -2.955065359477124e-02,
                          1.796443723688307e-01, -1.39243221690590e+00};
    double gl0 = c[9];
#pragma unroll
    for (int k = 8; k >= 0; --k) {
        gl0 = __dmul_rn(gl0, x2);
        gl0 = __dadd_rn(gl0, c[k]);
    }
    double gl = gl0 / x0 + 0.5 * 1.8378770664093453e+00 + (x0 - 0.5) * log(x0) - x0;
    for (int k = 1; k <= n; ++k) {
        gl -= log(x0 - 1.0);
        x0 -= 1.0;
    }
    return gl;
}

// The general form of poisson4: any stage whose rate is >= 10 draws by numpy's random_poisson_ptrs (Hoermann's
// transformed rejection: two doubles per trial, a quick-accept region, a quick-reject region, else a log-space test
// against loggam(k + 1)); the other stages run the multiplication method as above, in the same single loop and the same
// order of consumption.  Only instantiated into the kernels launched for such rates (template PTRS), so the default
// configuration (rates 3, 4, 3, 2) keeps the shorter loop.
__device__ int4 poisson4_general(DoubleStream& ds, const CarArgs& a) {
    int32_t x[4] = {0, 0, 0, 0};
    int stage = 0;
    while (stage < 4 && a.lam[stage] == 0.0) ++stage;
    double prod = 1.0;
    for (int it = 0; it < kMaxPoissonTrips && stage < 4; ++it) {
        const double lam = a.lam[stage];
        bool done_stage;
        if (lam >= 10.0) {
            const double pa = a.ptrs_a[stage], pb = a.ptrs_b[stage];
            const double U = ds.next() - 0.5;
            const double V = ds.next();
            const double us = 0.5 - fabs(U);
            const double kf = floor((2.0 * pa / us + pb) * U + lam + 0.43);  // us == 0: -inf, rejected by kf < 0 below
            bool accept;
            if (us >= 0.07 && V <= a.ptrs_vr[stage]) {
                accept = true;
            } else if (kf < 0.0 || (us < 0.013 && V > us)) {
                accept = false;
            } else {
                accept = (log(V) + a.ptrs_log_invalpha[stage] - log(pa / (us * us) + pb)) <=
                         (-lam + kf * a.loglam[stage] - loggam(kf + 1.0));
            }
            if (accept) x[stage] = static_cast<int32_t>(kf);
            done_stage = accept;
        } else {
            prod *= ds.next();
            done_stage = !(prod > a.enlam[stage]);
            if (!done_stage) x[stage] += 1;
        }
        if (done_stage) {
            prod = 1.0;
            ++stage;
            while (stage < 4 && a.lam[stage] == 0.0) ++stage;
        }
    }
    return make_int4(x[0], x[1], x[2], x[3]);
}

__device__ __forceinline__ int2 car_initial_state(const CarArgs& a, uint32_t g, uint64_t counter, uint32_t stream) {
    if (a.reset_mode == 1) return make_int2(a.max_cars / 2, a.max_cars / 2);
    const uint4 w = draw_block(a.seed, g, counter, stream);
    return make_int2(static_cast<int32_t>(__umulhi(w.x, static_cast<uint32_t>(a.max_cars) + 1u)),
                     static_cast<int32_t>(__umulhi(w.y, static_cast<uint32_t>(a.max_cars) + 1u)));
}

// One JacksCarRental transition of env g at step t from raw action `raw`; returns the reward.
template <bool PTRS>
__device__ __forceinline__ float car_transition(const CarArgs& a, uint32_t g, uint64_t t, int2& s, int32_t raw) {
    // _move_cars_and_calculate_cost (car_rental.py:54-74); the reference never validates the action
    const int32_t act = raw - a.max_move;
    const int32_t mag = abs(act), sign = (act > 0) - (act < 0);
    const int32_t moved = min(min(act > 0 ? s.x : s.y, mag), a.max_move);
    int32_t ca = max(s.x - moved * sign, 0), cb = max(s.y + moved * sign, 0);
    const int32_t cost = mag * a.move_cost;
    // _simulate_rentals_and_returns (:76-104): requests A, requests B, returns A, returns B -- in that order
    DoubleStream ds(a.rk, g, t);
    const int4 draws = PTRS ? poisson4_general(ds, a) : poisson4(ds, a);
    const int32_t rent_a = min(ca, draws.x), rent_b = min(cb, draws.y);
    const float rew = static_cast<float>((rent_a + rent_b) * a.credit - cost);
    s = make_int2(min(ca - rent_a + draws.z, a.max_cars), min(cb - rent_b + draws.w, a.max_cars));
    return rew;
}

// Fused K-step rollout: one env per thread, state in registers for all K steps; `action` [K, n] or NULL (uniform random
// action), `reward` [K, n] / `flags` [K, n] optional.  Episodes only end through the optional step limit.
template <bool PTRS>
__global__ void __launch_bounds__(kThreads) car_rental_rollout_kernel(const CarArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    EpisodeAcc acc;
    unsigned long long ticket;
    const uint64_t t0 = a.t + clock_begin(a.clock, ticket);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint32_t g = static_cast<uint32_t>(a.env_offset + i);
        int2 s = a.state[i];
        int32_t len = a.ep_len[i];
        float ret = a.ep_ret[i];
        for (int k = 0; k < a.K; ++k) {
            const uint64_t t = t0 + static_cast<uint64_t>(k);
            const int64_t o = static_cast<int64_t>(k) * a.n + i;
            const int32_t raw = a.action ? __ldg(a.action + o)
                                         : static_cast<int32_t>(__umulhi(group_word1(a.seed, g, t, kStreamPolicy),
                                                                         2u * static_cast<uint32_t>(a.max_move) + 1u));
            const float rew = car_transition<PTRS>(a, g, t, s, raw);
            ret += rew;
            len += 1;
            const bool trunc = a.limit > 0 && len >= a.limit;
            if (a.reward) a.reward[o] = rew;
            if (a.flags) a.flags[o] = trunc ? 2 : 0;
            if (trunc) {
                acc.finish(len, ret);
                s = car_initial_state(a, g, t, kStreamResetAuto);
                len = 0;
                ret = 0.0f;
            }
        }
        acc.steps += static_cast<uint32_t>(a.K);
        a.state[i] = s;
        a.ep_len[i] = len;
        a.ep_ret[i] = ret;
    }
    cta_epilogue(acc, ErrorAcc(), a.stats, nullptr, a.clock, ticket, static_cast<unsigned long long>(a.K));
}

template <bool PTRS>
__global__ void __launch_bounds__(kThreads) car_rental_step_kernel(const CarArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    EpisodeAcc acc;
    ErrorAcc errs;
    unsigned long long ticket;
    const uint64_t t = a.t + clock_begin(a.clock, ticket);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint32_t g = static_cast<uint32_t>(a.env_offset + i);
        int2 s = a.state[i];
        int32_t len = a.ep_len[i];
        float ret = a.ep_ret[i];
        const int32_t raw = a.action ? __ldg(a.action + i)
                                     : static_cast<int32_t>(__umulhi(group_word1(a.seed, g, t, kStreamPolicy),
                                                                     2u * static_cast<uint32_t>(a.max_move) + 1u));
        const float rew = car_transition<PTRS>(a, g, t, s, raw);
        ret += rew;
        len += 1;
        acc.steps += 1u;
        const bool trunc = a.limit > 0 && len >= a.limit;  // the reference never ends an episode; limit is an extension
        if (a.final_obs) a.final_obs[i] = s;
        if (a.autoreset && trunc) {
            acc.finish(len, ret);
            s = car_initial_state(a, g, t, kStreamResetAuto);
            len = 0;
            ret = 0.0f;
        }
        a.state[i] = s;
        a.ep_len[i] = len;
        a.ep_ret[i] = ret;
        if (a.obs) a.obs[i] = s;
        if (a.reward) a.reward[i] = rew;
        if (a.terminated) a.terminated[i] = 0;
        if (a.truncated) a.truncated[i] = static_cast<uint8_t>(trunc);
    }
    cta_epilogue(acc, errs, a.stats, nullptr, a.clock, ticket, 1ull);
}

__global__ void __launch_bounds__(kThreads) car_rental_reset_kernel(const CarArgs a, uint64_t epoch, const uint8_t* mask) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        if (mask && !mask[i]) continue;
        const int2 s = car_initial_state(a, static_cast<uint32_t>(a.env_offset + i), epoch, kStreamResetUser);
        a.state[i] = s;
        a.ep_len[i] = 0;
        a.ep_ret[i] = 0.0f;
        if (a.obs) a.obs[i] = s;
    }
}

__global__ void __launch_bounds__(kThreads) gambler_reset_kernel(int32_t start, int32_t* state, int32_t* ep_len, float* ep_ret,
                                                                 int32_t* obs, const uint8_t* mask, int64_t n) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        if (mask && !mask[i]) continue;
        state[i] = start;
        ep_len[i] = 0;
        ep_ret[i] = 0.0f;
        if (obs) obs[i] = start;
    }
}

static bool uses_ptrs(const CarArgs& a) {
    return a.lam[0] >= 10.0 || a.lam[1] >= 10.0 || a.lam[2] >= 10.0 || a.lam[3] >= 10.0;
}

static int fill_car_args(const ef_car_rental_params* p, CarArgs& a) {
    if (!p || p->max_cars < 0 || p->max_cars > (1 << 20) || p->max_move_cars < 0 || p->max_move_cars > (1 << 20))
        return EF_ERR_INVALID_ARGUMENT;
    for (int j = 0; j < 4; ++j) {
        if (!(p->lam[j] >= 0.0)) return EF_ERR_INVALID_ARGUMENT;
        if (p->lam[j] > 1.0e6) return EF_ERR_UNSUPPORTED;  // the draw must fit an int32 with room to spare
        a.lam[j] = p->lam[j];
        a.enlam[j] = p->exp_neg_lam[j];
        a.loglam[j] = p->log_lam[j];
        a.ptrs_a[j] = p->ptrs_a[j];
        a.ptrs_b[j] = p->ptrs_b[j];
        a.ptrs_log_invalpha[j] = p->ptrs_log_invalpha[j];
        a.ptrs_vr[j] = p->ptrs_vr[j];
        if (p->lam[j] >= 10.0 && !(p->ptrs_b[j] > 3.4 && p->ptrs_a[j] > 0.0)) return EF_ERR_INVALID_ARGUMENT;  // constants missing
    }
    a.max_cars = p->max_cars;
    a.max_move = p->max_move_cars;
    a.credit = p->rental_credit;
    a.move_cost = p->move_car_cost;
    a.limit = p->episode_limit;
    a.reset_mode = p->init_equal ? 1 : 0;
    return EF_OK;
}

}  // namespace ef

using namespace ef;

extern "C" int ef_gambler_step(const ef_gambler_params* p, int32_t* state, int32_t* ep_len, float* ep_ret,
                               const int32_t* action, const uint32_t* draw, uint64_t seed, uint64_t t, uint64_t env_offset,
                               int32_t* obs, float* reward, uint8_t* terminated, uint8_t* truncated, int32_t* final_obs,
                               double* stats, uint32_t* err, uint64_t* clock, int64_t n, int32_t autoreset, void* stream) {
    if (!p || p->goal_amount < 0 || p->goal_amount > (1 << 30) || p->win_threshold > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n < 0 || !state || !ep_len || !ep_ret) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    GamblerArgs a{};
    a.goal = p->goal_amount; a.start = p->start_capital; a.limit = p->episode_limit; a.win_thr = p->win_threshold;
    a.state = state; a.ep_len = ep_len; a.ep_ret = ep_ret; a.action = action; a.draw = draw;
    a.seed = seed; a.t = t; a.env_offset = env_offset;
    a.obs = obs; a.reward = reward; a.terminated = terminated; a.truncated = truncated; a.final_obs = final_obs;
    a.stats = stats; a.err = err; a.clock = reinterpret_cast<unsigned long long*>(clock); a.n = n; a.autoreset = autoreset;
    const bool vec = draw == nullptr && n >= 4 && aligned(state, 16) && aligned(ep_len, 16) && aligned(ep_ret, 16) &&
                     aligned(action, 16) && aligned(obs, 16) && aligned(reward, 16) && aligned(terminated, 4) &&
                     aligned(truncated, 4) && aligned(final_obs, 16);
    const cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (vec) {
        const int64_t quads = n >> 2;
        const int64_t ctas = (quads + kThreads * kGamblerQ - 1) / (kThreads * kGamblerQ);
        return launch_kernel(gambler_step_kernel<true>, a, static_cast<int>(ctas), 0, st);
    }
    return launch_kernel(gambler_step_kernel<false>, a, grid_for(n, 8), 0, st);
}

extern "C" int ef_gambler_rollout(const ef_gambler_params* p, int32_t* state, int32_t* ep_len, float* ep_ret,
                                  const int32_t* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K,
                                  float* rewards, uint8_t* flags, double* stats, uint32_t* err, uint64_t* clock, int64_t n,
                                  void* stream) {
    if (!p || p->goal_amount < 0 || p->goal_amount > (1 << 30) || p->win_threshold > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n < 0 || K < 0 || !state || !ep_len || !ep_ret) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0 || K == 0) return EF_OK;
    GamblerArgs a{};
    a.goal = p->goal_amount; a.start = p->start_capital; a.limit = p->episode_limit; a.win_thr = p->win_threshold;
    a.state = state; a.ep_len = ep_len; a.ep_ret = ep_ret; a.action = actions;
    a.seed = seed; a.t = t0; a.env_offset = env_offset; a.K = K; a.reward = rewards; a.flags = flags;
    a.stats = stats; a.err = err; a.clock = reinterpret_cast<unsigned long long*>(clock); a.n = n; a.autoreset = 1;
    a.rk = philox_round_keys(seed);
    return launch_kernel(gambler_rollout_kernel, a, grid_for((n + 3) / 4, 8), 0, static_cast<cudaStream_t>(stream));
}

extern "C" int ef_gambler_reset(int32_t start_capital, int32_t* state, int32_t* ep_len, float* ep_ret, int32_t* obs,
                                const uint8_t* mask, int64_t n, void* stream) {
    if (n < 0 || !state || !ep_len || !ep_ret) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    gambler_reset_kernel<<<grid_for(n, 8), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(start_capital, state, ep_len,
                                                                                             ep_ret, obs, mask, n);
    return cuda_status(cudaGetLastError());
}

extern "C" int ef_car_rental_step(const ef_car_rental_params* p, int32_t* state, int32_t* ep_len, float* ep_ret,
                                  const int32_t* action, uint64_t seed, uint64_t t, uint64_t env_offset, int32_t* obs,
                                  float* reward, uint8_t* terminated, uint8_t* truncated, int32_t* final_obs, double* stats,
                                  uint64_t* clock, int64_t n, int32_t autoreset, void* stream) {
    CarArgs a{};
    if (int s = fill_car_args(p, a)) return s;
    if (n < 0 || !state || !ep_len || !ep_ret || !aligned(state, 8) || !aligned(obs, 8) || !aligned(final_obs, 8))
        return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    a.state = reinterpret_cast<int2*>(state); a.ep_len = ep_len; a.ep_ret = ep_ret; a.action = action;
    a.seed = seed; a.t = t; a.env_offset = env_offset;
    a.obs = reinterpret_cast<int2*>(obs); a.reward = reward; a.terminated = terminated; a.truncated = truncated;
    a.final_obs = reinterpret_cast<int2*>(final_obs);
    a.stats = stats; a.clock = reinterpret_cast<unsigned long long*>(clock); a.n = n; a.autoreset = autoreset;
    a.rk = philox_round_keys(seed);
    if (uses_ptrs(a)) return launch_kernel(car_rental_step_kernel<true>, a, grid_for(n, 8), 0, static_cast<cudaStream_t>(stream));
    return launch_kernel(car_rental_step_kernel<false>, a, grid_for(n, 8), 0, static_cast<cudaStream_t>(stream));
}

extern "C" int ef_car_rental_rollout(const ef_car_rental_params* p, int32_t* state, int32_t* ep_len, float* ep_ret,
                                     const int32_t* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K,
                                     float* rewards, uint8_t* flags, double* stats, uint64_t* clock, int64_t n,
                                     void* stream) {
    CarArgs a{};
    if (int s = fill_car_args(p, a)) return s;
    if (n < 0 || K < 0 || !state || !ep_len || !ep_ret || !aligned(state, 8)) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0 || K == 0) return EF_OK;
    a.state = reinterpret_cast<int2*>(state); a.ep_len = ep_len; a.ep_ret = ep_ret; a.action = actions;
    a.seed = seed; a.t = t0; a.env_offset = env_offset; a.K = K; a.reward = rewards; a.flags = flags;
    a.stats = stats; a.clock = reinterpret_cast<unsigned long long*>(clock); a.n = n; a.autoreset = 1;
    a.rk = philox_round_keys(seed);
    if (uses_ptrs(a)) return launch_kernel(car_rental_rollout_kernel<true>, a, grid_for(n, 8), 0, static_cast<cudaStream_t>(stream));
    return launch_kernel(car_rental_rollout_kernel<false>, a, grid_for(n, 8), 0, static_cast<cudaStream_t>(stream));
}

extern "C" int ef_car_rental_reset(const ef_car_rental_params* p, int32_t* state, int32_t* ep_len, float* ep_ret,
                                   uint64_t seed, uint64_t epoch, uint64_t env_offset, int32_t* obs, const uint8_t* mask,
                                   int64_t n, void* stream) {
    CarArgs a{};
    if (int s = fill_car_args(p, a)) return s;
    if (n < 0 || !state || !ep_len || !ep_ret || !aligned(state, 8) || !aligned(obs, 8)) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    a.state = reinterpret_cast<int2*>(state); a.ep_len = ep_len; a.ep_ret = ep_ret;
    a.seed = seed; a.env_offset = env_offset; a.obs = reinterpret_cast<int2*>(obs); a.n = n;
    car_rental_reset_kernel<<<grid_for(n, 8), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, epoch, mask);
    return cuda_status(cudaGetLastError());
}
