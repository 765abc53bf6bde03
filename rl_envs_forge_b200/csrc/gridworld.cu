// envforge_b200 -- GridWorld batched step / fused rollout / reset kernels (sm_100a).
//
// Replaces, for N independent instances, the reference's per-instance
//   GridWorld.step  rl_envs_forge/envs/grid_world/grid_world.py:495-528
//   GridWorld.reset rl_envs_forge/envs/grid_world/grid_world.py:530-551
// The dict MDP the reference builds at construction (build_mdp :245-274) arrives here as a flat outcome table
// (see include/envforge_b200.h) that each CTA stages into shared memory once; a step is then
//   r (uint32 draw) -> outcome index by integer thresholds -> one 8-byte shared-memory load -> state update.
//
// Memory behaviour (the step kernel is HBM-bound): per env 16 B read (action, pos, ep_len, ep_ret) and 26 B
// written (pos, ep_len, ep_ret, obs[2], reward, terminated, truncated) = 42 B; all traffic is 128-bit vector
// loads/stores, four envs per thread, grid-stride over a grid sized to the SM count.
#include "ef_common.cuh"

namespace ef {

struct GridArgs {
    const uint2* table;  // ef_grid_outcome as raw 8-byte words: x = next | flags << 16, y = reward bits
    int32_t table_len;   // entries
    int32_t cols;
    uint32_t thr0, thr1, thr2;
    int32_t idx_cap;
    int32_t start_cell;
    int32_t limit;       // <= 0: none
    int32_t* pos;
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
    int64_t n;
    int32_t autoreset;
    // rollout only
    int32_t K;
    uint8_t* flags;
    unsigned long long* clock;
};

template <bool SMEM>
__device__ __forceinline__ const uint2* stage_table(const GridArgs& a) {
    extern __shared__ __align__(16) uint2 s_table[];
    if constexpr (SMEM) {
        for (int i = threadIdx.x; i < a.table_len; i += blockDim.x) s_table[i] = __ldg(a.table + i);
        __syncthreads();
        return s_table;
    } else {
        return a.table;
    }
}

template <int SLOTS>
__device__ __forceinline__ uint32_t outcome_index(const GridArgs& a, uint32_t r) {
    if constexpr (SLOTS == 1) {
        return 0u;
    } else {
        const uint32_t idx = (r >= a.thr0) + (r >= a.thr1) + (r >= a.thr2);
        return min(idx, static_cast<uint32_t>(a.idx_cap));
    }
}

// One transition of one env.  Returns flags: bit0 terminated, bit1 truncated, bit2 error (state untouched).
template <bool SMEM, int SLOTS>
__device__ __forceinline__ uint32_t transition(const GridArgs& a, const uint2* __restrict__ tab, uint32_t g,
                                               int32_t& pos, int32_t& len, float& ret, int32_t act, uint32_t r,
                                               float& reward_out, ErrorAcc& errs) {
    if (static_cast<uint32_t>(act) > 3u) {
        errs.record(g, static_cast<uint32_t>(act), EF_STEP_BAD_ACTION);
        reward_out = 0.0f;
        return 4u;
    }
    const uint32_t row = static_cast<uint32_t>(pos) * 4u + static_cast<uint32_t>(act);
    const uint32_t slot = row * SLOTS + outcome_index<SLOTS>(a, r);
    const uint2 o = SMEM ? tab[slot] : __ldg(tab + slot);
    const uint32_t oflags = o.x >> 16;
    if (oflags & (EF_GRID_NO_OUTCOMES | EF_GRID_BAD_PROBS)) {
        errs.record(g, row, (oflags & EF_GRID_NO_OUTCOMES) ? EF_STEP_NO_OUTCOMES : EF_STEP_BAD_PROBS);
        reward_out = 0.0f;
        return 4u;
    }
    pos = static_cast<int32_t>(o.x & 0xffffu);
    reward_out = __uint_as_float(o.y);
    ret += reward_out;
    len += 1;
    const uint32_t trunc = (a.limit > 0 && len >= a.limit) ? 2u : 0u;
    return (oflags & EF_GRID_DONE) | trunc;
}

// ---- step -------------------------------------------------------------------------------------------------------------
template <bool SMEM, int SLOTS, int VEC>
__global__ void __launch_bounds__(kThreads) gridworld_step_kernel(const GridArgs a) {
    const uint2* __restrict__ tab = stage_table<SMEM>(a);
    EpisodeAcc acc;
    ErrorAcc errs;
    const int64_t groups = (a.n + VEC - 1) / VEC;
    const bool need_philox = (a.action == nullptr) || (SLOTS == 4 && a.draw == nullptr);
    const uint64_t t = a.t + clock_read(a.clock);
    for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < groups;
         v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t base = v * VEC;
        int32_t pos[VEC], len[VEC], act[VEC];
        float ret[VEC];
        uint32_t r[VEC];
        const bool full = (VEC == 4) && (base + VEC <= a.n);
        if (full) {  // 128-bit loads
            const int4 p4 = *reinterpret_cast<const int4*>(a.pos + base);
            const int4 l4 = *reinterpret_cast<const int4*>(a.ep_len + base);
            const float4 r4 = *reinterpret_cast<const float4*>(a.ep_ret + base);
            pos[0] = p4.x; pos[1] = p4.y; pos[2] = p4.z; pos[3] = p4.w;
            len[0] = l4.x; len[1] = l4.y; len[2] = l4.z; len[3] = l4.w;
            ret[0] = r4.x; ret[1] = r4.y; ret[2] = r4.z; ret[3] = r4.w;
            if (a.action) {
                const int4 a4 = __ldg(reinterpret_cast<const int4*>(a.action + base));
                act[0] = a4.x; act[1] = a4.y; act[2] = a4.z; act[3] = a4.w;
            }
            if (SLOTS == 4 && a.draw) {
                const uint4 d4 = __ldg(reinterpret_cast<const uint4*>(a.draw + base));
                r[0] = d4.x; r[1] = d4.y; r[2] = d4.z; r[3] = d4.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const bool ok = base + j < a.n;
                pos[j] = ok ? a.pos[base + j] : 0;
                len[j] = ok ? a.ep_len[base + j] : 0;
                ret[j] = ok ? a.ep_ret[base + j] : 0.0f;
                act[j] = (ok && a.action) ? __ldg(a.action + base + j) : 0;
                r[j] = (ok && SLOTS == 4 && a.draw) ? __ldg(a.draw + base + j) : 0u;
            }
        }
        int32_t obs[2 * VEC], fobs[2 * VEC];
        float rew[VEC];
        uint8_t term[VEC], trunc[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const uint32_t g = static_cast<uint32_t>(a.env_offset + base + j);
            if (need_philox) {
                const uint4 w = draw_block(a.seed, g, t >> 1, kStreamStep);
                const uint32_t wt = (t & 1) ? w.z : w.x, wp = (t & 1) ? w.w : w.y;
                if (a.action == nullptr) act[j] = static_cast<int32_t>(wp >> 30);
                if (SLOTS == 4 && a.draw == nullptr) r[j] = wt;
            }
            const uint32_t f = transition<SMEM, SLOTS>(a, tab, g, pos[j], len[j], ret[j], act[j], r[j], rew[j], errs);
            term[j] = f & 1u;
            trunc[j] = (f >> 1) & 1u;
            fobs[2 * j] = pos[j] / a.cols;
            fobs[2 * j + 1] = pos[j] - fobs[2 * j] * a.cols;
            const bool valid = base + j < a.n;
            if (valid && !(f & 4u)) acc.steps += 1u;
            if (a.autoreset && (f & 3u) && valid) {
                acc.finish(len[j], ret[j]);
                pos[j] = a.start_cell;
                len[j] = 0;
                ret[j] = 0.0f;
                obs[2 * j] = a.start_cell / a.cols;
                obs[2 * j + 1] = a.start_cell - obs[2 * j] * a.cols;
            } else {
                obs[2 * j] = fobs[2 * j];
                obs[2 * j + 1] = fobs[2 * j + 1];
            }
        }
        if (full) {  // 128-bit stores
            *reinterpret_cast<int4*>(a.pos + base) = make_int4(pos[0], pos[1], pos[2], pos[3]);
            *reinterpret_cast<int4*>(a.ep_len + base) = make_int4(len[0], len[1], len[2], len[3]);
            *reinterpret_cast<float4*>(a.ep_ret + base) = make_float4(ret[0], ret[1], ret[2], ret[3]);
            if (a.obs) {
                *reinterpret_cast<int4*>(a.obs + 2 * base) = make_int4(obs[0], obs[1], obs[2], obs[3]);
                *reinterpret_cast<int4*>(a.obs + 2 * base + 4) = make_int4(obs[4], obs[5], obs[6], obs[7]);
            }
            if (a.final_obs) {
                *reinterpret_cast<int4*>(a.final_obs + 2 * base) = make_int4(fobs[0], fobs[1], fobs[2], fobs[3]);
                *reinterpret_cast<int4*>(a.final_obs + 2 * base + 4) = make_int4(fobs[4], fobs[5], fobs[6], fobs[7]);
            }
            if (a.reward) *reinterpret_cast<float4*>(a.reward + base) = make_float4(rew[0], rew[1], rew[2], rew[3]);
            if (a.terminated) *reinterpret_cast<uchar4*>(a.terminated + base) = make_uchar4(term[0], term[1], term[2], term[3]);
            if (a.truncated) *reinterpret_cast<uchar4*>(a.truncated + base) = make_uchar4(trunc[0], trunc[1], trunc[2], trunc[3]);
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                if (base + j >= a.n) break;
                const int64_t i = base + j;
                a.pos[i] = pos[j];
                a.ep_len[i] = len[j];
                a.ep_ret[i] = ret[j];
                if (a.obs) { a.obs[2 * i] = obs[2 * j]; a.obs[2 * i + 1] = obs[2 * j + 1]; }
                if (a.final_obs) { a.final_obs[2 * i] = fobs[2 * j]; a.final_obs[2 * i + 1] = fobs[2 * j + 1]; }
                if (a.reward) a.reward[i] = rew[j];
                if (a.terminated) a.terminated[i] = term[j];
                if (a.truncated) a.truncated[i] = trunc[j];
            }
        }
    }
    flush_stats(acc, a.stats);
    flush_errors(errs, a.err);
    clock_advance(a.clock, 1ull);
}

// ---- fused K-step rollout -------------------------------------------------------------------------------------------
// One env per thread, state in registers for all K steps.  A warp votes (ballot) on "any lane finished an episode";
// the reset / statistics path runs only in warps where the vote is non-zero.
template <bool SMEM, int SLOTS, bool EXT_ACTIONS, bool EMIT>
__global__ void __launch_bounds__(kThreads) gridworld_rollout_kernel(const GridArgs a) {
    const uint2* __restrict__ tab = stage_table<SMEM>(a);
    EpisodeAcc acc;
    ErrorAcc errs;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t n_round = (a.n + 31) & ~static_cast<int64_t>(31);  // warp-uniform trip count
    const uint64_t t0 = a.t + clock_read(a.clock);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool valid = i < a.n;
        const uint32_t g = static_cast<uint32_t>(a.env_offset + i);
        int32_t pos = valid ? a.pos[i] : a.start_cell;
        int32_t len = valid ? a.ep_len[i] : 0;
        float ret = valid ? a.ep_ret[i] : 0.0f;
        uint4 w = make_uint4(0, 0, 0, 0);
        constexpr bool kNeedWords = !EXT_ACTIONS || SLOTS == 4;
        for (int k = 0; k < a.K; ++k) {
            const uint64_t t = t0 + static_cast<uint64_t>(k);
            uint32_t wt = 0, wp = 0;
            if constexpr (kNeedWords) {
                if (k == 0 || !(t & 1)) w = draw_block(a.seed, g, t >> 1, kStreamStep);
                wt = (t & 1) ? w.z : w.x;
                wp = (t & 1) ? w.w : w.y;
            }
            int32_t act;
            if constexpr (EXT_ACTIONS) act = valid ? __ldg(a.action + static_cast<int64_t>(k) * a.n + i) : 0;
            else act = static_cast<int32_t>(wp >> 30);
            float rew;
            const uint32_t f = transition<SMEM, SLOTS>(a, tab, g, pos, len, ret, act, wt, rew, errs);
            if constexpr (EMIT) {
                if (valid) {
                    if (a.reward) a.reward[static_cast<int64_t>(k) * a.n + i] = rew;
                    if (a.flags) a.flags[static_cast<int64_t>(k) * a.n + i] = static_cast<uint8_t>(f & 3u);
                }
            }
            if (valid && !(f & 4u)) acc.steps += 1u;
            const bool fin = valid && (f & 3u);
            if (__ballot_sync(0xffffffffu, fin)) {
                if (fin) {
                    acc.finish(len, ret);
                    pos = a.start_cell;
                    len = 0;
                    ret = 0.0f;
                }
            }
        }
        if (valid) {
            a.pos[i] = pos;
            a.ep_len[i] = len;
            a.ep_ret[i] = ret;
        }
    }
    flush_stats(acc, a.stats);
    flush_errors(errs, a.err);
    clock_advance(a.clock, static_cast<unsigned long long>(a.K));
}

__global__ void __launch_bounds__(kThreads) gridworld_reset_kernel(int32_t start_cell, int32_t cols, int32_t* pos,
                                                                   int32_t* ep_len, float* ep_ret, int32_t* obs,
                                                                   const uint8_t* mask, int64_t n) {
    const int32_t r0 = start_cell / cols, c0 = start_cell - r0 * cols;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        if (mask && !mask[i]) continue;
        pos[i] = start_cell;
        ep_len[i] = 0;
        ep_ret[i] = 0.0f;
        if (obs) { obs[2 * i] = r0; obs[2 * i + 1] = c0; }
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------------
constexpr size_t kMaxSmemTable = 96 * 1024;

static int fill_args(const ef_grid_desc* g, GridArgs& a) {
    if (!g || !g->table) return EF_ERR_INVALID_ARGUMENT;
    if (g->rows <= 0 || g->cols <= 0 || (g->n_slots != 1 && g->n_slots != 4)) return EF_ERR_INVALID_ARGUMENT;
    const int64_t cells = static_cast<int64_t>(g->rows) * g->cols;
    if (cells > 65536) return EF_ERR_UNSUPPORTED;
    if (g->start_cell < 0 || g->start_cell >= (g->table_cells > 0 ? g->table_cells : cells) || g->idx_cap < 0 || g->idx_cap > 3) return EF_ERR_INVALID_ARGUMENT;
    if (!aligned(g->table, 8)) return EF_ERR_INVALID_ARGUMENT;
    a.table = reinterpret_cast<const uint2*>(g->table);
    const int64_t tcells = g->table_cells > 0 ? g->table_cells : cells;
    if (tcells < cells || tcells > 65536) return EF_ERR_INVALID_ARGUMENT;
    a.table_len = static_cast<int32_t>(tcells * 4 * g->n_slots);
    a.cols = g->cols;
    a.thr0 = g->thr[0];
    a.thr1 = g->thr[1];
    a.thr2 = g->thr[2];
    a.idx_cap = g->idx_cap;
    a.start_cell = g->start_cell;
    a.limit = g->episode_limit;
    return EF_OK;
}

template <typename Kernel>
static int launch(Kernel kernel, const GridArgs& a, bool smem, int grid, cudaStream_t stream) {
    const size_t bytes = smem ? static_cast<size_t>(a.table_len) * sizeof(uint2) : 0;
    if (bytes > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
        if (e != cudaSuccess) return cuda_status(e);
    }
    kernel<<<grid, kThreads, bytes, stream>>>(a);
    return cuda_status(cudaGetLastError());
}

}  // namespace ef

using namespace ef;

extern "C" int ef_gridworld_step(const ef_grid_desc* grid, int32_t* pos, int32_t* ep_len, float* ep_ret,
                                 const int32_t* action, const uint32_t* draw, uint64_t seed, uint64_t t,
                                 uint64_t env_offset, int32_t* obs, float* reward, uint8_t* terminated,
                                 uint8_t* truncated, int32_t* final_obs, double* stats, uint32_t* err,
                                 uint64_t* clock, int64_t n, int32_t autoreset, void* stream) {
    GridArgs a{};
    if (int s = fill_args(grid, a)) return s;
    if (n < 0 || !pos || !ep_len || !ep_ret) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    a.pos = pos; a.ep_len = ep_len; a.ep_ret = ep_ret; a.action = action; a.draw = draw;
    a.seed = seed; a.t = t; a.env_offset = env_offset;
    a.obs = obs; a.reward = reward; a.terminated = terminated; a.truncated = truncated; a.final_obs = final_obs;
    a.stats = stats; a.err = err; a.n = n; a.autoreset = autoreset;
    a.clock = reinterpret_cast<unsigned long long*>(clock);
    const bool vec = aligned(pos, 16) && aligned(ep_len, 16) && aligned(ep_ret, 16) && aligned(action, 16) &&
                     aligned(draw, 16) && aligned(obs, 16) && aligned(reward, 16) && aligned(terminated, 4) &&
                     aligned(truncated, 4) && aligned(final_obs, 16);
    const bool smem = static_cast<size_t>(a.table_len) * sizeof(uint2) <= kMaxSmemTable;
    const cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid_v = grid_for((n + 3) / 4, 4), grid_s = grid_for(n, 8);
#define EF_DISPATCH(SM, SL)                                                                                   \
    return vec ? launch(gridworld_step_kernel<SM, SL, 4>, a, SM, grid_v, st)                                   \
               : launch(gridworld_step_kernel<SM, SL, 1>, a, SM, grid_s, st)
    if (smem) {
        if (grid->n_slots == 1) { EF_DISPATCH(true, 1); } else { EF_DISPATCH(true, 4); }
    } else {
        if (grid->n_slots == 1) { EF_DISPATCH(false, 1); } else { EF_DISPATCH(false, 4); }
    }
#undef EF_DISPATCH
}

extern "C" int ef_gridworld_rollout(const ef_grid_desc* grid, int32_t* pos, int32_t* ep_len, float* ep_ret,
                                    const int32_t* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K,
                                    float* rewards, uint8_t* flags, double* stats, uint32_t* err, uint64_t* clock,
                                    int64_t n, void* stream) {
    GridArgs a{};
    if (int s = fill_args(grid, a)) return s;
    if (n < 0 || K < 0 || !pos || !ep_len || !ep_ret) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0 || K == 0) return EF_OK;
    a.pos = pos; a.ep_len = ep_len; a.ep_ret = ep_ret; a.action = actions;
    a.seed = seed; a.t = t0; a.env_offset = env_offset; a.K = K;
    a.reward = rewards; a.flags = flags; a.stats = stats; a.err = err; a.n = n; a.autoreset = 1;
    a.clock = reinterpret_cast<unsigned long long*>(clock);
    const bool smem = static_cast<size_t>(a.table_len) * sizeof(uint2) <= kMaxSmemTable;
    const bool ext = actions != nullptr, emit = rewards != nullptr || flags != nullptr;
    const cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid_r = grid_for(n, 8);
#define EF_DISPATCH2(SM, SL)                                                                                  \
    do {                                                                                                      \
        if (ext && emit) return launch(gridworld_rollout_kernel<SM, SL, true, true>, a, SM, grid_r, st);     \
        if (ext) return launch(gridworld_rollout_kernel<SM, SL, true, false>, a, SM, grid_r, st);            \
        if (emit) return launch(gridworld_rollout_kernel<SM, SL, false, true>, a, SM, grid_r, st);           \
        return launch(gridworld_rollout_kernel<SM, SL, false, false>, a, SM, grid_r, st);                    \
    } while (0)
    if (smem) {
        if (grid->n_slots == 1) EF_DISPATCH2(true, 1); else EF_DISPATCH2(true, 4);
    } else {
        if (grid->n_slots == 1) EF_DISPATCH2(false, 1); else EF_DISPATCH2(false, 4);
    }
#undef EF_DISPATCH2
}

extern "C" int ef_gridworld_reset(int32_t start_cell, int32_t cols, int32_t* pos, int32_t* ep_len, float* ep_ret,
                                  int32_t* obs, const uint8_t* mask, int64_t n, void* stream) {
    if (n < 0 || cols <= 0 || start_cell < 0 || !pos || !ep_len || !ep_ret) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    gridworld_reset_kernel<<<grid_for(n, 8), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        start_cell, cols, pos, ep_len, ep_ret, obs, mask, n);
    return cuda_status(cudaGetLastError());
}
