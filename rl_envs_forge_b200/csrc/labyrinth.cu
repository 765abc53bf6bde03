// envforge_b200 -- Labyrinth batched step + observation kernels (sm_100a).
//
// Replaces, for N independent players on M host-generated mazes, the reference's per-instance
//   Labyrinth.step                      rl_envs_forge/envs/labyrinth/labyrinth.py:245-288
//   is_valid_move / agent_move          :290-309 / :333-345,  Player.potential_next_position  entities/player.py:61-82
//   state / build_state_matrix / build_state_around_the_player   :195-243
// Maze generation (maze/*.py) stays on the host: a maze arrives as its finished uint8 grid plus start / target cells.
//
// A CTA owns 256 consecutive envs.  Phase 1: one thread per env does the transition (bounds + wall check is one byte
// gathered from the maze grid, which M << N mazes keep L2-resident).  Phase 2: the CTA renders the observations of its
// envs -- the dominant cost: rows*cols bytes per env-step for the full view, (2v+1)^2 for a vision window -- with warps
// walking envs and lanes walking the bytes of one env, so every global store is a contiguous run.  HBM-bound on the
// observation write (full view) or issue-bound on the gather (small vision windows).
#include <algorithm>

#include "ef_common.cuh"

namespace ef {

constexpr uint32_t kLabWall = 0u, kLabTarget = 2u, kLabPlayer = 4u;  // labyrinth/constants.py:12-16
constexpr int kLabEnvsPerCta = kThreads;
constexpr int kLabMaxIters = 8;  // vision windows up to 16x16 bytes keep their (row, col) offsets in registers

struct LabArgs {
    const uint8_t* grids;        // [M, cells]
    const int32_t* maze_index;   // [n] or null: maze of env g = g % n_mazes
    const int32_t* start_cell;   // [M]
    const int32_t* target_cell;  // [M]
    int32_t rows, cols, cells, n_mazes, vision, limit;
    float r_neutral, r_wall, r_target;
    int32_t* pos;
    int32_t* ep_len;
    float* ep_ret;
    const int32_t* action;
    uint64_t seed, t, env_offset;
    uint8_t* obs;
    float* reward;
    uint8_t* terminated;
    uint8_t* truncated;
    uint8_t* final_obs;
    double* stats;
    uint32_t* err;
    unsigned long long* clock;
    int64_t n;
    int32_t autoreset;
    int32_t vision_smem;  // 1: vision windows are assembled per thread in shared memory (kLabEnvsPerCta * d * d bytes)
    // rollout only
    int32_t K;
    uint8_t* flags;
};

__device__ __forceinline__ int32_t lab_maze_of(const LabArgs& a, int64_t i) {
    return a.maze_index ? __ldg(a.maze_index + i) : static_cast<int32_t>((a.env_offset + i) % static_cast<uint64_t>(a.n_mazes));
}

// Full view of the envs [e0, e0 + count): grid bytes with TARGET, then PLAYER, written over them (build_state_matrix).
// CH = bytes per access (16 / 4 / 1, cells % CH == 0): warp w takes envs w, w + 8, ...; lanes walk the env's chunks.
template <int CH>
__device__ __forceinline__ void render_full(const LabArgs& a, uint8_t* __restrict__ out, int64_t e0, int count,
                                            const int32_t* s_pos, const int32_t* s_maze, const int32_t* s_tgt) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunks = a.cells / CH;
    for (int e = warp; e < count; e += kWarpsPerCta) {
        const uint32_t pos = static_cast<uint32_t>(s_pos[e]), tgt = static_cast<uint32_t>(s_tgt[e]);
        const uint8_t* __restrict__ src = a.grids + static_cast<int64_t>(s_maze[e]) * a.cells;
        uint8_t* __restrict__ dst = out + (e0 + e) * a.cells;
        for (int c = lane; c < chunks; c += 32) {
            if constexpr (CH == 1) {
                uint32_t v = __ldg(src + c);
                if (static_cast<uint32_t>(c) == tgt) v = kLabTarget;
                if (static_cast<uint32_t>(c) == pos) v = kLabPlayer;
                dst[c] = static_cast<uint8_t>(v);
            } else {
                constexpr int W = CH / 4;
                uint32_t w[W];
                if constexpr (CH == 16) {
                    const uint4 q = __ldg(reinterpret_cast<const uint4*>(src) + c);
                    w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
                } else {
                    w[0] = __ldg(reinterpret_cast<const uint32_t*>(src) + c);
                }
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    const uint32_t word = static_cast<uint32_t>(c) * W + k;  // index of this 4-byte word within the env
                    if ((tgt >> 2) == word) w[k] = (w[k] & ~(0xffu << (8 * (tgt & 3u)))) | (kLabTarget << (8 * (tgt & 3u)));
                    if ((pos >> 2) == word) w[k] = (w[k] & ~(0xffu << (8 * (pos & 3u)))) | (kLabPlayer << (8 * (pos & 3u)));
                }
                if constexpr (CH == 16) reinterpret_cast<uint4*>(dst)[c] = make_uint4(w[0], w[1], w[2], w[3]);
                else reinterpret_cast<uint32_t*>(dst)[c] = w[0];
            }
        }
    }
}

// Vision window (build_state_around_the_player): d x d bytes centred on the player, WALL outside the maze.
__device__ __forceinline__ void render_vision(const LabArgs& a, uint8_t* __restrict__ out, int64_t e0, int count,
                                              const int32_t* s_pos, const int32_t* s_maze, const int32_t* s_tgt) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int v = a.vision, d = 2 * v + 1, bytes = d * d;
    const int iters = (bytes + 31) / 32;
    int32_t off[kLabMaxIters];  // (di << 16) | dj of byte k = it * 32 + lane, relative to the window's corner
#pragma unroll
    for (int it = 0; it < kLabMaxIters; ++it) {
        const int k = it * 32 + lane;
        off[it] = ((k / d) << 16) | (k % d);
    }
    for (int e = warp; e < count; e += kWarpsPerCta) {
        const int32_t pos = s_pos[e], tgt = s_tgt[e];
        const int32_t pr = pos / a.cols, pc = pos - pr * a.cols;
        const uint8_t* __restrict__ src = a.grids + static_cast<int64_t>(s_maze[e]) * a.cells;
        uint8_t* __restrict__ dst = out + (e0 + e) * bytes;
        auto emit = [&](int k, int di, int dj) {
            const int r = pr - v + di, c = pc - v + dj;
            uint32_t val = kLabWall;
            if (r >= 0 && r < a.rows && c >= 0 && c < a.cols) {
                const int32_t cell = r * a.cols + c;
                val = __ldg(src + cell);
                if (cell == tgt) val = kLabTarget;
                if (cell == pos) val = kLabPlayer;
            }
            dst[k] = static_cast<uint8_t>(val);
        };
        if (iters <= kLabMaxIters) {
#pragma unroll
            for (int it = 0; it < kLabMaxIters; ++it) {
                const int k = it * 32 + lane;
                if (it < iters && k < bytes) emit(k, off[it] >> 16, off[it] & 0xffff);
            }
        } else {
            for (int k = lane; k < bytes; k += 32) emit(k, k / d, k % d);
        }
    }
}

// Vision window, one THREAD per player: the thread walks its d x d window with two nested loops (no per-byte index
// division, ~8 instructions per byte instead of ~100 in the lane-per-byte version above), assembling it in shared
// memory; the CTA then streams its contiguous count * d * d output bytes to global memory with 16-byte stores.
__device__ __forceinline__ void render_vision_smem(const LabArgs& a, uint8_t* __restrict__ out, int64_t e0, int count,
                                                   const int32_t* s_pos, const int32_t* s_maze, const int32_t* s_tgt,
                                                   uint8_t* __restrict__ s_obs) {
    const int v = a.vision, d = 2 * v + 1, bytes = d * d;
    if (static_cast<int>(threadIdx.x) < count) {
        const int e = threadIdx.x;
        const int32_t pos = s_pos[e], tgt = s_tgt[e];
        const int32_t pr = pos / a.cols, pc = pos - pr * a.cols;
        const uint8_t* __restrict__ src = a.grids + static_cast<int64_t>(s_maze[e]) * a.cells;
        uint8_t* __restrict__ dst = s_obs + e * bytes;
        for (int di = 0; di < d; ++di) {
            const int r = pr - v + di;
            const bool row_ok = r >= 0 && r < a.rows;
            const int32_t row_cell = r * a.cols;
            for (int dj = 0; dj < d; ++dj) {
                const int c = pc - v + dj;
                uint32_t val = kLabWall;
                if (row_ok && c >= 0 && c < a.cols) {
                    const int32_t cell = row_cell + c;
                    val = __ldg(src + cell);
                    if (cell == tgt) val = kLabTarget;
                    if (cell == pos) val = kLabPlayer;
                }
                dst[di * d + dj] = static_cast<uint8_t>(val);
            }
        }
    }
    __syncthreads();
    const int total = count * bytes;
    uint8_t* __restrict__ g = out + e0 * bytes;  // e0 is a multiple of 256, so g keeps the 16-byte alignment of `out`
    if (aligned(out, 16)) {
        const int vecs = total >> 4;
        for (int i = threadIdx.x; i < vecs; i += kThreads)
            reinterpret_cast<uint4*>(g)[i] = reinterpret_cast<const uint4*>(s_obs)[i];
        for (int i = (vecs << 4) + threadIdx.x; i < total; i += kThreads) g[i] = s_obs[i];
    } else {
        for (int i = threadIdx.x; i < total; i += kThreads) g[i] = s_obs[i];
    }
    __syncthreads();
}

__device__ __forceinline__ void lab_render(const LabArgs& a, uint8_t* out, int64_t e0, int count, const int32_t* s_pos,
                                           const int32_t* s_maze, const int32_t* s_tgt) {
    extern __shared__ __align__(16) uint8_t s_obs[];
    if (a.vision >= 0 && a.vision_smem) {
        render_vision_smem(a, out, e0, count, s_pos, s_maze, s_tgt, s_obs);
    } else if (a.vision >= 0) {
        render_vision(a, out, e0, count, s_pos, s_maze, s_tgt);
    } else if ((a.cells & 15) == 0 && aligned(a.grids, 16) && aligned(out, 16)) {
        render_full<16>(a, out, e0, count, s_pos, s_maze, s_tgt);
    } else if ((a.cells & 3) == 0 && aligned(a.grids, 4) && aligned(out, 4)) {
        render_full<4>(a, out, e0, count, s_pos, s_maze, s_tgt);
    } else {
        render_full<1>(a, out, e0, count, s_pos, s_maze, s_tgt);
    }
}

// STEP == false: observation only (reset / `state` property); the positions are read, nothing else is touched.
template <bool STEP>
__global__ void __launch_bounds__(kThreads) labyrinth_kernel(const LabArgs a) {
    __shared__ int32_t s_pos[kLabEnvsPerCta], s_fpos[kLabEnvsPerCta], s_maze[kLabEnvsPerCta], s_tgt[kLabEnvsPerCta];
    pdl_launch_dependents();
    pdl_wait();
    EpisodeAcc acc;
    ErrorAcc errs;
    unsigned long long ticket;
    const uint64_t t = a.t + clock_begin(a.clock, ticket, STEP);
    for (int64_t e0 = static_cast<int64_t>(blockIdx.x) * kLabEnvsPerCta; e0 < a.n;
         e0 += static_cast<int64_t>(gridDim.x) * kLabEnvsPerCta) {
        const int count = static_cast<int>(min(static_cast<int64_t>(kLabEnvsPerCta), a.n - e0));
        const int64_t i = e0 + threadIdx.x;
        if (threadIdx.x < count) {
            const int32_t maze = lab_maze_of(a, i);
            const int32_t tgt = __ldg(a.target_cell + maze);
            int32_t pos = a.pos[i];
            int32_t fpos = pos;
            if constexpr (STEP) {
                const uint32_t g = static_cast<uint32_t>(a.env_offset + i);
                int32_t len = a.ep_len[i];
                float ret = a.ep_ret[i], rew = 0.0f;
                const int32_t act = a.action ? __ldg(a.action + i)
                                             : static_cast<int32_t>(group_word1(a.seed, g, t, kStreamPolicy) >> 30);
                uint32_t f = 0u;
                if (static_cast<uint32_t>(act) > 3u) {  // Action(action) raises ValueError (labyrinth.py:261)
                    errs.record(g, static_cast<uint32_t>(act), EF_STEP_BAD_ACTION);
                    f = 4u;
                } else {
                    const int32_t pr = pos / a.cols, pc = pos - pr * a.cols;
                    const int32_t r = pr + (act == 2) - (act == 0), c = pc + (act == 1) - (act == 3);  // player.py:72-80
                    bool valid = r >= 0 && r < a.rows && c >= 0 && c < a.cols;                           // :300-305
                    const int32_t nxt = r * a.cols + c;
                    if (valid) valid = __ldg(a.grids + static_cast<int64_t>(maze) * a.cells + nxt) != kLabWall;  // :306-308
                    rew = a.r_wall;
                    if (valid) {
                        pos = nxt;
                        rew = a.r_neutral;
                        if (pos == tgt) {  // :279-286
                            rew = a.r_target;
                            f = 1u;
                        }
                    }
                    ret += rew;
                    len += 1;
                    if (a.limit > 0 && len >= a.limit) f |= 2u;
                    acc.steps += 1u;
                }
                fpos = pos;
                if (a.autoreset && (f & 3u)) {
                    acc.finish(len, ret);
                    pos = __ldg(a.start_cell + maze);
                    len = 0;
                    ret = 0.0f;
                }
                a.pos[i] = pos;
                a.ep_len[i] = len;
                a.ep_ret[i] = ret;
                if (a.reward) a.reward[i] = rew;
                if (a.terminated) a.terminated[i] = static_cast<uint8_t>(f & 1u);
                if (a.truncated) a.truncated[i] = static_cast<uint8_t>((f >> 1) & 1u);
            }
            s_pos[threadIdx.x] = pos;
            s_fpos[threadIdx.x] = fpos;
            s_maze[threadIdx.x] = maze;
            s_tgt[threadIdx.x] = tgt;
        }
        __syncthreads();
        if (a.obs) lab_render(a, a.obs, e0, count, s_pos, s_maze, s_tgt);
        if (STEP && a.final_obs) lab_render(a, a.final_obs, e0, count, s_fpos, s_maze, s_tgt);
        __syncthreads();
    }
    if constexpr (STEP) cta_epilogue(acc, errs, a.stats, a.err, a.clock, ticket, 1ull);
}

// One transition of one player (labyrinth.py:245-288).  Returns bit0 terminated, bit2 rejected (invalid Action).
__device__ __forceinline__ uint32_t lab_transition(const LabArgs& a, const uint8_t* __restrict__ grid, int32_t tgt, uint32_t g,
                                                   int32_t& pos, int32_t act, float& rew, ErrorAcc& errs) {
    rew = 0.0f;
    if (static_cast<uint32_t>(act) > 3u) {  // Action(action) raises ValueError (:261)
        errs.record(g, static_cast<uint32_t>(act), EF_STEP_BAD_ACTION);
        return 4u;
    }
    const int32_t pr = pos / a.cols, pc = pos - pr * a.cols;
    const int32_t r = pr + (act == 2) - (act == 0), c = pc + (act == 1) - (act == 3);  // player.py:72-80
    bool ok = r >= 0 && r < a.rows && c >= 0 && c < a.cols;                              // :300-305
    const int32_t nxt = r * a.cols + c;
    if (ok) ok = __ldg(grid + nxt) != kLabWall;                                           // :306-308
    rew = a.r_wall;
    if (!ok) return 0u;
    pos = nxt;
    rew = a.r_neutral;
    if (pos == tgt) {  // :279-286
        rew = a.r_target;
        return 1u;
    }
    return 0u;
}

// Fused K-step rollout: one player per thread, position and counters in registers for all K steps, no observation
// rendering (call ef_labyrinth_observe afterwards).  `action` [K, n] or NULL (uniform random direction), `reward` [K, n] /
// `flags` [K, n] (bit0 terminated, bit1 truncated) optional; auto-reset to the maze's start cell always on.
__global__ void __launch_bounds__(kThreads) labyrinth_rollout_kernel(const LabArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    EpisodeAcc acc;
    ErrorAcc errs;
    unsigned long long ticket;
    const uint64_t t0 = a.t + clock_begin(a.clock, ticket);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint32_t g = static_cast<uint32_t>(a.env_offset + i);
        const int32_t maze = lab_maze_of(a, i);
        const int32_t tgt = __ldg(a.target_cell + maze), start = __ldg(a.start_cell + maze);
        const uint8_t* __restrict__ grid = a.grids + static_cast<int64_t>(maze) * a.cells;
        int32_t pos = a.pos[i], len = a.ep_len[i];
        float ret = a.ep_ret[i];
        for (int k = 0; k < a.K; ++k) {
            const uint64_t t = t0 + static_cast<uint64_t>(k);
            const int64_t o = static_cast<int64_t>(k) * a.n + i;
            const int32_t act = a.action ? __ldg(a.action + o)
                                         : static_cast<int32_t>(group_word1(a.seed, g, t, kStreamPolicy) >> 30);
            float rew;
            uint32_t f = lab_transition(a, grid, tgt, g, pos, act, rew, errs);
            if (!(f & 4u)) {
                ret += rew;
                len += 1;
                acc.steps += 1u;
                if (a.limit > 0 && len >= a.limit) f |= 2u;
            }
            if (a.reward) a.reward[o] = rew;
            if (a.flags) a.flags[o] = static_cast<uint8_t>(f & 3u);
            if (f & 3u) {
                acc.finish(len, ret);
                pos = start;
                len = 0;
                ret = 0.0f;
            }
        }
        a.pos[i] = pos;
        a.ep_len[i] = len;
        a.ep_ret[i] = ret;
    }
    cta_epilogue(acc, errs, a.stats, a.err, a.clock, ticket, static_cast<unsigned long long>(a.K));
}

__global__ void __launch_bounds__(kThreads) labyrinth_reset_kernel(const LabArgs a, const uint8_t* mask) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        if (mask && !mask[i]) continue;
        a.pos[i] = __ldg(a.start_cell + lab_maze_of(a, i));
        a.ep_len[i] = 0;
        a.ep_ret[i] = 0.0f;
    }
}

static int fill_lab_args(const ef_labyrinth_desc* d, LabArgs& a) {
    if (!d || !d->grids || !d->start_cell || !d->target_cell) return EF_ERR_INVALID_ARGUMENT;
    if (d->rows <= 0 || d->cols <= 0 || d->n_mazes <= 0) return EF_ERR_INVALID_ARGUMENT;
    if (static_cast<int64_t>(d->rows) * d->cols > (1 << 24)) return EF_ERR_UNSUPPORTED;
    if (d->vision_range > 16383) return EF_ERR_UNSUPPORTED;
    a.grids = d->grids;
    a.maze_index = d->maze_index;
    a.start_cell = d->start_cell;
    a.target_cell = d->target_cell;
    a.rows = d->rows;
    a.cols = d->cols;
    a.cells = d->rows * d->cols;
    a.n_mazes = d->n_mazes;
    a.vision = d->vision_range > 0 ? d->vision_range : -1;
    a.limit = d->episode_limit;
    a.r_neutral = d->neutral_reward;
    a.r_wall = d->wall_collision_reward;
    a.r_target = d->target_reached_reward;
    return EF_OK;
}

// Shared memory for the thread-per-player vision path: used while the CTA's windows fit the default 48 KB window.
static size_t lab_vision_smem(LabArgs& a) {
    a.vision_smem = 0;
    if (a.vision < 0) return 0;
    const size_t bytes = static_cast<size_t>(kLabEnvsPerCta) * (2 * a.vision + 1) * (2 * a.vision + 1);
    if (bytes > 48 * 1024) return 0;
    a.vision_smem = 1;
    return bytes;
}

static int lab_grid(int64_t n) {
    const int64_t ctas = (n + kLabEnvsPerCta - 1) / kLabEnvsPerCta;
    return static_cast<int>(std::min<int64_t>(ctas, static_cast<int64_t>(sm_count()) * 64));
}

}  // namespace ef

using namespace ef;

extern "C" int ef_labyrinth_step(const ef_labyrinth_desc* maze, int32_t* pos, int32_t* ep_len, float* ep_ret,
                                 const int32_t* action, uint64_t seed, uint64_t t, uint64_t env_offset, uint8_t* obs,
                                 float* reward, uint8_t* terminated, uint8_t* truncated, uint8_t* final_obs, double* stats,
                                 uint32_t* err, uint64_t* clock, int64_t n, int32_t autoreset, void* stream) {
    LabArgs a{};
    if (int s = fill_lab_args(maze, a)) return s;
    if (n < 0 || !pos || !ep_len || !ep_ret) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    a.pos = pos; a.ep_len = ep_len; a.ep_ret = ep_ret; a.action = action;
    a.seed = seed; a.t = t; a.env_offset = env_offset;
    a.obs = obs; a.reward = reward; a.terminated = terminated; a.truncated = truncated; a.final_obs = final_obs;
    a.stats = stats; a.err = err; a.clock = reinterpret_cast<unsigned long long*>(clock); a.n = n; a.autoreset = autoreset;
    const size_t smem = lab_vision_smem(a);
    return launch_kernel(labyrinth_kernel<true>, a, lab_grid(n), smem, static_cast<cudaStream_t>(stream));
}

extern "C" int ef_labyrinth_rollout(const ef_labyrinth_desc* maze, int32_t* pos, int32_t* ep_len, float* ep_ret,
                                    const int32_t* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K,
                                    float* rewards, uint8_t* flags, double* stats, uint32_t* err, uint64_t* clock,
                                    int64_t n, void* stream) {
    LabArgs a{};
    if (int s = fill_lab_args(maze, a)) return s;
    if (n < 0 || K < 0 || !pos || !ep_len || !ep_ret) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0 || K == 0) return EF_OK;
    a.pos = pos; a.ep_len = ep_len; a.ep_ret = ep_ret; a.action = actions;
    a.seed = seed; a.t = t0; a.env_offset = env_offset; a.K = K; a.reward = rewards; a.flags = flags;
    a.stats = stats; a.err = err; a.clock = reinterpret_cast<unsigned long long*>(clock); a.n = n; a.autoreset = 1;
    return launch_kernel(labyrinth_rollout_kernel, a, grid_for(n, 8), 0, static_cast<cudaStream_t>(stream));
}

extern "C" int ef_labyrinth_observe(const ef_labyrinth_desc* maze, const int32_t* pos, uint64_t env_offset, uint8_t* obs,
                                    int64_t n, void* stream) {
    LabArgs a{};
    if (int s = fill_lab_args(maze, a)) return s;
    if (n < 0 || !pos || !obs) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    a.pos = const_cast<int32_t*>(pos); a.env_offset = env_offset; a.obs = obs; a.n = n;
    const size_t smem = lab_vision_smem(a);
    return launch_kernel(labyrinth_kernel<false>, a, lab_grid(n), smem, static_cast<cudaStream_t>(stream));
}

extern "C" int ef_labyrinth_reset(const ef_labyrinth_desc* maze, int32_t* pos, int32_t* ep_len, float* ep_ret,
                                  uint64_t env_offset, const uint8_t* mask, int64_t n, void* stream) {
    LabArgs a{};
    if (int s = fill_lab_args(maze, a)) return s;
    if (n < 0 || !pos || !ep_len || !ep_ret) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    a.pos = pos; a.ep_len = ep_len; a.ep_ret = ep_ret; a.env_offset = env_offset; a.n = n;
    labyrinth_reset_kernel<<<grid_for(n, 8), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, mask);
    return cuda_status(cudaGetLastError());
}
