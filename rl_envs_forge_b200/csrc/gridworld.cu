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
// written (pos, ep_len, ep_ret, obs[2], reward, terminated, truncated) = 42 B.  A thread owns four consecutive envs:
// every access is a 128-bit vector load/store (flags: one 32-bit store), and one Philox4x32 call yields the four
// outcome draws.  Grid-stride over a grid sized to the SM count; the table (8 KB for 8x8) is staged once per CTA.
//
// Packed state (ep_len == NULL): position and episode_length_counter share one 32-bit word per env, cell in bits 0-15
// and the counter in bits 16-31.  Possible whenever the counter cannot exceed 65535 (auto-reset on, 0 < limit <= 65535;
// cells are < 65536 by construction of the table), it removes 8 of the 42 bytes an env-step moves.
#include <algorithm>
#include <cstdlib>

#include "ef_common.cuh"

// Work of the step kernel AHEAD of griddepcontrol.wait, and who reads the device clock.  A/B on B200, us per 2^20-env launch in a
// CUDA-graph chain: one env set stepped repeatedly / eight sets in rotation (the headline: cold L2) / one set with a short
// policy kernel ahead of every step (tools/chain_probe.py, tools/policy_loop_probe.py; profiles/r02a*_chain_probe.txt,
// r02at..r02bb_*.txt):
//   nothing ahead of the wait, every thread reads the clock after it                 7.03 / 8.94 / 9.57   (start of the round's second half)
//   + bulk L2 prefetch of the CTA's inputs                                           7.02 / 7.88 /  --
//   + ONE thread per CTA reads the clock after the wait (barrier inside the loads)   6.92 / 7.57 / 9.53
//   + the clock word's line prefetched into L2 ahead of the wait                     6.99 / 7.02 / 9.70   <- COLD instantiation (EF_STEP_COLD_INPUTS)
//   + the clock word LOADED (and discarded) by thread 0 ahead of the wait instead    6.76 / 7.30 / 9.58   <- default instantiation
//   + Philox draws of the first 1 / 2 / 4 quads ahead of the wait (validated after)  6.73 / 7.21 / 9.76, 6.75 / 7.23 / --, 7.03 / 7.34 / --
// The early draws were withdrawn: their gain in the chains was the early clock read they came with, and with a short
// kernel ahead a step CTA gets its slot only 0.4 us before the release, so everything ahead of the wait is on the
// critical path there.  Load vs prefetch of the clock word is a template parameter, not a branch: with both in the
// instruction stream neither chain got its better number (6.82 / 7.32 whichever way the branch went).  Not adopted either:
// two launches co-resident (grid 148 x 2 passes, 4 CTAs/SM x Q = 2: 8.5-9.2 us), all draws right after issuing the
// loads (7.4 / 7.5), griddepcontrol.launch_dependents earlier or later (no effect).
#ifndef EF_STEP_L2_PREFETCH
#define EF_STEP_L2_PREFETCH 1   // 1: bulk L2 prefetch of the CTA's input chunk ahead of griddepcontrol.wait; 2: per-line prefetch; 0: none (A/B)
#endif
#ifndef EF_STEP_POST_CLOCK
#define EF_STEP_POST_CLOCK 1    // 1: after the wait, thread 0 reads the clock word and the CTA takes it from shared memory behind a barrier; 0: every thread (A/B)
#endif
#ifndef EF_STEP_PRE_CLOCK
#define EF_STEP_PRE_CLOCK 1     // 1: ahead of the wait thread 0 warms the clock word (load, or line prefetch under EF_STEP_COLD_INPUTS); 0: not (A/B)
#endif
#ifndef EF_STEP_LEAN
#define EF_STEP_LEAN 1       // packed-word arithmetic of step_quad_packed (0: the generic step_quad_compute; same results)
#endif

namespace ef {

struct GridArgs {
    const uint2* table;  // ef_grid_outcome as raw 8-byte words: x = next | flags << 16, y = reward bits
    int32_t table_len;   // entries
    int32_t cols;
    uint32_t cols_magic; // ceil(2^32 / cols): row = umulhi(cell, magic), exact for cell < 65536 (cols > 1)
    uint64_t thr0, thr1, thr2;  // 64-bit: a threshold at or beyond the index cap is 2^32, which no draw reaches
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
    // several layouts in one batch (scalar step kernel only): `table` holds n_layouts consecutive tables of layout_len
    // entries; env i uses layout[i], or (env_offset + i) % n_layouts when layout is null; start_cells[l] replaces start_cell
    const int32_t* layout;
    const int32_t* start_cells;
    int32_t n_layouts, layout_len;
    int64_t quads_per_cta;  // vector step kernel: contiguous quads per CTA
    int32_t cold;           // vector step kernel: EF_STEP_COLD_INPUTS hint (see the kernel's prologue)
    PhiloxKeys rk;          // vector step kernel: round keys of `seed`
    // rollout only
    int32_t K;
    uint8_t* flags;
    unsigned long long* clock;
    // step server only (ef_gridworld_serve)
    uint32_t* act_seq;     // [chunks] written by the action producer: k + 1 once chunk c's actions of step k are in place
    uint32_t* obs_seq;     // [chunks] written by the server: k + 1 once chunk c's outputs of step k are in place
    uint32_t* status;      // set to 1 if a wait gave up
    int32_t chunk_quads;   // quads (of four envs) per chunk = per CTA
};

// Timeline hooks of the step kernels: compiled to nothing in the product build.  `python -m rl_envs_forge_b200._build
// -DEF_TRACE=1 --tag=_trace` builds a tools-only variant (tools/step_trace.py) in which thread 0 of every CTA stores
// %globaltimer / clock64 at the marked points; the results of a step are the same either way.
#ifdef EF_TRACE
__device__ unsigned long long* g_trace_buf = nullptr;   // [8 sets][4 steps][1024 CTAs][16 points][2]
#define EF_TRACE_DECL unsigned long long tr_[32] = {0ull};
#define EF_TRACE_POINT(i)                                                   \
    do {                                                                    \
        if (threadIdx.x == 0) {                                             \
            unsigned long long gt_;                                         \
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt_));          \
            tr_[2 * (i)] = gt_;                                             \
            tr_[2 * (i) + 1] = static_cast<unsigned long long>(clock64());  \
        }                                                                   \
    } while (0)
#define EF_TRACE_FLUSH(t_)                                                                                               \
    do {                                                                                                                 \
        if (threadIdx.x == 0 && g_trace_buf != nullptr && blockIdx.x < 1024) {                                           \
            const unsigned long long set_ = (a.env_offset / static_cast<unsigned long long>(a.n)) & 7ull;                \
            unsigned long long* e_ = g_trace_buf + ((set_ * 4ull + ((t_) & 3ull)) * 1024ull + blockIdx.x) * 32ull;       \
            for (int i_ = 0; i_ < 32; ++i_) e_[i_] = tr_[i_];                                                            \
        }                                                                                                                \
    } while (0)
#else
#define EF_TRACE_DECL
#define EF_TRACE_POINT(i)
#define EF_TRACE_FLUSH(t_)
#endif

// Table staging by TMA bulk copy: thread 0 arms an mbarrier and issues ONE cp.async.bulk for the whole table; the copy
// runs while the CTA goes on (griddepcontrol.wait, input loads), and stage_table_wait() is called right before the first
// table lookup.  Replaces a 256-thread LDG/STS loop (5 % of the step kernel's instructions) that also had to finish
// before anything else could start.  Falls back to the loop when the table is not 16-byte aligned.
template <bool SMEM>
__device__ __forceinline__ const uint2* stage_table_begin(const GridArgs& a, uint64_t* bar, bool& pending) {
    extern __shared__ __align__(16) uint2 s_table[];
    pending = false;
    if constexpr (SMEM) {
        if (aligned(a.table, 16)) {  // table_len is a multiple of 4 entries, i.e. of 32 bytes
            if (threadIdx.x == 0) {
                mbar_init(bar, 1u);
                bulk_copy_g2s(s_table, a.table, static_cast<uint32_t>(a.table_len) * static_cast<uint32_t>(sizeof(uint2)), bar);
            }
            pending = true;
        } else {
            for (int i = threadIdx.x; i < a.table_len; i += blockDim.x) s_table[i] = __ldg(a.table + i);
        }
        __syncthreads();  // the mbarrier (or the table itself) is visible to every thread from here on
        return s_table;
    } else {
        return a.table;
    }
}

__device__ __forceinline__ void stage_table_wait(uint64_t* bar, bool& pending) {
    if (pending) {
        mbar_wait(bar, 0u);
        pending = false;
    }
}

// Synchronous form for the kernels that need the table right away (scalar step, PCG64 step, rollouts).
template <bool SMEM>
__device__ __forceinline__ const uint2* stage_table(const GridArgs& a) {
    __shared__ uint64_t bar;
    bool pending;
    const uint2* tab = stage_table_begin<SMEM>(a, &bar, pending);
    stage_table_wait(&bar, pending);
    return tab;
}

__device__ __forceinline__ int2 cell_to_obs(const GridArgs& a, int32_t cell) {
    const int32_t row = a.cols == 1 ? cell : static_cast<int32_t>(__umulhi(static_cast<uint32_t>(cell), a.cols_magic));
    return make_int2(row, cell - row * a.cols);
}

// Narrow wire format (W8): actions arrive as uint8 [n] and observations leave as uint8 [n,2] (rows, cols <= 256) -- the
// format of a host-side actor loop, where PCIe rather than HBM carries every byte.  `action` / `obs` / `final_obs` of
// GridArgs then point at bytes; everything between the load and the store is the same code.
template <bool W8>
__device__ __forceinline__ int32_t load_action1(const GridArgs& a, int64_t i) {
    if constexpr (W8) return static_cast<int32_t>(__ldg(reinterpret_cast<const uint8_t*>(a.action) + i));
    else return __ldg(a.action + i);
}
template <bool W8>
__device__ __forceinline__ void store_obs1(int32_t* obs, int64_t i, int2 ob) {
    if constexpr (W8) {
        reinterpret_cast<uchar2*>(obs)[i] = make_uchar2(static_cast<uint8_t>(ob.x), static_cast<uint8_t>(ob.y));
    } else {
        obs[2 * i] = ob.x;
        obs[2 * i + 1] = ob.y;
    }
}
// Observations of a quad: two 128-bit stores of int32 pairs, or one 64-bit store of byte pairs.
template <bool W8>
__device__ __forceinline__ void store_obs4(const GridArgs& a, int32_t* obs, int64_t base, const int32_t (&pos)[4]) {
    if constexpr (W8) {
        const int2 o0 = cell_to_obs(a, pos[0]), o1 = cell_to_obs(a, pos[1]);
        const int2 o2 = cell_to_obs(a, pos[2]), o3 = cell_to_obs(a, pos[3]);
        const uint32_t lo = static_cast<uint32_t>(o0.x) | static_cast<uint32_t>(o0.y) << 8 |
                            static_cast<uint32_t>(o1.x) << 16 | static_cast<uint32_t>(o1.y) << 24;
        const uint32_t hi = static_cast<uint32_t>(o2.x) | static_cast<uint32_t>(o2.y) << 8 |
                            static_cast<uint32_t>(o3.x) << 16 | static_cast<uint32_t>(o3.y) << 24;
        *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(obs) + 2 * base) = make_uint2(lo, hi);
    } else {
        int4* o = reinterpret_cast<int4*>(obs + 2 * base);
        const int2 o0 = cell_to_obs(a, pos[0]), o1 = cell_to_obs(a, pos[1]);
        o[0] = make_int4(o0.x, o0.y, o1.x, o1.y);
        const int2 o2 = cell_to_obs(a, pos[2]), o3 = cell_to_obs(a, pos[3]);
        o[1] = make_int4(o2.x, o2.y, o3.x, o3.y);
    }
}

// Scalar state access in either layout (see "Packed state" above).
__device__ __forceinline__ void load_state1(const GridArgs& a, int64_t i, int32_t& pos, int32_t& len) {
    if (a.ep_len) {
        pos = a.pos[i];
        len = a.ep_len[i];
    } else {
        const uint32_t w = static_cast<uint32_t>(a.pos[i]);
        pos = static_cast<int32_t>(w & 0xffffu);
        len = static_cast<int32_t>(w >> 16);
    }
}
__device__ __forceinline__ void store_state1(const GridArgs& a, int64_t i, int32_t pos, int32_t len) {
    if (a.ep_len) {
        a.pos[i] = pos;
        a.ep_len[i] = len;
    } else {
        a.pos[i] = static_cast<int32_t>(static_cast<uint32_t>(pos) | (static_cast<uint32_t>(len) << 16));
    }
}
__device__ __forceinline__ void unpack4(const int4 w, int32_t (&pos)[4], int32_t (&len)[4]) {
    const uint32_t v[4] = {static_cast<uint32_t>(w.x), static_cast<uint32_t>(w.y), static_cast<uint32_t>(w.z),
                           static_cast<uint32_t>(w.w)};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        pos[j] = static_cast<int32_t>(v[j] & 0xffffu);
        len[j] = static_cast<int32_t>(v[j] >> 16);
    }
}
__device__ __forceinline__ int4 pack4(const int32_t (&pos)[4], const int32_t (&len)[4]) {
    uint32_t v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = static_cast<uint32_t>(pos[j]) | (static_cast<uint32_t>(len[j]) << 16);
    return make_int4(static_cast<int32_t>(v[0]), static_cast<int32_t>(v[1]), static_cast<int32_t>(v[2]),
                     static_cast<int32_t>(v[3]));
}

// Outcome index = min(number of thresholds <= r, idx_cap).  The host stores the thresholds at or beyond idx_cap as 2^32
// (unreachable for a 32-bit draw), which makes the min implicit; the count itself is computed without predicate
// registers: the high word of the 64-bit difference r - thr is 0xffffffff exactly when r < thr.  (The comparison form
// made ptxas spill predicates into a general register in the 16-envs-per-thread step kernel.)
__device__ __forceinline__ uint32_t outcome_index(const GridArgs& a, uint32_t r) {
    const uint32_t b0 = static_cast<uint32_t>((static_cast<uint64_t>(r) - a.thr0) >> 32);
    const uint32_t b1 = static_cast<uint32_t>((static_cast<uint64_t>(r) - a.thr1) >> 32);
    const uint32_t b2 = static_cast<uint32_t>((static_cast<uint64_t>(r) - a.thr2) >> 32);
    return 3u + b0 + b1 + b2;
}

// One transition of one env.  Returns bit0 terminated, bit1 truncated, bit2 rejected (state untouched).
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
    uint32_t slot = row;
    if constexpr (SLOTS == 4) {
        const uint32_t idx = outcome_index(a, r);
        slot = row * 4u + idx;
    }
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

// Branch-free transition of the four envs a thread owns (hot path of both kernels).  The host compiles rejected rows
// (no outcomes / probabilities that fail numpy's check) as inert entries -- next = own cell, reward 0, not done -- so
// applying one changes nothing but the step counter, which is masked with the row's flag bits.  Bit j of `bad` marks
// env j as rejected; byte j of `term` / `trunc` is env j's flag.  Actions must already be known to lie in 0..3 (the
// callers route quads with an out-of-range action through the generic `transition`).
struct Quad {
    int32_t pos[4], len[4];
    float ret[4], rew[4];
    uint32_t term, trunc, bad;
};

template <bool SMEM, int SLOTS>
__device__ __forceinline__ void transition4(const GridArgs& a, const uint2* __restrict__ tab, Quad& q,
                                            const int32_t (&act)[4], const uint32_t (&r)[4], const uint32_t (&off)[4]) {
    q.term = q.trunc = q.bad = 0u;
    const int32_t limit = a.limit > 0 ? a.limit : 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t slot = static_cast<uint32_t>(q.pos[j]) * 4u + static_cast<uint32_t>(act[j]);
        if constexpr (SLOTS == 4) {
            const uint32_t idx = outcome_index(a, r[j]);
            slot = slot * 4u + idx;
        }
        slot += off[j];  // table of this env's layout (all zero, and folded away, for single-layout batches)
        const uint2 o = SMEM ? tab[slot] : __ldg(tab + slot);
        const uint32_t ok = ((o.x >> 17) & 3u) == 0u ? 1u : 0u;  // 0 for a rejected (inert) row
        q.rew[j] = __uint_as_float(o.y);
        q.pos[j] = static_cast<int32_t>(o.x & 0xffffu);
        q.ret[j] += q.rew[j];
        q.len[j] += static_cast<int32_t>(ok);
        const uint32_t tr = (q.len[j] >= limit ? 1u : 0u) & ok;
        q.term |= ((o.x >> 16) & 1u) << (8 * j);
        q.trunc |= tr << (8 * j);
        q.bad |= (ok ^ 1u) << j;
    }
}

// Rare path: classify and record the rejected env-steps of a quad (a rejected env keeps its position, so the row can
// be rebuilt from the post-step state).  Arguments by value: keeps the quad in registers.
template <bool SMEM, int SLOTS>
__device__ __noinline__ uint4 record_bad4(const uint2* __restrict__ tab, uint32_t bad, uint32_t g0, int4 pos, int4 act,
                                          uint4 off = make_uint4(0u, 0u, 0u, 0u)) {
    uint32_t e_count = 0u, e_g = 0u, e_detail = 0u, e_kind = 0u;
    const int32_t ps[4] = {pos.x, pos.y, pos.z, pos.w};
    const int32_t ac[4] = {act.x, act.y, act.z, act.w};
    const uint32_t of[4] = {off.x, off.y, off.z, off.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (!((bad >> j) & 1u)) continue;
        ++e_count;
        e_g = g0 + j;
        const uint32_t row = static_cast<uint32_t>(ps[j]) * 4u + static_cast<uint32_t>(ac[j]);
        const uint2 o = SMEM ? tab[of[j] + row * SLOTS] : __ldg(tab + of[j] + row * SLOTS);
        e_detail = row;
        e_kind = ((o.x >> 16) & EF_GRID_NO_OUTCOMES) ? EF_STEP_NO_OUTCOMES : EF_STEP_BAD_PROBS;
    }
    return make_uint4(e_count, e_g, e_detail, e_kind);
}

// Auto-reset + statistics of the finished envs of a quad (fin = byte-wise finished bits).
// `lay` = layout of each env in a mixed-layout batch (start cell fetched here, on the rare finishing path, rather than
// up front for every env); null for single-layout batches, which reset to a.start_cell.
__device__ __forceinline__ void finish4(const GridArgs& a, Quad& q, uint32_t fin, EpisodeAcc& acc, const int32_t* lay = nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if ((fin >> (8 * j)) & 1u) {
            acc.finish(q.len[j], q.ret[j]);
            q.pos[j] = lay ? __ldg(a.start_cells + lay[j]) : a.start_cell;
            q.len[j] = 0;
            q.ret[j] = 0.0f;
        }
    }
}

// Inputs of one quad (four consecutive envs), loaded with 128-bit accesses.
struct QuadIn {
    int4 p4, l4, a4;
    float4 r4;
    int4 y4;  // layout index of each env (mixed-layout batches with an explicit index only)
};

// PACKED: the state word layout is known at compile time (ep_len == NULL).  FAST: the common call shape is known at
// compile time -- actions given, obs / reward / terminated / truncated all requested, no final_obs, auto-reset on -- so
// the launch-uniform null checks and their branches disappear from the instruction stream.
template <bool PACKED, bool FAST, bool HET = false, bool W8 = false>
__device__ __forceinline__ QuadIn load_quad(const GridArgs& a, int64_t base, bool philox_act) {
    QuadIn in;
    if constexpr (HET) in.y4 = a.layout ? __ldg(reinterpret_cast<const int4*>(a.layout + base)) : make_int4(0, 0, 0, 0);
    in.p4 = *reinterpret_cast<const int4*>(a.pos + base);
    if constexpr (PACKED) in.l4 = make_int4(0, 0, 0, 0);
    else in.l4 = *reinterpret_cast<const int4*>(a.ep_len + base);
    in.r4 = *reinterpret_cast<const float4*>(a.ep_ret + base);
    if constexpr (W8) {  // four action bytes in one 32-bit load
        const uint32_t w = philox_act ? 0u : __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(a.action) + base));
        in.a4 = make_int4(static_cast<int32_t>(w & 0xffu), static_cast<int32_t>((w >> 8) & 0xffu),
                          static_cast<int32_t>((w >> 16) & 0xffu), static_cast<int32_t>(w >> 24));
    } else {
        in.a4 = (!FAST && philox_act) ? make_int4(0, 0, 0, 0) : __ldg(reinterpret_cast<const int4*>(a.action + base));
    }
    return in;
}

// One step of a full quad, in registers: draws, branch-free transitions, rejected-step bookkeeping, auto-reset.  Leaves
// the post-step quad in `q` (position / counter / return AFTER auto-reset; reward and flags of the step).  `base` is only
// used for the optional pre-reset observation (final_obs), which is stored here because the reset overwrites it.
template <bool SMEM, int SLOTS, bool PACKED, bool FAST, bool HET = false, bool W8 = false>
__device__ __forceinline__ void step_quad_compute(const GridArgs& a, const uint2* __restrict__ tab, int64_t base, uint32_t g0,
                                                  uint64_t t, const QuadIn& in, bool philox_act, EpisodeAcc& acc,
                                                  ErrorAcc& errs, Quad& q) {
    int32_t act[4] = {in.a4.x, in.a4.y, in.a4.z, in.a4.w};
    uint32_t r[4] = {0u, 0u, 0u, 0u};
    if constexpr (PACKED) {
        unpack4(in.p4, q.pos, q.len);
    } else {
        q.pos[0] = in.p4.x; q.pos[1] = in.p4.y; q.pos[2] = in.p4.z; q.pos[3] = in.p4.w;
        q.len[0] = in.l4.x; q.len[1] = in.l4.y; q.len[2] = in.l4.z; q.len[3] = in.l4.w;
    }
    q.ret[0] = in.r4.x; q.ret[1] = in.r4.y; q.ret[2] = in.r4.z; q.ret[3] = in.r4.w;
    if (!FAST && philox_act) {
        uint32_t w[4];
        group_words4_rk(a.rk, a.seed, g0, t, kStreamPolicy, w);
#pragma unroll
        for (int j = 0; j < 4; ++j) act[j] = static_cast<int32_t>(w[j] >> 30);
    }
    if constexpr (SLOTS == 4) group_words4_rk(a.rk, a.seed, g0, t, kStreamTransition, r);
    // HET: the batch mixes layouts -- every env walks its own table (offset into the staged tables) and resets to its own
    // start cell.  Otherwise both arrays are compile-time constants and fold away.
    uint32_t off[4] = {0u, 0u, 0u, 0u};
    int32_t lay[4] = {in.y4.x, in.y4.y, in.y4.z, in.y4.w};
    if constexpr (HET) {
        if (!a.layout) {  // layout = global env index mod n_layouts: one division per quad, the rest by wrap-around
            const uint32_t nl = static_cast<uint32_t>(a.n_layouts);
            const uint32_t l0 = g0 % nl;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t x = l0 + j;
                x -= x >= nl ? nl : 0u;
                x -= x >= nl ? nl : 0u;
                x -= x >= nl ? nl : 0u;
                lay[j] = static_cast<int32_t>(x);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) off[j] = static_cast<uint32_t>(lay[j]) * static_cast<uint32_t>(a.layout_len);
    }
    // An action outside 0..3 (the reference raises ValueError in Action(action)) must never index the table: a wild
    // value would read far outside it.  Such quads skip the branch-free walk and go one env at a time through the
    // checking path below.
    const bool wild = ((act[0] | act[1] | act[2] | act[3]) & ~3) != 0;
    if (!wild) transition4<SMEM, SLOTS>(a, tab, q, act, r, off);
    uint32_t rejected = 0u;
    if (wild) {
        q.term = q.trunc = q.bad = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t f = transition<SMEM, SLOTS>(a, tab + off[j], g0 + j, q.pos[j], q.len[j], q.ret[j], act[j], r[j],
                                                       q.rew[j], errs);
            q.term |= (f & 1u) << (8 * j);
            q.trunc |= ((f >> 1) & 1u) << (8 * j);
            rejected += (f >> 2) & 1u;
        }
    } else if (q.bad) {
        const uint4 e = record_bad4<SMEM, SLOTS>(tab, q.bad, g0, make_int4(q.pos[0], q.pos[1], q.pos[2], q.pos[3]),
                                                 make_int4(act[0], act[1], act[2], act[3]),
                                                 make_uint4(off[0], off[1], off[2], off[3]));
        errs.count += e.x;
        errs.g = e.y;
        errs.detail = e.z;
        errs.kind = e.w;
        rejected = e.x;
    }
    acc.steps += 4u - rejected;
    if (!FAST && a.final_obs) store_obs4<W8>(a, a.final_obs, base, q.pos);  // pre-reset observation, only on request
    const uint32_t fin = (FAST || a.autoreset) ? (q.term | q.trunc) : 0u;
    if (fin) finish4(a, q, fin, acc, HET ? lay : nullptr);
}

// 128-bit stores of a stepped quad straight to global memory.
template <bool PACKED, bool FAST, bool W8 = false>
__device__ __forceinline__ void store_quad_global(const GridArgs& a, int64_t base, const Quad& q) {
    if constexpr (PACKED) {
        *reinterpret_cast<int4*>(a.pos + base) = pack4(q.pos, q.len);
    } else {
        *reinterpret_cast<int4*>(a.pos + base) = make_int4(q.pos[0], q.pos[1], q.pos[2], q.pos[3]);
        *reinterpret_cast<int4*>(a.ep_len + base) = make_int4(q.len[0], q.len[1], q.len[2], q.len[3]);
    }
    *reinterpret_cast<float4*>(a.ep_ret + base) = make_float4(q.ret[0], q.ret[1], q.ret[2], q.ret[3]);
    if (FAST || a.obs) store_obs4<W8>(a, a.obs, base, q.pos);  // observation after auto-reset = (row, col) of the stored position
    if (FAST || a.reward) *reinterpret_cast<float4*>(a.reward + base) = make_float4(q.rew[0], q.rew[1], q.rew[2], q.rew[3]);
    if (FAST || a.terminated) *reinterpret_cast<uint32_t*>(a.terminated + base) = q.term;
    if (FAST || a.truncated) *reinterpret_cast<uint32_t*>(a.truncated + base) = q.trunc;
}

// base3 - [r < t0] - [r < t1] - [r < t2]  ==  base3 - 3 + #(thresholds <= r): three borrow-chained subtractions
// (IADD3 with carry-out + IADD3.X) instead of three 64-bit differences.
__device__ __forceinline__ uint32_t slot_minus_borrows(uint32_t base3, uint32_t r, uint32_t t0, uint32_t t1, uint32_t t2) {
    asm("{\n\t"
        ".reg .u32 t;\n\t"
        "sub.cc.u32 t, %1, %2;\n\t"
        "subc.u32 %0, %0, 0;\n\t"
        "sub.cc.u32 t, %1, %3;\n\t"
        "subc.u32 %0, %0, 0;\n\t"
        "sub.cc.u32 t, %1, %4;\n\t"
        "subc.u32 %0, %0, 0;\n\t"
        "}"
        : "+r"(base3)
        : "r"(r), "r"(t0), "r"(t1), "r"(t2));
    return base3;
}

// A stepped quad in the packed state layout: w = destination cell | episode counter << 16 (after auto-reset).
struct PackedQuad {
    uint32_t w[4];
    float ret[4], rew[4];
    uint32_t term, trunc;   // byte j = flag of env j
};

// One env of a packed quad through the checking path (an action outside 0..3, or a row the reference raises on), by
// value so that the caller's quad stays in registers.  Returns {packed word, return bits, reward bits, f}: f bit0
// terminated, bit1 truncated, bit2 rejected (state untouched), bits 8-15 the ef_step_error kind, bits 16-.. the detail.
template <bool SMEM, int SLOTS>
__device__ __noinline__ uint4 packed_slow_env(const uint2* __restrict__ tab, uint32_t thr0, uint32_t thr1, uint32_t thr2,
                                              uint32_t idx_cap, uint32_t lim_hi, int32_t act, uint32_t r, uint32_t w, float ret) {
    const uint32_t pos = w & 0xffffu;
    if (static_cast<uint32_t>(act) > 3u)
        return make_uint4(w, __float_as_uint(ret), 0u, 4u | (EF_STEP_BAD_ACTION << 8) | (static_cast<uint32_t>(act) << 16));
    const uint32_t row = pos * 4u + static_cast<uint32_t>(act);
    uint32_t slot = row * SLOTS;
    if (SLOTS == 4) slot += min(slot_minus_borrows(3u, r, thr0, thr1, thr2), idx_cap);
    const uint2 o = SMEM ? tab[slot] : __ldg(tab + slot);
    const uint32_t fl = o.x >> 16;
    if (fl & (EF_GRID_NO_OUTCOMES | EF_GRID_BAD_PROBS))
        return make_uint4(w, __float_as_uint(ret), 0u,
                          4u | (((fl & EF_GRID_NO_OUTCOMES) ? EF_STEP_NO_OUTCOMES : EF_STEP_BAD_PROBS) << 8) | (row << 16));
    w = __byte_perm(o.x, w, 0x7610) + 0x10000u;
    return make_uint4(w, __float_as_uint(ret + __uint_as_float(o.y)), o.y, (fl & EF_GRID_DONE) | (w >= lim_hi ? 2u : 0u));
}

// One step of a full quad in the PACKED state layout with auto-reset on (what the packed layout implies), single-layout
// batch.  Same Philox words, same table walk and same results as step_quad_compute, fewer instructions: the episode
// counter is stepped inside the packed word (one PRMT + one add per env), the four flag bytes are gathered with three
// PRMTs, the outcome index is a borrow chain, and rejected rows / out-of-range actions (rare) leave the straight-line code.
template <bool SMEM, int SLOTS, bool FAST, bool W8>
__device__ __forceinline__ void step_quad_packed(const GridArgs& a, const uint2* __restrict__ tab, int64_t base, uint32_t g0,
                                                 uint64_t t, const QuadIn& in, bool philox_act, EpisodeAcc& acc,
                                                 ErrorAcc& errs, PackedQuad& q, const uint4* spec = nullptr) {
    int32_t act[4] = {in.a4.x, in.a4.y, in.a4.z, in.a4.w};
    q.w[0] = static_cast<uint32_t>(in.p4.x); q.w[1] = static_cast<uint32_t>(in.p4.y);
    q.w[2] = static_cast<uint32_t>(in.p4.z); q.w[3] = static_cast<uint32_t>(in.p4.w);
    q.ret[0] = in.r4.x; q.ret[1] = in.r4.y; q.ret[2] = in.r4.z; q.ret[3] = in.r4.w;
    if (!FAST && philox_act) {
        uint32_t wp[4];
        group_words4_rk(a.rk, a.seed, g0, t, kStreamPolicy, wp);
#pragma unroll
        for (int j = 0; j < 4; ++j) act[j] = static_cast<int32_t>(wp[j] >> 30);
    }
    uint32_t r[4] = {0u, 0u, 0u, 0u};
    if constexpr (SLOTS == 4) {
        if (spec != nullptr) {   // the quad's outcome draws of step t were computed ahead of griddepcontrol.wait
            const uint4 d = *spec;
            r[0] = d.x; r[1] = d.y; r[2] = d.z; r[3] = d.w;
        } else {
            group_words4_rk(a.rk, a.seed, g0, t, kStreamTransition, r);
        }
    }
    // thresholds at or beyond idx_cap are 2^32 (unreachable): clipped to 2^32 - 1 here, the index cap does the rest
    const uint32_t thr0 = a.thr0 > 0xffffffffull ? 0xffffffffu : static_cast<uint32_t>(a.thr0);
    const uint32_t thr1 = a.thr1 > 0xffffffffull ? 0xffffffffu : static_cast<uint32_t>(a.thr1);
    const uint32_t thr2 = a.thr2 > 0xffffffffull ? 0xffffffffu : static_cast<uint32_t>(a.thr2);
    const uint32_t lim_hi = static_cast<uint32_t>(a.limit) << 16;   // packed layout: 0 < limit <= 65535
    const bool wild = ((act[0] | act[1] | act[2] | act[3]) & ~3) != 0;
    uint32_t gathered = 0u;
    uint2 o[4];
    if (!wild) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t slot = ((q.w[j] & 0xffffu) * 4u + static_cast<uint32_t>(act[j])) * SLOTS;
            if constexpr (SLOTS == 4) {
                if (a.idx_cap == 3) slot = slot_minus_borrows(slot + 3u, r[j], thr0, thr1, thr2);   // launch-uniform
                else slot += min(slot_minus_borrows(3u, r[j], thr0, thr1, thr2), static_cast<uint32_t>(a.idx_cap));
            }
            o[j] = SMEM ? tab[slot] : __ldg(tab + slot);
        }
        // byte j of `gathered` = flag byte of env j's outcome (bit0 done, bit1 no outcomes, bit2 bad probabilities)
        gathered = __byte_perm(__byte_perm(o[0].x, o[1].x, 0x0062), __byte_perm(o[2].x, o[3].x, 0x0062), 0x5410);
    }
    q.term = q.trunc = 0u;
    if (!wild) {
        // Rows the reference raises on are compiled INERT (next = own cell, reward 0, not done): applying one changes
        // nothing but the episode counter, which the fix-up below takes back -- so the straight-line code never branches
        // on them (they are ~0.1 % of the env-steps of BASELINE config 2, i.e. some lane of every eighth warp).
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            q.w[j] = __byte_perm(o[j].x, q.w[j], 0x7610) + 0x10000u;   // destination cell | (counter + 1) << 16
            q.rew[j] = __uint_as_float(o[j].y);
            q.ret[j] += q.rew[j];
        }
        uint32_t rejected = 0u;
        const uint32_t badmask = gathered & 0x06060606u;
        if (badmask) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t fl = (badmask >> (8 * j)) & 0xffu;
                if (fl) {
                    q.w[j] -= 0x10000u;
                    ++rejected;
                    errs.record(g0 + j, (q.w[j] & 0xffffu) * 4u + static_cast<uint32_t>(act[j]),
                                (fl & EF_GRID_NO_OUTCOMES) ? EF_STEP_NO_OUTCOMES : EF_STEP_BAD_PROBS);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) q.trunc |= (q.w[j] >= lim_hi ? 1u : 0u) << (8 * j);
        q.term = gathered & 0x01010101u;
        acc.steps += 4u - rejected;
    } else {   // an action outside 0..3 somewhere in the quad: one env at a time through the checking path
        uint32_t rejected = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 x = packed_slow_env<SMEM, SLOTS>(tab, thr0, thr1, thr2, static_cast<uint32_t>(a.idx_cap), lim_hi, act[j],
                                                         r[j], q.w[j], q.ret[j]);
            q.w[j] = x.x;
            q.ret[j] = __uint_as_float(x.y);
            q.rew[j] = __uint_as_float(x.z);
            q.term |= (x.w & 1u) << (8 * j);
            q.trunc |= ((x.w >> 1) & 1u) << (8 * j);
            if (x.w & 4u) {
                ++rejected;
                errs.record(g0 + j, x.w >> 16, (x.w >> 8) & 0xffu);
            }
        }
        acc.steps += 4u - rejected;
    }
    if (!FAST && a.final_obs) {  // pre-reset observation, only on request
        const int32_t fp[4] = {static_cast<int32_t>(q.w[0] & 0xffffu), static_cast<int32_t>(q.w[1] & 0xffffu),
                               static_cast<int32_t>(q.w[2] & 0xffffu), static_cast<int32_t>(q.w[3] & 0xffffu)};
        store_obs4<W8>(a, a.final_obs, base, fp);
    }
    const uint32_t fin = q.term | q.trunc;   // auto-reset is on whenever the state is packed
    if (fin) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if ((fin >> (8 * j)) & 1u) {
                acc.finish(static_cast<int>(q.w[j] >> 16), q.ret[j]);
                q.w[j] = static_cast<uint32_t>(a.start_cell);
                q.ret[j] = 0.0f;
            }
        }
    }
}

template <bool FAST, bool W8>
__device__ __forceinline__ void store_quad_packed(const GridArgs& a, int64_t base, const PackedQuad& q) {
    *reinterpret_cast<uint4*>(a.pos + base) = make_uint4(q.w[0], q.w[1], q.w[2], q.w[3]);
    *reinterpret_cast<float4*>(a.ep_ret + base) = make_float4(q.ret[0], q.ret[1], q.ret[2], q.ret[3]);
    if (FAST || a.obs) {  // observation after auto-reset = (row, col) of the stored position
        const int32_t p[4] = {static_cast<int32_t>(q.w[0] & 0xffffu), static_cast<int32_t>(q.w[1] & 0xffffu),
                              static_cast<int32_t>(q.w[2] & 0xffffu), static_cast<int32_t>(q.w[3] & 0xffffu)};
        store_obs4<W8>(a, a.obs, base, p);
    }
    if (FAST || a.reward) *reinterpret_cast<float4*>(a.reward + base) = make_float4(q.rew[0], q.rew[1], q.rew[2], q.rew[3]);
    if (FAST || a.terminated) *reinterpret_cast<uint32_t*>(a.terminated + base) = q.term;
    if (FAST || a.truncated) *reinterpret_cast<uint32_t*>(a.truncated + base) = q.trunc;
}

template <bool SMEM, int SLOTS, bool PACKED, bool FAST, bool HET = false, bool W8 = false>
__device__ __forceinline__ void step_quad(const GridArgs& a, const uint2* __restrict__ tab, int64_t base, uint32_t g0,
                                          uint64_t t, const QuadIn& in, bool philox_act, EpisodeAcc& acc,
                                          ErrorAcc& errs, const uint4* spec = nullptr) {
    if constexpr (EF_STEP_LEAN && PACKED && !HET) {
        PackedQuad q;
        step_quad_packed<SMEM, SLOTS, FAST, W8>(a, tab, base, g0, t, in, philox_act, acc, errs, q, spec);
        store_quad_packed<FAST, W8>(a, base, q);
    } else {
        Quad q;
        step_quad_compute<SMEM, SLOTS, PACKED, FAST, HET, W8>(a, tab, base, g0, t, in, philox_act, acc, errs, q);
        store_quad_global<PACKED, FAST, W8>(a, base, q);
    }
}

// Ragged tail of the vector kernels: the at most three envs past the last full quad, stepped by the last CTA through the
// generic one-env code.
template <bool SMEM, int SLOTS, bool HET, bool W8>
__device__ __forceinline__ void step_tail(const GridArgs& a, const uint2* __restrict__ tab, uint64_t t, bool philox_act,
                                          EpisodeAcc& acc, ErrorAcc& errs) {
    const int64_t quads = a.n >> 2;
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x < (a.n & 3)) {
        const int64_t i = (quads << 2) + threadIdx.x;
        const uint32_t g = static_cast<uint32_t>(a.env_offset + i);
        int32_t pos, len;
        load_state1(a, i, pos, len);
        float ret = a.ep_ret[i], rew;
        const int32_t act = philox_act ? static_cast<int32_t>(group_word1(a.seed, g, t, kStreamPolicy) >> 30)
                                       : load_action1<W8>(a, i);
        uint32_t r = 0u;
        if constexpr (SLOTS == 4) r = group_word1(a.seed, g, t, kStreamTransition);
        const uint2* __restrict__ tab_i = tab;
        int32_t start = a.start_cell;
        if constexpr (HET) {
            const int32_t lay = a.layout ? __ldg(a.layout + i) : static_cast<int32_t>(g % static_cast<uint32_t>(a.n_layouts));
            tab_i = tab + static_cast<int64_t>(lay) * a.layout_len;
            start = __ldg(a.start_cells + lay);
        }
        const uint32_t f = transition<SMEM, SLOTS>(a, tab_i, g, pos, len, ret, act, r, rew, errs);
        if (!(f & 4u)) acc.steps += 1u;
        if (a.final_obs) {
            const int2 fo = cell_to_obs(a, pos);
            store_obs1<W8>(a.final_obs, i, fo);
        }
        if (a.autoreset && (f & 3u)) {
            acc.finish(len, ret);
            pos = start;
            len = 0;
            ret = 0.0f;
        }
        store_state1(a, i, pos, len);
        a.ep_ret[i] = ret;
        if (a.obs) {
            const int2 ob = cell_to_obs(a, pos);
            store_obs1<W8>(a.obs, i, ob);
        }
        if (a.reward) a.reward[i] = rew;
        if (a.terminated) a.terminated[i] = static_cast<uint8_t>(f & 1u);
        if (a.truncated) a.truncated[i] = static_cast<uint8_t>((f >> 1) & 1u);
    }
}

#ifndef EF_STEP_MINBLOCKS
#define EF_STEP_MINBLOCKS 2  // resident CTAs per SM the step kernel is compiled for (register cap 65536 / (256 * this))
#endif
// Large batches, dynamic CTA scheduling.  Measured on B200 with the packed state layout (34 B moved per env-step), us per
// launch at 4M / 16M envs: Q=2 x 3 CTAs/SM 28.9 / 105.1 (adopted), Q=4 x 2 32.8 / 110.4, Q=3 x 2 32.5 / 115.3,
// Q=1 x 4 31.4 / 116.6, Q=2 x 4 36.5 / 125.9, Q=4 x 3 40.5 / 135.2 (the last two spill registers).
#ifndef EF_STEP_BIG_Q
#define EF_STEP_BIG_Q 2
#endif
#ifndef EF_STEP_BIG_MINBLOCKS
#define EF_STEP_BIG_MINBLOCKS 3
#endif
#ifndef EF_STEP_BIG_GRID_PER_SM
#define EF_STEP_BIG_GRID_PER_SM 64
#endif
#ifndef EF_STEP_GRID_PER_SM
#define EF_STEP_GRID_PER_SM EF_STEP_MINBLOCKS  // grid cap in CTAs per SM
#endif
#ifndef EF_STEP_PASSES
#define EF_STEP_PASSES 1     // load/compute passes of EF_STEP_Q quads a thread may make in the one-wave ("small") geometry
#endif
#ifndef EF_STEP_Q
#define EF_STEP_Q 4          // quads (of four envs) whose loads a thread keeps in flight at once
#endif
// ---- step -------------------------------------------------------------------------------------------------------------
// Vector kernel (all pointers 16-byte aligned, Philox or no outcome draws).  The batch is cut into one contiguous chunk of
// quads per CTA (grid = a few CTAs per SM, all resident), and a thread first ISSUES the 128-bit loads of all its quads
// (up to EF_STEP_Q per pass), then steps them one after the other: a launch is one round of DRAM latency with every
// load in flight at once, instead of a grid-stride loop that pays the latency once per iteration.
template <bool SMEM, int SLOTS, int Q, int MINBLOCKS, bool PACKED, bool FAST, bool HET = false, bool W8 = false, bool COLD = false>
__global__ void __launch_bounds__(kThreads, MINBLOCKS) gridworld_step_kernel(const GridArgs a) {
    __shared__ uint64_t table_bar;
    bool table_pending;
    EF_TRACE_DECL
    EF_TRACE_POINT(0);   // kernel entry
    const bool philox_act = FAST ? false : a.action == nullptr;
    const int64_t quads = a.n >> 2;  // full quads; the (n & 3)-env tail goes through the generic code below
    const int64_t per_cta = a.quads_per_cta;  // ceil(quads / gridDim.x), divided on the host
    const int64_t q_begin = static_cast<int64_t>(blockIdx.x) * per_cta;
    const int64_t q_end = min(q_begin + per_cta, quads);
#if EF_STEP_L2_PREFETCH
    // This CTA is resident up to a few microseconds before the previous launch releases it: ask L2 for the chunk's input
    // lines now (a hint -- whatever the previous launch still writes to them lands in the same L2 lines), so that the
    // loads below find them there instead of queueing in DRAM behind every other CTA's at the moment of release.
    if (q_begin < q_end) {
        const uint32_t bytes = static_cast<uint32_t>(q_end - q_begin) * 16u;   // 4 B per env of a quad
        const int64_t e0 = q_begin * 4;
#if EF_STEP_L2_PREFETCH == 1
        if (threadIdx.x == 0) l2_prefetch_bulk(a.pos + e0, bytes);
        if (threadIdx.x == 32) l2_prefetch_bulk(a.ep_ret + e0, bytes);
        if (!PACKED && threadIdx.x == 64) l2_prefetch_bulk(a.ep_len + e0, bytes);
        if (!W8 && !philox_act && threadIdx.x == 96) l2_prefetch_bulk(a.action + e0, bytes);
#else
        for (uint32_t off = threadIdx.x * 128u; off < bytes; off += kThreads * 128u) {
            l2_prefetch_line(reinterpret_cast<const char*>(a.pos + e0) + off);
            l2_prefetch_line(reinterpret_cast<const char*>(a.ep_ret + e0) + off);
            if (!PACKED) l2_prefetch_line(reinterpret_cast<const char*>(a.ep_len + e0) + off);
            if (!W8 && !philox_act) l2_prefetch_line(reinterpret_cast<const char*>(a.action + e0) + off);
        }
#endif
        if (W8 && !philox_act)
            for (uint32_t off = threadIdx.x * 128u; off < bytes / 4u; off += kThreads * 128u)
                l2_prefetch_line(reinterpret_cast<const char*>(a.action) + e0 + off);
    }
#endif
#if EF_STEP_PRE_CLOCK
    // The clock word is read right after the wait and everything waits for it.  Warm it now: with the env set's data in L2
    // (the usual loop) a load whose result is dropped -- measured 0.2 us per launch better than a prefetch hint there;
    // in the COLD instantiation (EF_STEP_COLD_INPUTS: other env sets' traffic has pushed this set's lines out of L2 since
    // its last step) a line prefetch, 0.3 us better there.  Compile-time, not a branch: the mere presence of the load in
    // the instruction stream costs the cold chain those 0.3 us (profiles/r02bd_cold_hint_matrix.txt, r02be_*).
    if (threadIdx.x == 0 && a.clock != nullptr) {
        if constexpr (COLD) {
            l2_prefetch_line(a.clock);
        } else {
            unsigned long long sink;
            asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(sink) : "l"(a.clock) : "memory");
        }
    }
#endif
    const uint2* __restrict__ tab = stage_table_begin<SMEM>(a, &table_bar, table_pending);  // independent of the previous launch
    pdl_launch_dependents();
    EF_TRACE_POINT(12);  // prefetches issued, table copy under way
    pdl_wait();                                            // previous launch complete, its writes visible
    EF_TRACE_POINT(1);   // released by the previous launch
    EpisodeAcc acc;
    ErrorAcc errs;
    QuadIn in[Q];
    auto issue_loads = [&](int64_t chunk) {
        // 32-bit bookkeeping inside the chunk (a CTA's share is far below 2^31 quads); 64-bit only for the addresses
        const uint32_t left = static_cast<uint32_t>(min(q_end - chunk, static_cast<int64_t>(Q) * kThreads));
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            const uint32_t local = static_cast<uint32_t>(j) * kThreads + threadIdx.x;
            if (local < left) in[j] = load_quad<PACKED, FAST, HET, W8>(a, (chunk + local) * 4, philox_act);
        }
    };
#if EF_STEP_POST_CLOCK == 1
    // ONE thread of the CTA reads the clock word (2368 warps asking L2 for the same sector queue up behind each other:
    // about 0.7 us at the median, and the word is under atomic fire in a same-set chain), ahead of the input loads; the
    // CTA picks it up behind a barrier that falls into the loads' latency.
    __shared__ unsigned long long s_clk1;
    if (threadIdx.x == 0) s_clk1 = clock_fetch(a.clock);
    if (q_begin < q_end) issue_loads(q_begin);
    __syncthreads();
    const unsigned long long clk = s_clk1;
#else
    const unsigned long long clk = clock_fetch(a.clock);   // in flight ahead of the input loads, consumed by Philox
    if (q_begin < q_end) issue_loads(q_begin);
#endif
    const uint64_t t = a.t + clock_steps(clk);
    EF_TRACE_POINT(2);
    for (int64_t chunk = q_begin; chunk < q_end; chunk += static_cast<int64_t>(Q) * kThreads) {
        if (chunk != q_begin) issue_loads(chunk);
        const uint32_t left = static_cast<uint32_t>(min(q_end - chunk, static_cast<int64_t>(Q) * kThreads));
        stage_table_wait(&table_bar, table_pending);  // the bulk copy had the whole load phase to land
        EF_TRACE_POINT(3);   // loads issued, table staged
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            const uint32_t local = static_cast<uint32_t>(j) * kThreads + threadIdx.x;
            if (local < left)
                step_quad<SMEM, SLOTS, PACKED, FAST, HET, W8>(a, tab, (chunk + local) * 4,
                                                          static_cast<uint32_t>(a.env_offset + (chunk + local) * 4), t, in[j],
                                                          philox_act, acc, errs);
            if (j == 0) EF_TRACE_POINT(4);   // first quad stepped (its loads had to land)
        }
    }
    stage_table_wait(&table_bar, table_pending);  // CTAs that own no full quad (tiny batches) still need the table below
    step_tail<SMEM, SLOTS, HET, W8>(a, tab, t, philox_act, acc, errs);
    EF_TRACE_POINT(13);  // before the CTA epilogue
    cta_epilogue(acc, errs, a.stats, a.err, a.clock, 0ull, 1ull);
    EF_TRACE_POINT(14);  // done
    EF_TRACE_FLUSH(t);
}

// Scalar kernel: one env per thread, any alignment, optional injected outcome draws (parity tests, unaligned views).
template <bool SMEM, int SLOTS, bool W8 = false>
__global__ void __launch_bounds__(kThreads) gridworld_step_scalar_kernel(const GridArgs a) {
    const uint2* __restrict__ tab = stage_table<SMEM>(a);
    pdl_launch_dependents();
    pdl_wait();
    EpisodeAcc acc;
    ErrorAcc errs;
    unsigned long long ticket;
    const uint64_t t = a.t + clock_begin(a.clock, ticket);
    const bool philox_act = a.action == nullptr;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint32_t g = static_cast<uint32_t>(a.env_offset + i);
        int32_t pos, len;
        load_state1(a, i, pos, len);
        float ret = a.ep_ret[i], rew;
        const int32_t act = philox_act ? static_cast<int32_t>(group_word1(a.seed, g, t, kStreamPolicy) >> 30)
                                       : load_action1<W8>(a, i);
        uint32_t r = 0u;
        if constexpr (SLOTS == 4) r = a.draw ? __ldg(a.draw + i) : group_word1(a.seed, g, t, kStreamTransition);
        const uint2* __restrict__ tab_i = tab;
        int32_t start = a.start_cell;
        if (a.start_cells) {  // launch-uniform: several layouts in the batch -- this env's own table and start cell
            const int32_t lay = a.layout ? __ldg(a.layout + i) : static_cast<int32_t>(g % static_cast<uint32_t>(a.n_layouts));
            tab_i = tab + static_cast<int64_t>(lay) * a.layout_len;
            start = __ldg(a.start_cells + lay);
        }
        const uint32_t f = transition<SMEM, SLOTS>(a, tab_i, g, pos, len, ret, act, r, rew, errs);
        if (!(f & 4u)) acc.steps += 1u;
        if (a.final_obs) {
            const int2 fo = cell_to_obs(a, pos);
            store_obs1<W8>(a.final_obs, i, fo);
        }
        if (a.autoreset && (f & 3u)) {
            acc.finish(len, ret);
            pos = start;
            len = 0;
            ret = 0.0f;
        }
        store_state1(a, i, pos, len);
        a.ep_ret[i] = ret;
        if (a.obs) {
            const int2 ob = cell_to_obs(a, pos);
            store_obs1<W8>(a.obs, i, ob);
        }
        if (a.reward) a.reward[i] = rew;
        if (a.terminated) a.terminated[i] = static_cast<uint8_t>(f & 1u);
        if (a.truncated) a.truncated[i] = static_cast<uint8_t>((f >> 1) & 1u);
    }
    cta_epilogue(acc, errs, a.stats, a.err, a.clock, ticket, 1ull);
}

// ---- stream-exact step (numpy PCG64 per env) ------------------------------------------------------------------------
// `GridWorld(seed=s)` draws one Generator.random() double per successful step from Generator(PCG64(SeedSequence(s)))
// (grid_world.py:69, :515).  Here every env carries that generator's 128-bit state and increment; a step advances it
// (state = state * MULT + inc, XSL-RR output) and compares m = raw >> 11 against 53-bit integer thresholds, which is
// numpy's searchsorted(cdf, m * 2^-53, 'right').  Rejected steps do not consume a draw, like the reference (it raises
// before calling choice / inside choice before drawing).
struct Pcg64Args {
    unsigned long long thr[3];
    int32_t idx_cap;
    ulonglong2* rng_state;      // x = high word, y = low word
    const ulonglong2* rng_inc;
};

template <bool SMEM, int SLOTS>
__global__ void __launch_bounds__(kThreads) gridworld_step_pcg64_kernel(const GridArgs a, const Pcg64Args r) {
    const uint2* __restrict__ tab = stage_table<SMEM>(a);
    pdl_launch_dependents();
    pdl_wait();
    EpisodeAcc acc;
    ErrorAcc errs;
    constexpr unsigned long long kMultHi = 0x2360ED051FC65DA4ull, kMultLo = 0x4385DF649FCCF645ull;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint32_t g = static_cast<uint32_t>(a.env_offset + i);
        int32_t pos, len;
        load_state1(a, i, pos, len);
        float ret = a.ep_ret[i], rew = 0.0f;
        const int32_t act = __ldg(a.action + i);
        uint32_t f = 0u;
        if (static_cast<uint32_t>(act) > 3u) {
            errs.record(g, static_cast<uint32_t>(act), EF_STEP_BAD_ACTION);
            f = 4u;
        } else {
            const uint32_t row = static_cast<uint32_t>(pos) * 4u + static_cast<uint32_t>(act);
            const uint2 probe = SMEM ? tab[row * SLOTS] : __ldg(tab + row * SLOTS);  // every slot of a row carries its flags
            if ((probe.x >> 16) & (EF_GRID_NO_OUTCOMES | EF_GRID_BAD_PROBS)) {
                errs.record(g, row, ((probe.x >> 16) & EF_GRID_NO_OUTCOMES) ? EF_STEP_NO_OUTCOMES : EF_STEP_BAD_PROBS);
                f = 4u;
            } else {
                ulonglong2 st = r.rng_state[i];
                const ulonglong2 inc = r.rng_inc[i];
                // 128-bit LCG step
                unsigned long long lo = st.y * kMultLo;
                unsigned long long hi = __umul64hi(st.y, kMultLo) + st.x * kMultLo + st.y * kMultHi;
                lo += inc.y;
                hi += inc.x + (lo < inc.y ? 1ull : 0ull);
                r.rng_state[i] = make_ulonglong2(hi, lo);
                const unsigned long long x = hi ^ lo;
                const unsigned rot = static_cast<unsigned>(hi >> 58);
                const unsigned long long raw = (x >> rot) | (x << ((64u - rot) & 63u));
                const unsigned long long m = raw >> 11;
                uint32_t idx = 0u;
                if constexpr (SLOTS == 4) {
                    idx = (m >= r.thr[0]) + (m >= r.thr[1]) + (m >= r.thr[2]);
                    idx = min(idx, static_cast<uint32_t>(r.idx_cap));
                }
                const uint2 o = SMEM ? tab[row * SLOTS + idx] : __ldg(tab + row * SLOTS + idx);
                pos = static_cast<int32_t>(o.x & 0xffffu);
                rew = __uint_as_float(o.y);
                ret += rew;
                len += 1;
                f = ((o.x >> 16) & EF_GRID_DONE) | ((a.limit > 0 && len >= a.limit) ? 2u : 0u);
                acc.steps += 1u;
            }
        }
        if (a.final_obs) {
            const int2 fo = cell_to_obs(a, pos);
            a.final_obs[2 * i] = fo.x;
            a.final_obs[2 * i + 1] = fo.y;
        }
        if (a.autoreset && (f & 3u)) {
            acc.finish(len, ret);
            pos = a.start_cell;
            len = 0;
            ret = 0.0f;
        }
        store_state1(a, i, pos, len);
        a.ep_ret[i] = ret;
        if (a.obs) {
            const int2 ob = cell_to_obs(a, pos);
            a.obs[2 * i] = ob.x;
            a.obs[2 * i + 1] = ob.y;
        }
        if (a.reward) a.reward[i] = rew;
        if (a.terminated) a.terminated[i] = static_cast<uint8_t>(f & 1u);
        if (a.truncated) a.truncated[i] = static_cast<uint8_t>((f >> 1) & 1u);
    }
    cta_epilogue(acc, errs, a.stats, a.err, nullptr, 0ull, 0ull);
}

// ---- fused K-step rollout -------------------------------------------------------------------------------------------
// A thread keeps FOUR consecutive envs in registers for all K steps (one Philox call per step yields their four outcome
// draws, one more their four random actions; the four table walks are independent, which gives the scheduler ILP).
// A warp votes (ballot) on "any of my lanes finished an episode this step"; the reset / statistics path is skipped by
// warps where nobody did.  State is read and written once per launch (24 B / K per env-step).
// HET: the batch mixes layouts (ef_gridworld_rollout_layouts) -- each env walks its own table at a fixed offset into the
// staged tables and resets to its own start cell; both are fetched once, before the K steps.
template <bool SMEM, int SLOTS, bool EXT_ACTIONS, bool EMIT, bool HET = false>
__global__ void __launch_bounds__(kThreads) gridworld_rollout_kernel(const GridArgs a) {
    const uint2* __restrict__ tab = stage_table<SMEM>(a);
    pdl_launch_dependents();
    pdl_wait();
    EpisodeAcc acc;
    ErrorAcc errs;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t groups = (a.n + 3) / 4;
    const int64_t groups_round = (groups + 31) & ~static_cast<int64_t>(31);  // warp-uniform trip count
    unsigned long long ticket;
    const uint64_t t0 = a.t + clock_begin(a.clock, ticket);
    const bool packed = a.ep_len == nullptr;
    const bool vec_ok = aligned(a.pos, 16) && aligned(a.ep_len, 16) && aligned(a.ep_ret, 16);
    for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < groups_round; v += stride) {
        const int64_t base = v * 4;
        const uint32_t g0 = static_cast<uint32_t>(a.env_offset + base);
        Quad q;
        bool valid[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) valid[j] = base + j < a.n;
        const bool full = valid[3] && vec_ok;
        if (full) {
            const int4 p4 = *reinterpret_cast<const int4*>(a.pos + base);
            const float4 r4 = *reinterpret_cast<const float4*>(a.ep_ret + base);
            if (packed) {
                unpack4(p4, q.pos, q.len);
            } else {
                const int4 l4 = *reinterpret_cast<const int4*>(a.ep_len + base);
                q.pos[0] = p4.x; q.pos[1] = p4.y; q.pos[2] = p4.z; q.pos[3] = p4.w;
                q.len[0] = l4.x; q.len[1] = l4.y; q.len[2] = l4.z; q.len[3] = l4.w;
            }
            q.ret[0] = r4.x; q.ret[1] = r4.y; q.ret[2] = r4.z; q.ret[3] = r4.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                q.pos[j] = a.start_cell;
                q.len[j] = 0;
                if (valid[j]) load_state1(a, base + j, q.pos[j], q.len[j]);
                q.ret[j] = valid[j] ? a.ep_ret[base + j] : 0.0f;
            }
        }
        uint32_t off[4] = {0u, 0u, 0u, 0u};
        int32_t lay[4] = {0, 0, 0, 0};
        if constexpr (HET) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (a.layout) lay[j] = valid[j] ? __ldg(a.layout + base + j) : 0;
                else lay[j] = static_cast<int32_t>((g0 + j) % static_cast<uint32_t>(a.n_layouts));
                off[j] = static_cast<uint32_t>(lay[j]) * static_cast<uint32_t>(a.layout_len);
                if (!valid[j]) q.pos[j] = __ldg(a.start_cells + lay[j]);   // dummy env: any cell of its own table
            }
        }
        // lanes past the end of the batch (only in the last, ragged quad) carry dummy envs: masks keep them out of
        // statistics, errors and stores
        const uint32_t valid_bytes = (valid[0] ? 0x1u : 0u) | (valid[1] ? 0x100u : 0u) | (valid[2] ? 0x10000u : 0u) |
                                     (valid[3] ? 0x1000000u : 0u);
        const uint32_t valid_bits = (valid[0] ? 1u : 0u) | (valid[1] ? 2u : 0u) | (valid[2] ? 4u : 0u) | (valid[3] ? 8u : 0u);
        const bool act_vec = EXT_ACTIONS && valid[3] && ((a.n & 3) == 0) && aligned(a.action, 16);
        for (int k = 0; k < a.K; ++k) {
            const uint64_t t = t0 + static_cast<uint64_t>(k);
            const int64_t o = static_cast<int64_t>(k) * a.n + base;
            uint32_t wt[4] = {0u, 0u, 0u, 0u};
            int32_t act[4];
            if constexpr (SLOTS == 4) group_words4_rk(a.rk, a.seed, g0, t, kStreamTransition, wt);
            if constexpr (EXT_ACTIONS) {
                if (act_vec) {
                    const int4 a4 = __ldg(reinterpret_cast<const int4*>(a.action + o));
                    act[0] = a4.x; act[1] = a4.y; act[2] = a4.z; act[3] = a4.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) act[j] = valid[j] ? __ldg(a.action + o + j) : 0;
                }
            } else {
                uint32_t wp[4];
                group_words4_rk(a.rk, a.seed, g0, t, kStreamPolicy, wp);
#pragma unroll
                for (int j = 0; j < 4; ++j) act[j] = static_cast<int32_t>(wp[j] >> 30);
            }
            uint32_t rejected = 0u;
            if (EXT_ACTIONS && ((act[0] | act[1] | act[2] | act[3]) & ~3) != 0) {  // out-of-range action: checking path
                q.term = q.trunc = q.bad = 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    ErrorAcc scratch;
                    const uint32_t f = transition<SMEM, SLOTS>(a, tab + off[j], g0 + j, q.pos[j], q.len[j], q.ret[j], act[j],
                                                               wt[j], q.rew[j], valid[j] ? errs : scratch);
                    q.term |= (f & 1u) << (8 * j);
                    q.trunc |= ((f >> 1) & 1u) << (8 * j);
                    rejected += valid[j] ? ((f >> 2) & 1u) : 0u;
                }
            } else {
                transition4<SMEM, SLOTS>(a, tab, q, act, wt, off);  // off: all zero, and folded away, unless HET
                q.bad &= valid_bits;
                if (q.bad) {
                    const uint4 e = record_bad4<SMEM, SLOTS>(tab, q.bad, g0, make_int4(q.pos[0], q.pos[1], q.pos[2], q.pos[3]),
                                                             make_int4(act[0], act[1], act[2], act[3]),
                                                             make_uint4(off[0], off[1], off[2], off[3]));
                    errs.count += e.x;
                    errs.g = e.y;
                    errs.detail = e.z;
                    errs.kind = e.w;
                    rejected = e.x;
                }
            }
            acc.steps += __popc(valid_bits) - rejected;
            if constexpr (EMIT) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (valid[j]) {
                        if (a.reward) a.reward[o + j] = q.rew[j];
                        if (a.flags) a.flags[o + j] = static_cast<uint8_t>(((q.term >> (8 * j)) & 1u) | (((q.trunc >> (8 * j)) & 1u) << 1));
                    }
                }
            }
            const uint32_t fin = (q.term | q.trunc) & valid_bytes;
            if (__ballot_sync(0xffffffffu, fin != 0u)) {  // warp vote: skip the reset path where nobody finished
                if (fin) finish4(a, q, fin, acc, HET ? lay : nullptr);
            }
            if constexpr (EMIT) {
                if (a.obs) {  // trajectory of observations [K, n, 2]: (row, col) after auto-reset, like step()
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (valid[j]) reinterpret_cast<int2*>(a.obs)[o + j] = cell_to_obs(a, q.pos[j]);
                    }
                }
            }
        }
        if (full) {
            if (packed) {
                *reinterpret_cast<int4*>(a.pos + base) = pack4(q.pos, q.len);
            } else {
                *reinterpret_cast<int4*>(a.pos + base) = make_int4(q.pos[0], q.pos[1], q.pos[2], q.pos[3]);
                *reinterpret_cast<int4*>(a.ep_len + base) = make_int4(q.len[0], q.len[1], q.len[2], q.len[3]);
            }
            *reinterpret_cast<float4*>(a.ep_ret + base) = make_float4(q.ret[0], q.ret[1], q.ret[2], q.ret[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (valid[j]) {
                    store_state1(a, base + j, q.pos[j], q.len[j]);
                    a.ep_ret[base + j] = q.ret[j];
                }
            }
        }
    }
    cta_epilogue(acc, errs, a.stats, a.err, a.clock, ticket, static_cast<unsigned long long>(a.K));
}

__global__ void __launch_bounds__(kThreads) gridworld_reset_kernel(int32_t start_cell, int32_t cols, int32_t* pos,
                                                                   int32_t* ep_len, float* ep_ret, int32_t* obs,
                                                                   const uint8_t* mask, int64_t n) {
    const int32_t r0 = start_cell / cols, c0 = start_cell - r0 * cols;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        if (mask && !mask[i]) continue;
        pos[i] = start_cell;              // packed layout (ep_len == NULL): counter bits are zero
        if (ep_len) ep_len[i] = 0;
        ep_ret[i] = 0.0f;
        if (obs) { obs[2 * i] = r0; obs[2 * i + 1] = c0; }
    }
}

// Mixed-layout reset: env i goes back to the start cell of ITS layout (layout[i], or (env_offset + i) % n_layouts).
__global__ void __launch_bounds__(kThreads) gridworld_reset_layouts_kernel(int32_t n_layouts, const int32_t* __restrict__ layout,
                                                                           const int32_t* __restrict__ start_cells, int32_t cols,
                                                                           int32_t* pos, int32_t* ep_len, float* ep_ret,
                                                                           int32_t* obs, const uint8_t* mask,
                                                                           uint64_t env_offset, int64_t n) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        if (mask && !mask[i]) continue;
        const int32_t lay = layout ? __ldg(layout + i) : static_cast<int32_t>((env_offset + i) % static_cast<uint64_t>(n_layouts));
        const int32_t start = __ldg(start_cells + lay);
        pos[i] = start;                   // packed layout (ep_len == NULL): counter bits are zero
        if (ep_len) ep_len[i] = 0;
        ep_ret[i] = 0.0f;
        if (obs) { obs[2 * i] = start / cols; obs[2 * i + 1] = start % cols; }
    }
}

// ---- persistent step server ---------------------------------------------------------------------------------------------
// K steps of ONE env set without a launch per step, for a policy that lives on the device in a kernel of its own.  The
// server is resident for all K steps, keeps the packed state of its envs in registers and exchanges one chunk of envs per
// CTA with the action producer through two sequence words per chunk:
//     producer: (k > 0: wait obs_seq[c] >= k)  read obs of step k-1  ->  write actions of step k  ->  act_seq[c] = k + 1
//     server  : wait act_seq[c] >= k + 1       read actions          ->  step, write outputs      ->  obs_seq[c] = k + 1
// (release stores / acquire loads at gpu scope; everything another kernel wrote while this one runs is read through L2).
// Same Philox words and table walk as K step() calls: results are bit-identical (tests/test_gpu_gridworld_serve.py).
// Both kernels must be resident together: the server runs one CTA of kServeThreads (256) per chunk and at most
// EF_SERVE_CHUNKS_PER_SM (2) chunks per SM -- 48 K of every SM's 64 K registers -- which leaves room for the producer.
// Publishing a chunk is `bar.sync` + one release store by thread 0 (cumulative: the CTA's stores are ordered before it).  A wait that is not satisfied within ~5 s gives up, sets
// `status` and ends the kernel instead of hanging the GPU.
#ifndef EF_SERVE_THREADS
#define EF_SERVE_THREADS 256
#endif
#ifndef EF_SERVE_THREADFENCE
#define EF_SERVE_THREADFENCE 0     // 1: every thread fences before the publishing barrier (measured slower: 11.5 vs 10.3 us per step)
#endif
#ifndef EF_SERVE_PREDRAW
#define EF_SERVE_PREDRAW 1         // the step's Philox draws are computed while the CTA waits for the step's actions
#endif
#ifndef EF_SERVE_NANOSLEEP
#define EF_SERVE_NANOSLEEP 100       // back-off between two polls of a sequence word (ns); 0 = spin
#endif
constexpr int kServeThreads = EF_SERVE_THREADS;   // CTAs of this size, EF_SERVE_CHUNKS_PER_SM of them per SM: every chunk is its
constexpr int kServeQ = 4;           // own dependency chain (wait, step, publish); the chains of one SM hide each other's latencies
constexpr int kServeProducerThreads = 128;
constexpr int kProducerBatch = 4;    // quads a producer thread loads before it stores (a chunk of 896 quads = two batches)

__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Thread 0 spins until *seq >= want (or gives up); returns false for the whole CTA after a give-up.
__device__ __forceinline__ bool serve_wait(const uint32_t* seq, uint32_t want, uint32_t* status, int* s_ok) {
    if (threadIdx.x == 0) {
        unsigned long long t0 = 0ull;
        uint32_t polls = 0u;
        int ok = 1;
        while (ld_acquire_gpu_u32(seq) < want) {
            if (EF_SERVE_NANOSLEEP) __nanosleep(EF_SERVE_NANOSLEEP);
            if ((++polls & 1023u) == 0u) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
                if (t0 == 0ull) t0 = now;
                else if (now - t0 > 5000000000ull) {
                    ok = 0;
                    if (status) atomicExch(status, 1u);
                    break;
                }
            }
        }
        *s_ok = ok;
    }
    __syncthreads();
    return *s_ok != 0;
}

// 96 registers x 256 threads x 2 CTAs = 48 K of the SM's 64 K registers: the producer's CTAs (128 threads, ~40 registers) fit.
template <int SLOTS>
__global__ void __maxnreg__(96) gridworld_serve_kernel(const GridArgs a) {
    const uint2* __restrict__ tab = stage_table<true>(a);
    __shared__ int s_ok;
    EpisodeAcc acc;
    ErrorAcc errs;
    const uint32_t c = blockIdx.x;
    const int64_t q_begin = static_cast<int64_t>(c) * a.chunk_quads;
    const int64_t quads = a.n >> 2;
    const int32_t count = static_cast<int32_t>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(a.chunk_quads), quads - q_begin)));
    const uint64_t t0 = a.t + clock_steps(clock_fetch(a.clock));
    // state of this thread's quads: registers for all K steps
    uint32_t w[kServeQ][4];
    float ret[kServeQ][4];
#pragma unroll
    for (int j = 0; j < kServeQ; ++j) {
        const int32_t local = j * kServeThreads + static_cast<int32_t>(threadIdx.x);
        if (local < count) {
            const int64_t base = (q_begin + local) * 4;
            const uint4 p4 = *reinterpret_cast<const uint4*>(a.pos + base);
            const float4 r4 = *reinterpret_cast<const float4*>(a.ep_ret + base);
            w[j][0] = p4.x; w[j][1] = p4.y; w[j][2] = p4.z; w[j][3] = p4.w;
            ret[j][0] = r4.x; ret[j][1] = r4.y; ret[j][2] = r4.z; ret[j][3] = r4.w;
        }
    }
    // The outcome draws of step k depend on (seed, env, t0 + k) only: every thread computes those of its quads BEFORE the
    // wait for the step's actions and parks them in shared memory -- a third of a step's instructions off the
    // actions -> observations critical path (EF_SERVE_PREDRAW=0: computed after the wait, as step() does; same bits).
    __shared__ uint4 s_draw[(EF_SERVE_PREDRAW && SLOTS == 4) ? kServeQ * kServeThreads : 1];
    int k = 0;
    for (; k < a.K; ++k) {
        const uint64_t t = t0 + static_cast<uint64_t>(k);
        if constexpr (EF_SERVE_PREDRAW && SLOTS == 4) {
#pragma unroll
            for (int j = 0; j < kServeQ; ++j) {
                const int32_t local = j * kServeThreads + static_cast<int32_t>(threadIdx.x);
                if (local < count) {
                    uint32_t r[4];
                    group_words4_rk(a.rk, a.seed, static_cast<uint32_t>(a.env_offset + (q_begin + local) * 4), t, kStreamTransition, r);
                    s_draw[local] = make_uint4(r[0], r[1], r[2], r[3]);
                }
            }
        }
        if (!serve_wait(a.act_seq + c, static_cast<uint32_t>(k) + 1u, a.status, &s_ok)) break;
#pragma unroll
        for (int j = 0; j < kServeQ; ++j) {
            const int32_t local = j * kServeThreads + static_cast<int32_t>(threadIdx.x);
            if (local < count) {
                const int64_t base = (q_begin + local) * 4;
                QuadIn in;
                in.p4 = make_int4(static_cast<int32_t>(w[j][0]), static_cast<int32_t>(w[j][1]), static_cast<int32_t>(w[j][2]),
                                  static_cast<int32_t>(w[j][3]));
                in.r4 = make_float4(ret[j][0], ret[j][1], ret[j][2], ret[j][3]);
                in.a4 = __ldcg(reinterpret_cast<const int4*>(a.action + base));   // written by the producer while we run: L2
                PackedQuad q;
                step_quad_packed<true, SLOTS, true, false>(a, tab, base, static_cast<uint32_t>(a.env_offset + base), t, in, false,
                                                           acc, errs, q,
                                                           (EF_SERVE_PREDRAW && SLOTS == 4) ? &s_draw[local] : nullptr);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    w[j][e] = q.w[e];
                    ret[j][e] = q.ret[e];
                }
                const int32_t p[4] = {static_cast<int32_t>(q.w[0] & 0xffffu), static_cast<int32_t>(q.w[1] & 0xffffu),
                                      static_cast<int32_t>(q.w[2] & 0xffffu), static_cast<int32_t>(q.w[3] & 0xffffu)};
                store_obs4<false>(a, a.obs, base, p);
                *reinterpret_cast<float4*>(a.reward + base) = make_float4(q.rew[0], q.rew[1], q.rew[2], q.rew[3]);
                *reinterpret_cast<uint32_t*>(a.terminated + base) = q.term;
                *reinterpret_cast<uint32_t*>(a.truncated + base) = q.trunc;
            }
        }
#if EF_SERVE_THREADFENCE
        __threadfence();   // every thread drains its own stores (in parallel) instead of thread 0's release waiting for all
#endif
        __syncthreads();   // every thread's stores precede the barrier; thread 0's release store is cumulative over them
        if (threadIdx.x == 0) st_release_gpu_u32(a.obs_seq + c, static_cast<uint32_t>(k) + 1u);
    }
    // state back to memory, so that step() / rollout() continue where the server stopped
#pragma unroll
    for (int j = 0; j < kServeQ; ++j) {
        const int32_t local = j * kServeThreads + static_cast<int32_t>(threadIdx.x);
        if (local < count) {
            const int64_t base = (q_begin + local) * 4;
            *reinterpret_cast<uint4*>(a.pos + base) = make_uint4(w[j][0], w[j][1], w[j][2], w[j][3]);
            *reinterpret_cast<float4*>(a.ep_ret + base) = make_float4(ret[j][0], ret[j][1], ret[j][2], ret[j][3]);
        }
    }
    cta_epilogue<kServeThreads / 32>(acc, errs, a.stats, a.err, a.clock, 0ull, static_cast<unsigned long long>(k));
}

// A stand-in for the user's policy kernel (tests, bench): the action of env i at step k is a fixed function of the
// observation the env returned for step k - 1 (the observation buffer's content at launch for k = 0),
//     action = (row * 3 + col + k) & 3,
// computed chunk by chunk against the server's sequence words.  128 threads per chunk, resident next to the server.
__global__ void __launch_bounds__(kServeProducerThreads) policy_demo_serve_kernel(const int32_t* obs, int32_t* action, const uint32_t* obs_seq,
                                                                     uint32_t* act_seq, int32_t chunk_quads, int32_t K,
                                                                     uint32_t* status, int64_t n) {
    __shared__ int s_ok;
    const uint32_t c = blockIdx.x;
    const int64_t q_begin = static_cast<int64_t>(c) * chunk_quads;
    const int64_t quads = n >> 2;
    const int32_t count = static_cast<int32_t>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(chunk_quads), quads - q_begin)));
    for (int k = 0; k < K; ++k) {
        if (k > 0 && !serve_wait(obs_seq + c, static_cast<uint32_t>(k), status, &s_ok)) break;
        // kProducerBatch quads per thread at a time, all their loads in flight before the first is used: a chunk is then
        // two L2 round trips instead of one per quad (the plain loop cost the served step 2-3 us)
        for (int32_t first = 0; first < count; first += kServeProducerThreads * kProducerBatch) {
            int4 o0[kProducerBatch], o1[kProducerBatch];
#pragma unroll
            for (int b = 0; b < kProducerBatch; ++b) {
                const int32_t local = first + b * kServeProducerThreads + static_cast<int32_t>(threadIdx.x);
                if (local < count) {
                    const int64_t base = (q_begin + local) * 4;
                    o0[b] = __ldcg(reinterpret_cast<const int4*>(obs + 2 * base));       // (row, col) of envs 0, 1
                    o1[b] = __ldcg(reinterpret_cast<const int4*>(obs + 2 * base) + 1);   // envs 2, 3
                }
            }
#pragma unroll
            for (int b = 0; b < kProducerBatch; ++b) {
                const int32_t local = first + b * kServeProducerThreads + static_cast<int32_t>(threadIdx.x);
                if (local < count) {
                    const int64_t base = (q_begin + local) * 4;
                    *reinterpret_cast<int4*>(action + base) =
                        make_int4((o0[b].x * 3 + o0[b].y + k) & 3, (o0[b].z * 3 + o0[b].w + k) & 3,
                                  (o1[b].x * 3 + o1[b].y + k) & 3, (o1[b].z * 3 + o1[b].w + k) & 3);
                }
            }
        }
#if EF_SERVE_THREADFENCE
        __threadfence();
#endif
        __syncthreads();
        if (threadIdx.x == 0) st_release_gpu_u32(act_seq + c, static_cast<uint32_t>(k) + 1u);
    }
}

// The stand-in policy as an ordinary kernel (one launch per step): what the step server is measured against.
__global__ void __launch_bounds__(kThreads) policy_demo_step_kernel(const int32_t* __restrict__ obs, int32_t* __restrict__ action,
                                                                    int32_t k, int64_t quads) {
    for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < quads;
         v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int4 o0 = reinterpret_cast<const int4*>(obs)[2 * v], o1 = reinterpret_cast<const int4*>(obs)[2 * v + 1];
        reinterpret_cast<int4*>(action)[v] = make_int4((o0.x * 3 + o0.y + k) & 3, (o0.z * 3 + o0.w + k) & 3,
                                                       (o1.x * 3 + o1.y + k) & 3, (o1.z * 3 + o1.w + k) & 3);
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------------
constexpr size_t kMaxSmemTable = 96 * 1024;

static int fill_args(const ef_grid_desc* g, GridArgs& a) {
    if (!g || !g->table) return EF_ERR_INVALID_ARGUMENT;
    if (g->rows <= 0 || g->cols <= 0 || (g->n_slots != 1 && g->n_slots != 4)) return EF_ERR_INVALID_ARGUMENT;
    const int64_t cells = static_cast<int64_t>(g->rows) * g->cols;
    if (cells > 65536) return EF_ERR_UNSUPPORTED;
    const int64_t tcells = g->table_cells > 0 ? g->table_cells : cells;
    if (tcells < cells || tcells > 65536) return EF_ERR_INVALID_ARGUMENT;
    if (g->start_cell < 0 || g->start_cell >= tcells || g->idx_cap < 0 || g->idx_cap > 3) return EF_ERR_INVALID_ARGUMENT;
    if (!aligned(g->table, 8)) return EF_ERR_INVALID_ARGUMENT;
    a.table = reinterpret_cast<const uint2*>(g->table);
    a.table_len = static_cast<int32_t>(tcells * 4 * g->n_slots);
    a.cols = g->cols;
    a.cols_magic = g->cols > 1 ? static_cast<uint32_t>(((1ull << 32) + g->cols - 1) / static_cast<uint64_t>(g->cols)) : 0u;
    a.thr0 = g->idx_cap > 0 ? g->thr[0] : (1ull << 32);
    a.thr1 = g->idx_cap > 1 ? g->thr[1] : (1ull << 32);
    a.thr2 = g->idx_cap > 2 ? g->thr[2] : (1ull << 32);
    a.idx_cap = g->idx_cap;
    a.start_cell = g->start_cell;
    a.limit = g->episode_limit;
    return EF_OK;
}

// Packed state (ep_len == NULL) needs a counter that provably stays below 2^16: auto-reset on and 0 < limit <= 65535.
static bool packed_ok(const GridArgs& a, int32_t autoreset) { return autoreset != 0 && a.limit > 0 && a.limit <= 65535; }

template <typename Kernel>
static int launch(Kernel kernel, const GridArgs& a, bool smem, int grid, cudaStream_t stream, size_t extra_smem = 0) {
    const size_t bytes = (smem ? static_cast<size_t>(a.table_len) * sizeof(uint2) : 0) + extra_smem;
    if (bytes > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
        if (e != cudaSuccess) return cuda_status(e);
    }
    return launch_kernel(kernel, a, grid, bytes, stream);
}

}  // namespace ef

using namespace ef;

// Shared body of ef_gridworld_step (W8 = false: int32 actions / observations) and ef_gridworld_step_u8 (W8 = true).
template <bool W8>
static int gridworld_step_impl(const ef_grid_desc* grid, int32_t* pos, int32_t* ep_len, float* ep_ret, const void* action,
                               const uint32_t* draw, uint64_t seed, uint64_t t, uint64_t env_offset, void* obs,
                               float* reward, uint8_t* terminated, uint8_t* truncated, void* final_obs, double* stats,
                               uint32_t* err, uint64_t* clock, int64_t n, int32_t autoreset, void* stream,
                               int32_t hints = 0) {
    GridArgs a{};
    a.cold = (hints & EF_STEP_COLD_INPUTS) ? 1 : 0;
    if (int s = fill_args(grid, a)) return s;
    if (n < 0 || !pos || !ep_ret) return EF_ERR_INVALID_ARGUMENT;
    if (W8 && (grid->rows > 256 || grid->cols > 256)) return EF_ERR_UNSUPPORTED;  // (row, col) must fit a byte each
    if (!ep_len && !packed_ok(a, autoreset)) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    a.pos = pos; a.ep_len = ep_len; a.ep_ret = ep_ret; a.action = static_cast<const int32_t*>(action); a.draw = draw;
    a.seed = seed; a.t = t; a.env_offset = env_offset;
    a.obs = static_cast<int32_t*>(obs); a.reward = reward; a.terminated = terminated; a.truncated = truncated;
    a.final_obs = static_cast<int32_t*>(final_obs);
    a.stats = stats; a.err = err; a.n = n; a.autoreset = autoreset;
    a.clock = reinterpret_cast<unsigned long long*>(clock);
    // a quad's accesses: 16 B of int32 actions / 2 x 16 B of int32 observations, or 4 B / 8 B in the narrow format
    const bool vec = draw == nullptr && aligned(pos, 16) && aligned(ep_len, 16) && aligned(ep_ret, 16) &&
                     aligned(action, W8 ? 4 : 16) && aligned(obs, W8 ? 8 : 16) && aligned(reward, 16) &&
                     aligned(terminated, 4) && aligned(truncated, 4) && aligned(final_obs, W8 ? 8 : 16);
    const bool smem = static_cast<size_t>(a.table_len) * sizeof(uint2) <= kMaxSmemTable;
    const cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Small batches (one pass fits the resident grid): EF_STEP_Q quads per thread, all loads in flight, persistent grid.
    // Large batches: fewer quads per thread, more resident warps, CTAs scheduled dynamically over many passes.
    const int64_t quads = (n + 3) / 4;
    const bool small = quads <= static_cast<int64_t>(sm_count()) * EF_STEP_GRID_PER_SM * kThreads * EF_STEP_Q * EF_STEP_PASSES;
    // One-pass grid: spread the quads over ALL resident CTA slots (at least a warp's worth each) rather than filling
    // CTAs to Q * 256 quads -- 2^20 envs would otherwise make 256 full CTAs, leaving 40 of the 148 SMs with one CTA and
    // 108 with two, and the launch lasts as long as the doubly loaded SMs.
    const int grid_v = static_cast<int>(std::max<int64_t>(
        1, std::min<int64_t>(static_cast<int64_t>(sm_count()) * EF_STEP_GRID_PER_SM, (quads + 31) / 32)));
    const int grid_b = grid_for((quads + EF_STEP_BIG_Q - 1) / EF_STEP_BIG_Q, EF_STEP_BIG_GRID_PER_SM);
    const int grid_s = grid_for(n, 8);
    const int grid_used = small ? grid_v : grid_b;
    a.quads_per_cta = ((n >> 2) + grid_used - 1) / grid_used;  // full quads only; the ragged tail is handled separately
    a.rk = philox_round_keys(seed);
    const bool packed = ep_len == nullptr;
    const bool fast = action && obs && reward && terminated && truncated && !final_obs && autoreset != 0;
#define EF_LAUNCH_VEC(SM, SL, PK, FS)                                                                                    \
    if (small && FS && a.cold) /* the hinted instantiation exists for the one-wave kernel of the common call shape */    \
        return launch(gridworld_step_kernel<SM, SL, EF_STEP_Q, EF_STEP_MINBLOCKS, PK, FS, false, W8, FS>, a, SM, grid_v, \
                      st);                                                                                               \
    return small ? launch(gridworld_step_kernel<SM, SL, EF_STEP_Q, EF_STEP_MINBLOCKS, PK, FS, false, W8>, a, SM, grid_v, \
                          st)                                                                                            \
                 : launch(gridworld_step_kernel<SM, SL, EF_STEP_BIG_Q, EF_STEP_BIG_MINBLOCKS, PK, FS, false, W8>, a, SM, \
                          grid_b, st)
    // The narrow format has the compile-time call-shape specialisation (FAST) only: its callers are actor loops that
    // pass actions and take every output; other call shapes (and tables too large for shared memory) use the generic
    // build, which for W8 is instantiated for the packed state only when the table is staged.
#define EF_DISPATCH(SM, SL)                                                                                              \
    if (!vec) return launch(gridworld_step_scalar_kernel<SM, SL, W8>, a, SM, grid_s, st);                                \
    if (SM && fast) { if (packed) { EF_LAUNCH_VEC(SM, SL, true, SM); } else { EF_LAUNCH_VEC(SM, SL, false, SM); } }      \
    if (packed) { EF_LAUNCH_VEC(SM, SL, true, false); } else { EF_LAUNCH_VEC(SM, SL, false, false); }
    if (smem) {
        if (grid->n_slots == 1) { EF_DISPATCH(true, 1); } else { EF_DISPATCH(true, 4); }
    } else {
        if (grid->n_slots == 1) { EF_DISPATCH(false, 1); } else { EF_DISPATCH(false, 4); }
    }
#undef EF_LAUNCH_VEC
#undef EF_DISPATCH
}

extern "C" int ef_gridworld_step(const ef_grid_desc* grid, int32_t* pos, int32_t* ep_len, float* ep_ret,
                                 const int32_t* action, const uint32_t* draw, uint64_t seed, uint64_t t,
                                 uint64_t env_offset, int32_t* obs, float* reward, uint8_t* terminated,
                                 uint8_t* truncated, int32_t* final_obs, double* stats, uint32_t* err,
                                 uint64_t* clock, int64_t n, int32_t autoreset, void* stream) {
    return gridworld_step_impl<false>(grid, pos, ep_len, ep_ret, action, draw, seed, t, env_offset, obs, reward, terminated,
                                      truncated, final_obs, stats, err, clock, n, autoreset, stream);
}

extern "C" int ef_gridworld_step_u8(const ef_grid_desc* grid, int32_t* pos, int32_t* ep_len, float* ep_ret,
                                    const uint8_t* action, const uint32_t* draw, uint64_t seed, uint64_t t,
                                    uint64_t env_offset, uint8_t* obs, float* reward, uint8_t* terminated,
                                    uint8_t* truncated, uint8_t* final_obs, double* stats, uint32_t* err,
                                    uint64_t* clock, int64_t n, int32_t autoreset, void* stream) {
    return gridworld_step_impl<true>(grid, pos, ep_len, ep_ret, action, draw, seed, t, env_offset, obs, reward, terminated,
                                     truncated, final_obs, stats, err, clock, n, autoreset, stream);
}

extern "C" int ef_gridworld_step_v(const ef_gridworld_step_args* x) {
    if (!x) return EF_ERR_INVALID_ARGUMENT;
    if (x->wire & EF_WIRE_U8)  // action / obs / final_obs point at bytes (see ef_gridworld_step_u8)
        return gridworld_step_impl<true>(x->grid, x->pos, x->ep_len, x->ep_ret, x->action, x->draw, x->seed, x->t, x->env_offset,
                                         x->obs, x->reward, x->terminated, x->truncated, x->final_obs, x->stats, x->err,
                                         x->clock, x->n, x->autoreset, x->stream, x->wire);
    return gridworld_step_impl<false>(x->grid, x->pos, x->ep_len, x->ep_ret, static_cast<const int32_t*>(x->action), x->draw,
                                      x->seed, x->t, x->env_offset, x->obs, x->reward, x->terminated, x->truncated,
                                      x->final_obs, x->stats, x->err, x->clock, x->n, x->autoreset, x->stream, x->wire);
}

extern "C" int ef_gridworld_step_layouts(const ef_grid_desc* grid, int32_t n_layouts, const int32_t* layout_index,
                                         const int32_t* start_cells, int32_t* pos, int32_t* ep_len, float* ep_ret,
                                         const int32_t* action, const uint32_t* draw, uint64_t seed, uint64_t t,
                                         uint64_t env_offset, int32_t* obs, float* reward, uint8_t* terminated,
                                         uint8_t* truncated, int32_t* final_obs, double* stats, uint32_t* err,
                                         uint64_t* clock, int64_t n, int32_t autoreset, void* stream) {
    GridArgs a{};
    if (int s = fill_args(grid, a)) return s;
    if (n_layouts < 1 || !start_cells || n < 0 || !pos || !ep_ret) return EF_ERR_INVALID_ARGUMENT;
    if (static_cast<int64_t>(a.table_len) * n_layouts > (1ll << 30)) return EF_ERR_UNSUPPORTED;
    if (!ep_len && !packed_ok(a, autoreset)) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    a.layout = layout_index; a.start_cells = start_cells; a.n_layouts = n_layouts; a.layout_len = a.table_len;
    a.pos = pos; a.ep_len = ep_len; a.ep_ret = ep_ret; a.action = action; a.draw = draw;
    a.seed = seed; a.t = t; a.env_offset = env_offset;
    a.obs = obs; a.reward = reward; a.terminated = terminated; a.truncated = truncated; a.final_obs = final_obs;
    a.stats = stats; a.err = err; a.n = n; a.autoreset = autoreset;
    a.clock = reinterpret_cast<unsigned long long*>(clock);
    // the staged table is all layouts back to back when that fits in shared memory, else every lookup goes through L1/L2
    const int64_t total = static_cast<int64_t>(a.table_len) * n_layouts;
    const bool smem = static_cast<size_t>(total) * sizeof(uint2) <= kMaxSmemTable;
    a.table_len = static_cast<int32_t>(total);  // what stage_table copies
    const int grid_s = grid_for(n, 8);
    const cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Vector kernel (four envs per thread, 128-bit accesses, dynamically scheduled CTAs) whenever the buffers allow it;
    // injected draws and unaligned views take the one-env-per-thread kernel.
    const bool vec = draw == nullptr && n >= 4 && aligned(pos, 16) && aligned(ep_len, 16) && aligned(ep_ret, 16) &&
                     aligned(action, 16) && aligned(obs, 16) && aligned(reward, 16) && aligned(terminated, 4) &&
                     aligned(truncated, 4) && aligned(final_obs, 16) && aligned(layout_index, 16);
    if (vec) {
        const int64_t quads = n >> 2;
        // Q = 2 quads per thread like the single-layout large-batch variant, but compiled for 2 CTAs/SM: the per-env table
        // offsets and layout indices push the 3-CTA (80-register) build into local-memory spills (156 B of spill stores).
        // Persistent grid (2 resident CTAs per SM, each looping over its contiguous share) when the tables are staged:
        // every CTA copies ALL layouts into its shared memory (64 KB for eight 8x8 layouts), so 64 dynamically scheduled
        // CTAs per SM would move more table bytes out of L2 than the batch moves through HBM.
        static const int het_per_sm = [] {
            const char* e = getenv("ENVFORGE_HET_GRID_PER_SM");  // A/B switch; 0 = the dynamic grid for staged tables too
            return e ? atoi(e) : 2;
        }();
        const int per_sm = (smem && het_per_sm > 0) ? het_per_sm : EF_STEP_BIG_GRID_PER_SM;
        const int grid_b = grid_for((quads + EF_STEP_BIG_Q - 1) / EF_STEP_BIG_Q, per_sm);
        a.quads_per_cta = (quads + grid_b - 1) / grid_b;
        a.rk = philox_round_keys(seed);
        const bool packed = ep_len == nullptr;
#define EF_HET(SM, SL)                                                                                                      \
    return packed ? launch(gridworld_step_kernel<SM, SL, EF_STEP_BIG_Q, 2, true, false, true>, a, SM, grid_b, st)           \
                  : launch(gridworld_step_kernel<SM, SL, EF_STEP_BIG_Q, 2, false, false, true>, a, SM, grid_b, st)
        if (smem) {
            if (grid->n_slots == 1) { EF_HET(true, 1); } else { EF_HET(true, 4); }
        } else {
            if (grid->n_slots == 1) { EF_HET(false, 1); } else { EF_HET(false, 4); }
        }
#undef EF_HET
    }
    if (smem) {
        if (grid->n_slots == 1) return launch(gridworld_step_scalar_kernel<true, 1>, a, true, grid_s, st);
        return launch(gridworld_step_scalar_kernel<true, 4>, a, true, grid_s, st);
    }
    if (grid->n_slots == 1) return launch(gridworld_step_scalar_kernel<false, 1>, a, false, grid_s, st);
    return launch(gridworld_step_scalar_kernel<false, 4>, a, false, grid_s, st);
}

extern "C" int ef_gridworld_step_pcg64(const ef_grid_desc* grid, const uint64_t* thr53, int32_t idx_cap53, int32_t* pos,
                                       int32_t* ep_len, float* ep_ret, uint64_t* rng_state, const uint64_t* rng_inc,
                                       const int32_t* action, uint64_t env_offset, int32_t* obs, float* reward,
                                       uint8_t* terminated, uint8_t* truncated, int32_t* final_obs, double* stats,
                                       uint32_t* err, int64_t n, int32_t autoreset, void* stream) {
    GridArgs a{};
    if (int s = fill_args(grid, a)) return s;
    if (n < 0 || !pos || !ep_ret || !rng_state || !rng_inc || !action || !thr53) return EF_ERR_INVALID_ARGUMENT;
    if (!ep_len && !packed_ok(a, autoreset)) return EF_ERR_INVALID_ARGUMENT;
    if (!aligned(rng_state, 16) || !aligned(rng_inc, 16) || idx_cap53 < 0 || idx_cap53 > 3) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    a.pos = pos; a.ep_len = ep_len; a.ep_ret = ep_ret; a.action = action; a.env_offset = env_offset;
    a.obs = obs; a.reward = reward; a.terminated = terminated; a.truncated = truncated; a.final_obs = final_obs;
    a.stats = stats; a.err = err; a.n = n; a.autoreset = autoreset;
    Pcg64Args r{};
    r.thr[0] = thr53[0]; r.thr[1] = thr53[1]; r.thr[2] = thr53[2];
    r.idx_cap = idx_cap53;
    r.rng_state = reinterpret_cast<ulonglong2*>(rng_state);
    r.rng_inc = reinterpret_cast<const ulonglong2*>(rng_inc);
    const bool smem = static_cast<size_t>(a.table_len) * sizeof(uint2) <= kMaxSmemTable;
    const size_t bytes = smem ? static_cast<size_t>(a.table_len) * sizeof(uint2) : 0;
    const int grid_s = grid_for(n, 8);
    const cudaStream_t st = static_cast<cudaStream_t>(stream);
#define EF_DISPATCH3(SM, SL)                                                                                          \
    do {                                                                                                              \
        auto kernel = gridworld_step_pcg64_kernel<SM, SL>;                                                            \
        if (bytes > 48 * 1024) {                                                                                      \
            const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                                       static_cast<int>(bytes));                                      \
            if (e != cudaSuccess) return cuda_status(e);                                                              \
        }                                                                                                             \
        kernel<<<grid_s, kThreads, bytes, st>>>(a, r);                                                                \
        return cuda_status(cudaGetLastError());                                                                       \
    } while (0)
    if (smem) {
        if (grid->n_slots == 1) EF_DISPATCH3(true, 1); else EF_DISPATCH3(true, 4);
    } else {
        if (grid->n_slots == 1) EF_DISPATCH3(false, 1); else EF_DISPATCH3(false, 4);
    }
#undef EF_DISPATCH3
}

extern "C" int ef_gridworld_rollout(const ef_grid_desc* grid, int32_t* pos, int32_t* ep_len, float* ep_ret,
                                    const int32_t* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K,
                                    float* rewards, uint8_t* flags, double* stats, uint32_t* err, uint64_t* clock,
                                    int64_t n, void* stream) {
    GridArgs a{};
    if (int s = fill_args(grid, a)) return s;
    if (n < 0 || K < 0 || !pos || !ep_ret) return EF_ERR_INVALID_ARGUMENT;
    if (!ep_len && !packed_ok(a, 1)) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0 || K == 0) return EF_OK;
    a.pos = pos; a.ep_len = ep_len; a.ep_ret = ep_ret; a.action = actions;
    a.seed = seed; a.t = t0; a.env_offset = env_offset; a.K = K;
    a.reward = rewards; a.flags = flags; a.stats = stats; a.err = err; a.n = n; a.autoreset = 1;
    a.rk = philox_round_keys(seed);
    a.clock = reinterpret_cast<unsigned long long*>(clock);
    const bool smem = static_cast<size_t>(a.table_len) * sizeof(uint2) <= kMaxSmemTable;
    const bool ext = actions != nullptr, emit = rewards != nullptr || flags != nullptr;
    const cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid_r = grid_for((n + 3) / 4, 8);
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

extern "C" int ef_gridworld_rollout_traj(const ef_grid_desc* grid, int32_t* pos, int32_t* ep_len, float* ep_ret,
                                         const int32_t* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K,
                                         int32_t* obs, float* rewards, uint8_t* flags, double* stats, uint32_t* err,
                                         uint64_t* clock, int64_t n, void* stream) {
    GridArgs a{};
    if (int s = fill_args(grid, a)) return s;
    if (n < 0 || K < 0 || !pos || !ep_ret || !aligned(obs, 8)) return EF_ERR_INVALID_ARGUMENT;
    if (!ep_len && !packed_ok(a, 1)) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0 || K == 0) return EF_OK;
    a.pos = pos; a.ep_len = ep_len; a.ep_ret = ep_ret; a.action = actions;
    a.seed = seed; a.t = t0; a.env_offset = env_offset; a.K = K;
    a.obs = obs; a.reward = rewards; a.flags = flags; a.stats = stats; a.err = err; a.n = n; a.autoreset = 1;
    a.rk = philox_round_keys(seed);
    a.clock = reinterpret_cast<unsigned long long*>(clock);
    const bool smem = static_cast<size_t>(a.table_len) * sizeof(uint2) <= kMaxSmemTable;
    const bool ext = actions != nullptr;
    const cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid_r = grid_for((n + 3) / 4, 8);
#define EF_TRAJ(SM, SL)                                                                                       \
    do {                                                                                                      \
        if (ext) return launch(gridworld_rollout_kernel<SM, SL, true, true>, a, SM, grid_r, st);              \
        return launch(gridworld_rollout_kernel<SM, SL, false, true>, a, SM, grid_r, st);                      \
    } while (0)
    if (smem) {
        if (grid->n_slots == 1) EF_TRAJ(true, 1); else EF_TRAJ(true, 4);
    } else {
        if (grid->n_slots == 1) EF_TRAJ(false, 1); else EF_TRAJ(false, 4);
    }
#undef EF_TRAJ
}

extern "C" int ef_gridworld_rollout_layouts(const ef_grid_desc* grid, int32_t n_layouts, const int32_t* layout_index,
                                            const int32_t* start_cells, int32_t* pos, int32_t* ep_len, float* ep_ret,
                                            const int32_t* actions, uint64_t seed, uint64_t t0, uint64_t env_offset,
                                            int32_t K, int32_t* obs, float* rewards, uint8_t* flags, double* stats,
                                            uint32_t* err, uint64_t* clock, int64_t n, void* stream) {
    GridArgs a{};
    if (int s = fill_args(grid, a)) return s;
    if (n_layouts < 1 || !start_cells || n < 0 || K < 0 || !pos || !ep_ret || !aligned(obs, 8)) return EF_ERR_INVALID_ARGUMENT;
    if (static_cast<int64_t>(a.table_len) * n_layouts > (1ll << 30)) return EF_ERR_UNSUPPORTED;
    if (!ep_len && !packed_ok(a, 1)) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0 || K == 0) return EF_OK;
    a.layout = layout_index; a.start_cells = start_cells; a.n_layouts = n_layouts; a.layout_len = a.table_len;
    a.pos = pos; a.ep_len = ep_len; a.ep_ret = ep_ret; a.action = actions;
    a.seed = seed; a.t = t0; a.env_offset = env_offset; a.K = K;
    a.obs = obs; a.reward = rewards; a.flags = flags; a.stats = stats; a.err = err; a.n = n; a.autoreset = 1;
    a.rk = philox_round_keys(seed);
    a.clock = reinterpret_cast<unsigned long long*>(clock);
    const int64_t total = static_cast<int64_t>(a.table_len) * n_layouts;
    const bool smem = static_cast<size_t>(total) * sizeof(uint2) <= kMaxSmemTable;
    a.table_len = static_cast<int32_t>(total);  // what stage_table copies: all layouts back to back
    const bool ext = actions != nullptr, emit = obs != nullptr || rewards != nullptr || flags != nullptr;
    const cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid_r = grid_for((n + 3) / 4, 8);
#define EF_ROLL_HET(SM, SL)                                                                                         \
    do {                                                                                                            \
        if (ext && emit) return launch(gridworld_rollout_kernel<SM, SL, true, true, true>, a, SM, grid_r, st);     \
        if (ext) return launch(gridworld_rollout_kernel<SM, SL, true, false, true>, a, SM, grid_r, st);            \
        if (emit) return launch(gridworld_rollout_kernel<SM, SL, false, true, true>, a, SM, grid_r, st);           \
        return launch(gridworld_rollout_kernel<SM, SL, false, false, true>, a, SM, grid_r, st);                    \
    } while (0)
    if (smem) {
        if (grid->n_slots == 1) EF_ROLL_HET(true, 1); else EF_ROLL_HET(true, 4);
    } else {
        if (grid->n_slots == 1) EF_ROLL_HET(false, 1); else EF_ROLL_HET(false, 4);
    }
#undef EF_ROLL_HET
}

extern "C" int ef_gridworld_reset_layouts(int32_t n_layouts, const int32_t* layout_index, const int32_t* start_cells, int32_t cols,
                                          int32_t* pos, int32_t* ep_len, float* ep_ret, int32_t* obs, const uint8_t* mask,
                                          uint64_t env_offset, int64_t n, void* stream) {
    if (n < 0 || cols <= 0 || n_layouts < 1 || !start_cells || !pos || !ep_ret) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    gridworld_reset_layouts_kernel<<<grid_for(n, 8), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        n_layouts, layout_index, start_cells, cols, pos, ep_len, ep_ret, obs, mask, env_offset, n);
    return cuda_status(cudaGetLastError());
}

extern "C" int ef_gridworld_reset(int32_t start_cell, int32_t cols, int32_t* pos, int32_t* ep_len, float* ep_ret,
                                  int32_t* obs, const uint8_t* mask, int64_t n, void* stream) {
    if (n < 0 || cols <= 0 || start_cell < 0 || start_cell > 65535 || !pos || !ep_ret) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    gridworld_reset_kernel<<<grid_for(n, 8), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        start_cell, cols, pos, ep_len, ep_ret, obs, mask, n);
    return cuda_status(cudaGetLastError());
}

#ifdef EF_TRACE
// tools-only build: where the step kernels store their timeline (device buffer of 8 * 4 * 1024 * 32 uint64, or NULL)
extern "C" int ef_debug_set_trace(unsigned long long* buf) {
    return cuda_status(cudaMemcpyToSymbol(ef::g_trace_buf, &buf, sizeof(buf)));
}
#endif

extern "C" int ef_gridworld_serve_chunk_envs(int64_t n) {
    // at most EF_SERVE_CHUNKS_PER_SM chunks per SM; whole warps of quads (128 envs), at most kServeQ quads per server thread
    if (n <= 0) return 0;
    const int64_t quads = (n + 3) / 4;
    const int64_t slots = static_cast<int64_t>(sm_count()) * EF_SERVE_CHUNKS_PER_SM;
    int64_t per = (quads + slots - 1) / slots;
    per = (per + 31) / 32 * 32;
    const int64_t cap = static_cast<int64_t>(kServeThreads) * kServeQ;
    if (per > cap) return -1;   // the batch is too large for one resident CTA per SM
    return static_cast<int>(per * 4);
}

extern "C" int ef_gridworld_serve(const ef_grid_desc* grid, int32_t* pos, float* ep_ret, const int32_t* action, uint64_t seed,
                                  uint64_t t0, uint64_t env_offset, int32_t K, int32_t* obs, float* reward,
                                  uint8_t* terminated, uint8_t* truncated, double* stats, uint32_t* err, uint64_t* clock,
                                  uint32_t* act_seq, uint32_t* obs_seq, uint32_t* status, int64_t n, void* stream) {
    GridArgs a{};
    if (int s = fill_args(grid, a)) return s;
    if (n <= 0 || K < 0 || !pos || !ep_ret || !action || !obs || !reward || !terminated || !truncated || !act_seq || !obs_seq)
        return EF_ERR_INVALID_ARGUMENT;
    if (!packed_ok(a, 1)) return EF_ERR_INVALID_ARGUMENT;                      // packed state word only
    if ((n & 3) != 0) return EF_ERR_UNSUPPORTED;                                // whole quads
    if (!aligned(pos, 16) || !aligned(ep_ret, 16) || !aligned(action, 16) || !aligned(obs, 16) || !aligned(reward, 16) ||
        !aligned(terminated, 4) || !aligned(truncated, 4))
        return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    const size_t table_bytes = static_cast<size_t>(a.table_len) * sizeof(uint2);
    if (table_bytes > kMaxSmemTable) return EF_ERR_UNSUPPORTED;
    const int chunk_envs = ef_gridworld_serve_chunk_envs(n);
    if (chunk_envs <= 0) return EF_ERR_UNSUPPORTED;
    if (K == 0) return EF_OK;
    a.pos = pos; a.ep_ret = ep_ret; a.action = action; a.seed = seed; a.t = t0; a.env_offset = env_offset; a.K = K;
    a.obs = obs; a.reward = reward; a.terminated = terminated; a.truncated = truncated;
    a.stats = stats; a.err = err; a.n = n; a.autoreset = 1;
    a.clock = reinterpret_cast<unsigned long long*>(clock);
    a.act_seq = act_seq; a.obs_seq = obs_seq; a.status = status; a.chunk_quads = chunk_envs / 4;
    a.rk = philox_round_keys(seed);
    const int grid_s = static_cast<int>(((n >> 2) + a.chunk_quads - 1) / a.chunk_quads);
    const cudaStream_t st = static_cast<cudaStream_t>(stream);
    {   // CUDA loads kernels lazily, and the FIRST launch of a kernel may synchronise the context: a producer kernel launched
        // for the first time while the server is already spinning would then wait for the server, which waits for it.  The
        // stand-in producer is loaded here, before the server starts; a user's own producer must have been loaded (launched
        // once, or cudaFuncGetAttributes) before ef_gridworld_serve is called.
        cudaFuncAttributes fa;
        const cudaError_t e = cudaFuncGetAttributes(&fa, policy_demo_serve_kernel);
        if (e != cudaSuccess) return cuda_status(e);
    }
    auto go = [&](auto kernel) -> int {
        if (table_bytes > 48 * 1024) {
            const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(table_bytes));
            if (e != cudaSuccess) return cuda_status(e);
        }
        // every chunk's CTA must be resident for all K steps (a chunk that waits for a slot would wait for CTAs that wait for
        // their actions): refuse a table too large to fit the grid next to the kernel's own shared memory
        int per_sm = 0;
        const cudaError_t eo = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kServeThreads, table_bytes);
        if (eo != cudaSuccess) return cuda_status(eo);
        if (static_cast<int64_t>(per_sm) * sm_count() < grid_s) return EF_ERR_UNSUPPORTED;
        kernel<<<grid_s, kServeThreads, table_bytes, st>>>(a);
        return cuda_status(cudaGetLastError());
    };
    return grid->n_slots == 1 ? go(gridworld_serve_kernel<1>) : go(gridworld_serve_kernel<4>);
}

extern "C" int ef_policy_demo_serve(const int32_t* obs, int32_t* action, const uint32_t* obs_seq, uint32_t* act_seq, int32_t K,
                                    uint32_t* status, int64_t n, void* stream) {
    if (n <= 0 || K < 0 || !obs || !action || !obs_seq || !act_seq || (n & 3) != 0) return EF_ERR_INVALID_ARGUMENT;
    const int chunk_envs = ef_gridworld_serve_chunk_envs(n);
    if (chunk_envs <= 0) return EF_ERR_UNSUPPORTED;
    if (K == 0) return EF_OK;
    const int chunk_quads = chunk_envs / 4;
    const int grid_s = static_cast<int>(((n >> 2) + chunk_quads - 1) / chunk_quads);
    policy_demo_serve_kernel<<<grid_s, kServeProducerThreads, 0, static_cast<cudaStream_t>(stream)>>>(obs, action, obs_seq, act_seq, chunk_quads, K,
                                                                                         status, n);
    return cuda_status(cudaGetLastError());
}

extern "C" int ef_policy_demo_step(const int32_t* obs, int32_t* action, int32_t k, int64_t n, void* stream) {
    if (n <= 0 || !obs || !action || (n & 3) != 0 || !aligned(obs, 16) || !aligned(action, 16)) return EF_ERR_INVALID_ARGUMENT;
    policy_demo_step_kernel<<<grid_for(n >> 2, 8), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(obs, action, k, n >> 2);
    return cuda_status(cudaGetLastError());
}
