// envforge_b200 -- shared device/host helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/envforge_b200.h"

namespace ef {

// ---- host-side status plumbing ----------------------------------------------------------------------------------
extern thread_local cudaError_t g_last_cuda_error;

inline int cuda_status(cudaError_t e) {
    if (e == cudaSuccess) return EF_OK;
    g_last_cuda_error = e;
    return EF_ERR_CUDA;
}

__host__ __device__ inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

// Launch geometry: B200 has 148 SMs; a grid of `per_sm` resident CTAs per SM looping grid-stride keeps every SM
// busy without a tail wave.  Queried once per process per device (cheap attribute reads, cached in a small table).
int sm_count();

constexpr int kThreads = 256;

inline int grid_for(int64_t work_items, int per_sm) {
    int64_t blocks = (work_items + kThreads - 1) / kThreads;
    int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// Step kernels are launched with the programmatic-stream-serialization attribute: the CTAs of launch i+1 may be
// scheduled while launch i drains, run their prologue (stage the transition table into shared memory) and then block in
// pdl_wait() until launch i has completed and its writes are visible.  That hides launch latency and the prologue behind
// the previous step's tail.  Both instructions are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();  // false when ENVFORGE_NO_PDL is set (A/B experiments)

template <typename Kernel, typename Args>
inline int launch_kernel(Kernel kernel, const Args& args, int grid, size_t smem_bytes, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cuda_status(cudaLaunchKernelEx(&cfg, kernel, args));
}

// ---- cp.async (LDGSTS): 16-byte global -> shared copies that hold no registers while in flight -------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- TMA bulk copy (cp.async.bulk, UBLKCP) + mbarrier: one thread moves a whole table into shared memory ----------------
// The copy is asynchronous: the issuing thread arms an mbarrier with the byte count, the copy engine completes it, and
// any thread that then observes the phase flip may read the data.  Source, destination and size must be 16-byte multiples.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// L2 prefetch hints.  A prefetch into L2 never breaks coherence (L2 is the point of coherence: a line that a still-running
// producer writes afterwards is simply updated there), so they may be issued AHEAD of griddepcontrol.wait.
__device__ __forceinline__ void l2_prefetch_bulk(const void* gmem_src, uint32_t bytes) {   // 16-byte aligned, multiple of 16
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void l2_prefetch_line(const void* gmem_src) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(gmem_src));
}

// The same in two steps, for several copies completing on one barrier: arm it with the total, then issue the copies.
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Shared -> global bulk store (bulk-group completion).  The shared-memory bytes were written through the generic proxy:
// the writers fence (fence_proxy_async) and synchronise with the issuing thread before it calls bulk_s2g.
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// ---- Philox4x32-10 ------------------------------------------------------------------------------------------------
// Stream layout (twin: oracle/philox.py).  key = (seed lo, seed hi); ctr = (x, c_lo, c_hi, stream | block << 8).
//   per-env streams: x = g (global env index).  Grouped streams: x = g >> 2 and env g owns word g & 3, so the one
//   Philox call a thread makes serves the four consecutive envs it processes.
constexpr uint32_t kStreamTransition = 0;  // grouped, per (g>>2, t): outcome draw of env g at step t
constexpr uint32_t kStreamResetAuto = 1;   // per (env, t): initial state of the episode starting after step t
constexpr uint32_t kStreamResetUser = 2;   // per (env, reset epoch)
constexpr uint32_t kStreamBandit = 3;      // per (env, step), block j: sampler primitives
constexpr uint32_t kStreamBanditMeans = 4; // per (env, arm): default-arm mean
constexpr uint32_t kStreamPolicy = 5;      // grouped, per (g>>2, t): random-policy draw of env g at step t

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = static_cast<uint64_t>(M0) * c0;
        const uint64_t p1 = static_cast<uint64_t>(M1) * c2;
        const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
        c1 = static_cast<uint32_t>(p1);
        c3 = static_cast<uint32_t>(p0);
        c0 = n0;
        c2 = n2;
        k0 += W0;
        k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

// The ten round keys of a launch are the same for every thread (key = seed): the host derives them once and passes them
// in the kernel parameters, so the XORs take them straight from the constant bank instead of every Philox call
// re-running the key schedule on the uniform datapath (20 UIADD3 per call, 4 % of the GridWorld step kernel).
struct PhiloxKeys {
    uint32_t k[20];  // k[2r] = key0 + r * W0, k[2r + 1] = key1 + r * W1
};

inline PhiloxKeys philox_round_keys(uint64_t seed) {
    PhiloxKeys rk;
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        rk.k[2 * r] = k0;
        rk.k[2 * r + 1] = k1;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return rk;
}

__device__ __forceinline__ uint4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& rk) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = static_cast<uint64_t>(M0) * c0;
        const uint64_t p1 = static_cast<uint64_t>(M1) * c2;
        const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ rk.k[2 * r];
        const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ rk.k[2 * r + 1];
        c1 = static_cast<uint32_t>(p1);
        c3 = static_cast<uint32_t>(p0);
        c0 = n0;
        c2 = n2;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint4 draw_block(uint64_t seed, uint32_t g, uint64_t counter, uint32_t stream,
                                            uint32_t block = 0) {
    return philox4x32_10(g, static_cast<uint32_t>(counter), static_cast<uint32_t>(counter >> 32),
                         stream | (block << 8), static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
}

// draw_block with host-derived round keys (kernel-parameter constant bank).
__device__ __forceinline__ uint4 draw_block_rk(const PhiloxKeys& rk, uint32_t g, uint64_t counter, uint32_t stream,
                                               uint32_t block = 0) {
    return philox4x32_10_rk(g, static_cast<uint32_t>(counter), static_cast<uint32_t>(counter >> 32), stream | (block << 8), rk);
}

// Grouped-stream words of the four consecutive envs g0 .. g0+3 at counter t.  g0 & 3 is the same for every thread of a
// launch (it equals env_offset & 3), so the second call is a launch-uniform branch taken only for unaligned shards.
// Unaligned shard (env_offset & 3 != 0): the quad straddles two Philox blocks.  Kept out of line: it is launch-uniformly
// rare and would otherwise double the code size of every unrolled call site.
static __device__ __noinline__ uint4 group_words4_unaligned(uint64_t seed, uint32_t g0, uint64_t t, uint32_t stream) {
    const uint4 a = draw_block(seed, g0 >> 2, t, stream);
    const uint4 b = draw_block(seed, (g0 >> 2) + 1u, t, stream);
    const uint32_t sh = g0 & 3u;
    const uint32_t all[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = sh == 1u ? all[j + 1] : (sh == 2u ? all[j + 2] : all[j + 3]);
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ void group_words4(uint64_t seed, uint32_t g0, uint64_t t, uint32_t stream, uint32_t (&w)[4]) {
    uint4 a;
    if ((g0 & 3u) == 0u) a = draw_block(seed, g0 >> 2, t, stream);
    else a = group_words4_unaligned(seed, g0, t, stream);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
}

// group_words4 with host-derived round keys (`seed` only feeds the rare unaligned fallback).
__device__ __forceinline__ void group_words4_rk(const PhiloxKeys& rk, uint64_t seed, uint32_t g0, uint64_t t, uint32_t stream,
                                                uint32_t (&w)[4]) {
    uint4 a;
    if ((g0 & 3u) == 0u)
        a = philox4x32_10_rk(g0 >> 2, static_cast<uint32_t>(t), static_cast<uint32_t>(t >> 32), stream, rk);
    else
        a = group_words4_unaligned(seed, g0, t, stream);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
}

__device__ __forceinline__ uint32_t group_word1(uint64_t seed, uint32_t g, uint64_t t, uint32_t stream) {
    const uint4 a = draw_block(seed, g >> 2, t, stream);
    const uint32_t l = g & 3u;
    return l == 0u ? a.x : (l == 1u ? a.y : (l == 2u ? a.z : a.w));
}

// uint32 word -> uniform in (0,1), ((w>>9)+0.5)*2^-23: exact in fp32, and so is 1-u.
__device__ __forceinline__ float u23(uint32_t w) { return (static_cast<float>(w >> 9) + 0.5f) * 1.1920928955078125e-07f; }

// lo + (hi-lo)*u with separate roundings (no FMA contraction) so the numpy fp32 twin is bit-identical.
__device__ __forceinline__ float uniform_f32(uint32_t w, float lo, float span) {
    return __fadd_rn(__fmul_rn(u23(w), span), lo);
}

// ---- episode statistics -------------------------------------------------------------------------------------------
struct EpisodeAcc {
    uint32_t episodes = 0, steps = 0;  // per-thread counts (a thread sees < 2^32 env-steps per launch)
    unsigned long long length = 0;
    double ret = 0.0, ret_sq = 0.0;
    __device__ __forceinline__ void finish(int len, float r) {
        episodes += 1u;
        length += static_cast<unsigned long long>(len);
        ret += static_cast<double>(r);
        ret_sq += static_cast<double>(r) * static_cast<double>(r);
    }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Rejected env-steps ("the reference would have raised here") are counted per thread and folded once per CTA.
struct ErrorAcc {
    uint32_t count = 0, g = 0, detail = 0, kind = 0;
    __device__ __forceinline__ void record(uint32_t g_, uint32_t detail_, uint32_t kind_) {
        ++count;
        g = g_;
        detail = detail_;
        kind = kind_;
    }
};

// Device-resident step clock (optional, for CUDA-graph capture where the host cannot bump `t` between replays).
// clock[0] is a counter in units of 2^-20 steps: a launch that advances the envs by `inc` steps adds inc * 2^20 to it in
// total -- every CTA adds its share (total / gridDim.x, CTA 0 also the remainder) with ONE fire-and-forget reduction when
// it retires -- and any thread reads the step index as clock[0] >> 20.  While a launch is running the counter lies in
// [t * 2^20, (t + inc) * 2^20) as long as at least one CTA (the reader itself) has not retired, so every thread of the
// launch reads the same t whenever it looks: no ticket, no "last CTA" hand-over, no barrier, and the read is a plain
// L2 hit that is issued ahead of the input loads.  (The previous scheme -- a ticket atomic next to the clock word, the
// last ticket holder bumping it -- put a contended atomic round trip on every CTA's critical path: profiles/r02_*.)
constexpr unsigned kClockShift = 20;
__device__ __forceinline__ unsigned long long clock_fetch(const unsigned long long* clock) {
    unsigned long long v = 0ull;
    if (clock != nullptr) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(clock) : "memory");
    return v;
}
// Step index for every calling thread (`clock` is launch-uniform; null = the host passes `t` itself).
__device__ __forceinline__ uint64_t clock_steps(unsigned long long fetched) { return fetched >> kClockShift; }
// fetch + convert in one go for the kernels that have nothing to overlap the fetch with.  `ticket` is kept for the
// call sites' sake (cta_epilogue ignores it).
__device__ __forceinline__ uint64_t clock_begin(unsigned long long* clock, unsigned long long& ticket, bool = true) {
    ticket = 0ull;
    return clock_steps(clock_fetch(clock));
}

constexpr int kWarpsPerCta = kThreads / 32;

// CTA epilogue: fold the per-thread statistics / error accumulators into global memory with ONE barrier and at most
// a handful of global reductions per CTA, then add this CTA's share to the device clock.
// All threads of the CTA must call it.  `stats`, `err`, `clock` may each be null.  WARPS = warps per CTA (8 everywhere but in
// the 512-thread step server).
template <int WARPS = kWarpsPerCta>
__device__ __forceinline__ void cta_epilogue(const EpisodeAcc& acc, const ErrorAcc& errs, double* __restrict__ stats,
                                             uint32_t* __restrict__ err, unsigned long long* clock,
                                             unsigned long long ticket, unsigned long long clock_inc) {
    __shared__ uint32_t s_u[WARPS][3];            // steps, episodes, rejected
    __shared__ unsigned long long s_len[WARPS];
    __shared__ double s_d[WARPS][2];              // return, return^2
    __shared__ uint32_t s_e[WARPS][3];            // one offender per warp: g + 1, detail, kind
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t steps = __reduce_add_sync(0xffffffffu, acc.steps);
    const uint32_t eps = __reduce_add_sync(0xffffffffu, acc.episodes);
    const uint32_t rejected = __reduce_add_sync(0xffffffffu, errs.count);
    unsigned long long len = 0ull;
    double r = 0.0, r2 = 0.0;
    if (eps) {  // warp-uniform: floating sums only where some lane finished an episode
        len = acc.length;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) len += __shfl_down_sync(0xffffffffu, len, o);
        r = warp_sum(acc.ret);
        r2 = warp_sum(acc.ret_sq);
    }
    if (rejected) {
        const uint32_t mask = __ballot_sync(0xffffffffu, errs.count != 0u);
        if (lane == static_cast<uint32_t>(__ffs(mask) - 1)) {
            s_e[warp][0] = errs.g + 1u;
            s_e[warp][1] = errs.detail;
            s_e[warp][2] = errs.kind;
        }
    }
    if (lane == 0u) {
        s_u[warp][0] = steps;
        s_u[warp][1] = eps;
        s_u[warp][2] = rejected;
        s_len[warp] = len;
        s_d[warp][0] = r;
        s_d[warp][1] = r2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t_steps = 0ull, t_eps = 0ull, t_len = 0ull;
        uint32_t t_rej = 0u, first = WARPS;
        double t_r = 0.0, t_r2 = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            t_steps += s_u[w][0];
            t_eps += s_u[w][1];
            if (s_u[w][2] != 0u && first == WARPS) first = w;
            t_rej += s_u[w][2];
            t_len += s_len[w];
            t_r += s_d[w][0];
            t_r2 += s_d[w][1];
        }
        if (stats != nullptr) {
            // CTAs finish together and would serialise on one L2 atomic unit if they all added into the same five
            // addresses (measured: 1.5 us per launch at 296 CTAs): spread them over EF_STATS_REPLICAS copies of the vector
            double* mine = stats + static_cast<size_t>(blockIdx.x % EF_STATS_REPLICAS) * EF_STATS_STRIDE;
            if (t_eps) {
                atomicAdd(&mine[0], static_cast<double>(t_eps));
                atomicAdd(&mine[1], static_cast<double>(t_len));
                atomicAdd(&mine[2], t_r);
                atomicAdd(&mine[3], t_r2);
            }
            if (t_steps) atomicAdd(&mine[4], static_cast<double>(t_steps));
        }
        if (err != nullptr && t_rej != 0u) {  // no returning atomics: count by RED.ADD, offender by RED.MAX on a packed word
            atomicAdd(&err[0], t_rej);
            const unsigned long long packed = (static_cast<unsigned long long>(s_e[first][0]) << 32) |
                                              (static_cast<unsigned long long>(s_e[first][2] & 0xffu) << 24) |
                                              static_cast<unsigned long long>(s_e[first][1] & 0xffffffu);
            atomicMax(reinterpret_cast<unsigned long long*>(err + 2), packed);
        }
        if (clock != nullptr && clock_inc != 0ull) {   // this CTA's share of clock_inc * 2^20 (see "step clock" above)
            const unsigned long long total = clock_inc << kClockShift;
            const unsigned long long share = total / gridDim.x;
            atomicAdd(clock, blockIdx.x == 0 ? share + (total - share * gridDim.x) : share);
        }
        (void)ticket;
    }
}

}  // namespace ef
