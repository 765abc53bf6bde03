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

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

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

// ---- Philox4x32-10 ------------------------------------------------------------------------------------------------
// Stream layout (twin: oracle/philox.py).  key = (seed lo, seed hi); ctr = (g, c_lo, c_hi, stream | block << 8).
constexpr uint32_t kStreamStep = 0;        // per (env, t>>1): words (0,1) even t, (2,3) odd t = [transition, policy]
constexpr uint32_t kStreamResetAuto = 1;   // per (env, t): initial state of the episode starting after step t
constexpr uint32_t kStreamResetUser = 2;   // per (env, reset epoch)
constexpr uint32_t kStreamBandit = 3;      // per (env, step), block j: sampler primitives
constexpr uint32_t kStreamBanditMeans = 4; // per (env, arm): default-arm mean

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

__device__ __forceinline__ uint4 draw_block(uint64_t seed, uint32_t g, uint64_t counter, uint32_t stream,
                                            uint32_t block = 0) {
    return philox4x32_10(g, static_cast<uint32_t>(counter), static_cast<uint32_t>(counter >> 32),
                         stream | (block << 8), static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
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

// Block-level fold of the per-thread accumulators into `stats` (one atomicAdd set per CTA).  All threads of the CTA
// must call it (it synchronises).  `stats` may be null.
__device__ __forceinline__ void flush_stats(const EpisodeAcc& acc, double* __restrict__ stats) {
    if (stats == nullptr) return;
    __shared__ double s_part[5];
    if (threadIdx.x < 5) s_part[threadIdx.x] = 0.0;
    __syncthreads();
    double v[5] = {static_cast<double>(acc.episodes), static_cast<double>(acc.length), acc.ret, acc.ret_sq,
                   static_cast<double>(acc.steps)};
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const double w = warp_sum(v[j]);
        if ((threadIdx.x & 31) == 0 && w != 0.0) atomicAdd(&s_part[j], w);
    }
    __syncthreads();
    if (threadIdx.x < 5 && s_part[threadIdx.x] != 0.0) atomicAdd(&stats[threadIdx.x], s_part[threadIdx.x]);
}

// Device-resident step clock (optional, for CUDA-graph capture where the host cannot bump `t` between replays):
// clock[0] = steps already taken, clock[1] = CTA ticket.  Every CTA reads clock[0] when it starts; the CTA that
// finishes last (ticket == gridDim.x - 1) advances clock[0] by `inc`, i.e. only after all CTAs have read it.
__device__ __forceinline__ uint64_t clock_read(const unsigned long long* clock) {
    return clock ? *reinterpret_cast<const volatile unsigned long long*>(clock) : 0ull;
}
__device__ __forceinline__ void clock_advance(unsigned long long* clock, unsigned long long inc) {
    if (clock == nullptr) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&clock[1], 1ull) == static_cast<unsigned long long>(gridDim.x) - 1ull) {
            clock[1] = 0ull;
            clock[0] += inc;
            __threadfence();
        }
    }
}

// Rejected env-steps ("the reference would have raised here") are counted per thread and folded once per CTA, so a
// batch with many offenders costs one global atomic per CTA instead of one per offender.
struct ErrorAcc {
    uint32_t count = 0, g = 0, detail = 0, kind = 0;
    __device__ __forceinline__ void record(uint32_t g_, uint32_t detail_, uint32_t kind_) {
        ++count;
        g = g_;
        detail = detail_;
        kind = kind_;
    }
};

// All threads of the CTA must call it.  err[0] += offenders of this CTA; the CTA that finds err[0] == 0 also records
// one offender's (global env index + 1, detail, kind).
__device__ __forceinline__ void flush_errors(const ErrorAcc& e, uint32_t* __restrict__ err) {
    if (err == nullptr) return;
    __shared__ uint32_t s_err[4];
    if (threadIdx.x < 4) s_err[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t mask = __ballot_sync(0xffffffffu, e.count != 0u);
    if (mask != 0u) {
        uint32_t c = e.count;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
        if ((threadIdx.x & 31u) == 0u) atomicAdd(&s_err[0], c);
        if ((threadIdx.x & 31u) == static_cast<uint32_t>(__ffs(mask) - 1)) {  // lowest offending lane of the warp
            s_err[1] = e.g + 1u;
            s_err[2] = e.detail;
            s_err[3] = e.kind;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_err[0] != 0u) {
        if (atomicAdd(&err[0], s_err[0]) == 0u) {
            err[1] = s_err[1];
            err[2] = s_err[2];
            err[3] = s_err[3];
        }
    }
}

}  // namespace ef
