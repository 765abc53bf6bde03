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

__device__ __forceinline__ void report_error(uint32_t* __restrict__ err, uint32_t g, uint32_t detail, uint32_t kind) {
    if (err == nullptr) return;
    if (atomicAdd(&err[0], 1u) == 0u) {
        err[1] = g + 1u;
        err[2] = detail;
        err[3] = kind;
    }
}

}  // namespace ef
