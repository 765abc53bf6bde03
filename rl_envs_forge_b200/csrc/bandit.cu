// envforge_b200 -- KArmedBandit batched step kernel (sm_100a, fp32 samplers over a Philox primitive stream).
//
// Replaces, for N independent bandits, the reference's per-instance
//   KArmedBandit.step  rl_envs_forge/envs/k_armed_bandit/k_armed_bandit.py:175-200
//   Arm.sample         rl_envs_forge/envs/k_armed_bandit/k_armed_bandit.py:66-122
// The reference samples through numpy's legacy RandomState (normal/uniform/exponential/gamma/beta/logistic/weibull/
// lognormal).  Those algorithms are restated here in fp32 over *primitives*: U(0,1) = u23(word) and N(0,1) = plain
// Box-Muller of two words, the words being consumed in order from the (env, step) substream kStreamBandit.  The
// oracle (oracle/bandit.py) feeds the same primitives to the float64 legacy algorithms, so device and oracle differ
// only by fp32 rounding -- except when an accept/reject comparison of a rejection sampler lands within that rounding
// (a "flip"), which the parity test counts.
#include <algorithm>
#include <cstdlib>

#include "ef_common.cuh"

namespace ef {

struct BanditArgs {
    const ef_arm* arms;  // [K, k]
    int32_t k;
    const int32_t* action;  // [K, n] or null
    const float* means;     // [n, k] default-arm means (ef_bandit_default_means) or null: re-derived in-kernel
    uint64_t seed, t0, env_offset;
    int32_t K;
    int32_t* obs;
    float* reward;
    double* stats;
    uint32_t* err;
    int64_t n;
    PhiloxKeys rk;  // round keys of `seed`
};

// Primitive stream of one (env, step).  The first Philox block and the transcendental-heavy values every family starts
// from (first normal, first uniform, first exponential) are computed up front by ALL lanes of the warp, outside the
// per-family switch: lanes of a warp usually hold different families, and anything inside the switch is serialised
// across them.  The values are exactly those the lazy path would produce (same words, same formulas, same order).
struct Primitives {
    uint64_t seed, t;
    uint32_t g, consumed;
    uint4 blk;
    float u0, u1, u2, n0, e0;
    __device__ __forceinline__ Primitives(uint64_t seed_, uint32_t g_, uint64_t t_) : seed(seed_), t(t_), g(g_), consumed(0) {
        blk = draw_block(seed, g, t, kStreamBandit, 0);
        u0 = u23(blk.x);
        u1 = u23(blk.y);
        u2 = u23(blk.z);
        n0 = sqrtf(-2.0f * logf(u0)) * cospif(2.0f * u1);
        e0 = -logf(1.0f - u0);
    }
    __device__ __forceinline__ uint32_t word() {
        const uint32_t lane = consumed & 3u;
        if (lane == 0u && consumed != 0u) blk = draw_block(seed, g, t, kStreamBandit, consumed >> 2);
        ++consumed;
        return lane == 0u ? blk.x : lane == 1u ? blk.y : lane == 2u ? blk.z : blk.w;
    }
    __device__ __forceinline__ float u() {
        if (consumed < 3u) {
            const float v = consumed == 0u ? u0 : (consumed == 1u ? u1 : u2);
            ++consumed;
            return v;
        }
        return u23(word());
    }
    __device__ __forceinline__ float n() {
        if (consumed == 0u) {
            consumed = 2u;
            return n0;
        }
        const float a = u();
        const float b = u();
        return sqrtf(-2.0f * logf(a)) * cospif(2.0f * b);
    }
    __device__ __forceinline__ float std_exp() {
        if (consumed == 0u) {
            consumed = 1u;
            return e0;
        }
        return -logf(1.0f - u());
    }
};

// The same primitive stream, evaluated on demand.  Used by the family-sorted kernel, where the lanes of a warp hold the
// same family and nothing needs hoisting; word for word and formula for formula identical to `Primitives`.
struct LazyPrimitives {
    const PhiloxKeys& rk;
    uint64_t t;
    uint32_t g, consumed;
    uint4 blk;
    __device__ __forceinline__ LazyPrimitives(const PhiloxKeys& rk_, uint32_t g_, uint64_t t_) : rk(rk_), t(t_), g(g_), consumed(0) {
        blk = draw_block_rk(rk, g, t, kStreamBandit, 0);
    }
    __device__ __forceinline__ uint32_t word() {
        const uint32_t lane = consumed & 3u;
        if (lane == 0u && consumed != 0u) blk = draw_block_rk(rk, g, t, kStreamBandit, consumed >> 2);
        ++consumed;
        return lane == 0u ? blk.x : lane == 1u ? blk.y : lane == 2u ? blk.z : blk.w;
    }
    __device__ __forceinline__ float u() { return u23(word()); }
    __device__ __forceinline__ float n() {
        const float a = u();
        const float b = u();
        return sqrtf(-2.0f * logf(a)) * cospif(2.0f * b);
    }
    __device__ __forceinline__ float std_exp() { return -logf(1.0f - u()); }
};

constexpr int kMaxRejections = 1024;  // termination guard; P(reaching it) is astronomically small for valid params

#ifndef EF_BANDIT_GAMMA_ATTR
#define EF_BANDIT_GAMMA_ATTR   // tuning knob: __noinline__ keeps ONE copy of the gamma sampler per kernel (code size)
#endif
template <typename P>
__device__ EF_BANDIT_GAMMA_ATTR float std_gamma(P& pr, float shape) {
    if (shape == 1.0f) return pr.std_exp();
    if (shape == 0.0f) return 0.0f;
    if (shape < 1.0f) {  // Ahrens-Dieter style (numpy legacy_standard_gamma, shape < 1)
        const float inv = 1.0f / shape;
        float x = 0.0f;
        for (int it = 0; it < kMaxRejections; ++it) {
            const float u = pr.u();
            const float v = pr.std_exp();
            if (u <= 1.0f - shape) {
                x = powf(u, inv);
                if (x <= v) return x;
            } else {
                const float y = -logf((1.0f - u) / shape);
                x = powf(1.0f - shape + shape * y, inv);
                if (x <= v + y) return x;
            }
        }
        return x;
    }
    // Marsaglia-Tsang (shape > 1)
    const float b = shape - 1.0f / 3.0f;
    const float c = 1.0f / sqrtf(9.0f * b);
    float v = 1.0f;
    for (int it = 0; it < kMaxRejections; ++it) {
        float x;
        int guard = 0;
        do {
            x = pr.n();
            v = 1.0f + c * x;
        } while (v <= 0.0f && ++guard < kMaxRejections);
        v = v * v * v;
        const float u = pr.u();
        const float x2 = x * x;
        if (u < 1.0f - 0.0331f * x2 * x2) return b * v;
        if (logf(u) < 0.5f * x2 + b * (1.0f - v + logf(v))) return b * v;
    }
    return b * v;
}

template <typename P>
__device__ float beta_sample(P& pr, float a, float b) {
    if (a <= 1.0f && b <= 1.0f) {  // Johnk
        const float ia = 1.0f / a, ib = 1.0f / b;
        float x = 0.0f, y = 1.0f;
        for (int it = 0; it < kMaxRejections; ++it) {
            const float u = pr.u();
            const float v = pr.u();
            x = powf(u, ia);
            y = powf(v, ib);
            const float s = x + y;
            if (s <= 1.0f) {  // u + v > 0 always holds: primitives are strictly positive
                if (s > 0.0f) return x / s;
                float lx = logf(u) * ia, ly = logf(v) * ib;
                const float lm = fmaxf(lx, ly);
                lx -= lm;
                ly -= lm;
                return expf(lx - logf(expf(lx) + expf(ly)));
            }
        }
        return x / (x + y);
    }
    const float ga = std_gamma(pr, a);
    const float gb = std_gamma(pr, b);
    return ga / (ga + gb);
}

__device__ __forceinline__ float default_mean(uint64_t seed, uint32_t g, uint32_t arm) {
    const uint4 w = draw_block(seed, g, arm, kStreamBanditMeans);
    return sqrtf(-2.0f * logf(u23(w.x))) * cospif(2.0f * u23(w.y));
}

template <typename P>
__device__ float sample_arm(const ef_arm arm, P& pr, uint64_t seed, uint32_t g, uint32_t arm_index) {
    switch (arm.family) {
        case EF_ARM_NORMAL: return arm.p0 + arm.p1 * pr.n();
        case EF_ARM_UNIFORM: return arm.p0 + (arm.p1 - arm.p0) * pr.u();
        case EF_ARM_EXPONENTIAL: return arm.p0 * pr.std_exp();
        case EF_ARM_GAMMA: return arm.p1 * std_gamma(pr, arm.p0);
        case EF_ARM_BETA: return beta_sample(pr, arm.p0, arm.p1);
        case EF_ARM_LOGISTIC: {
            const float u = pr.u();
            return arm.p0 + arm.p1 * logf(u / (1.0f - u));
        }
        case EF_ARM_WEIBULL: return arm.p0 == 0.0f ? 0.0f : powf(pr.std_exp(), 1.0f / arm.p0);
        case EF_ARM_LOGNORMAL: return expf(arm.p0 + arm.p1 * pr.n());
        default: return (default_mean(seed, g, arm_index) + arm.p0) + arm.p1 * pr.n();  // EF_ARM_DEFAULT_NORMAL
    }
}

__global__ void __launch_bounds__(kThreads) bandit_step_kernel(const BanditArgs a) {
    EpisodeAcc acc;
    ErrorAcc errs;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint32_t g = static_cast<uint32_t>(a.env_offset + i);
        for (int j = 0; j < a.K; ++j) {
            const uint64_t t = a.t0 + static_cast<uint64_t>(j);
            const int64_t o = static_cast<int64_t>(j) * a.n + i;
            int32_t act;
            if (a.action) {
                act = __ldg(a.action + o);
            } else {
                act = static_cast<int32_t>(__umulhi(group_word1(a.seed, g, t, kStreamPolicy), static_cast<uint32_t>(a.k)));
            }
            if (a.obs) a.obs[o] = act;  // observation = action (k_armed_bandit.py:200)
            if (static_cast<uint32_t>(act) >= static_cast<uint32_t>(a.k)) {
                errs.record(g, static_cast<uint32_t>(act), EF_STEP_BAD_ACTION);
                if (a.reward) a.reward[o] = 0.0f;
                continue;
            }
            const int4 raw = __ldg(reinterpret_cast<const int4*>(a.arms) + static_cast<int64_t>(j) * a.k + act);
            ef_arm arm;
            arm.family = raw.x;
            arm.p0 = __int_as_float(raw.y);
            arm.p1 = __int_as_float(raw.z);
            arm.reserved = 0;
            Primitives pr(a.seed, g, t);
            float r;
            if (a.means != nullptr && static_cast<uint32_t>(arm.family) >= 8u)   // EF_ARM_DEFAULT_NORMAL, mean looked up
                r = (__ldg(a.means + i * a.k + act) + arm.p0) + arm.p1 * pr.n();
            else
                r = sample_arm(arm, pr, a.seed, g, static_cast<uint32_t>(act));
            if (a.reward) a.reward[o] = r;
            acc.steps += 1u;
            acc.finish(1, r);  // every bandit step is a finished one-step episode (done is always True, :198)
        }
    }
    cta_epilogue(acc, errs, a.stats, a.err, nullptr, 0ull, 0ull);
}

// ---- family-sorted step kernel ---------------------------------------------------------------------------------------
// The eight samplers share almost no code, so a warp whose lanes hold different families executes them one after the
// other (measured on the plain kernel with 10 mixed arms: 11.2 of 32 lanes active per instruction).  A reward is a pure
// function of (global env index, step, action, arm parameters), so the order in which a CTA evaluates its envs is free:
// this kernel takes a tile of kTile consecutive envs, bins them by sampler code path in shared memory (warp-aggregated
// counting sort), evaluates them in sorted order -- the 32 lanes of a warp then run the same sampler -- and writes the
// rewards back through shared memory so global stores stay coalesced.  Results are bit-identical to the plain kernel.
//
// Second pass of round 2 (profiles/r02z_bandit_ab.txt):
//  * every bin starts on a 32-slot boundary of the sorted array (the gaps hold a sentinel), so a warp never holds two
//    samplers; with the bins packed back to back each of the ~12 bin boundaries of a tile cost an extra warp pass;
//  * tiles of 4096 envs (16 per thread; key / rank / action of an env wait in the reward staging array instead of in
//    registers), which halves the share of partially filled warps again;
//  * the 32-env groups of the sorted array are handed to the warps of the CTA dynamically (one shared counter): a
//    group of beta(2,5) draws costs 2.5 x a group of uniform draws;
//  * a default arm's per-env mean comes from the [n,k] table of ef_bandit_default_means when the caller passes it
//    (ef_bandit_step_means) instead of being re-derived with a second Philox call + Box-Muller on every step.  (An L1 or
//    L2 prefetch of the mean's line while the tile is binned was measured SLOWER: 5.13e10 vs 5.48e10 bandit-steps/s,
//    profiles/r02bh_bandit_mean_prefetch_ab.txt.)
#ifndef EF_BANDIT_ITEMS
#define EF_BANDIT_ITEMS 16
#endif
constexpr int kItemsPerThread = EF_BANDIT_ITEMS;
constexpr int kTile = kThreads * kItemsPerThread;  // 4096 envs per tile; local index and rank fit 12 bits
static_assert(kItemsPerThread % 8 == 0 && kTile <= 4096, "local index / rank are packed in 12 bits");
constexpr int kBins = 16;
constexpr int kSortedSlots = kTile + 32 * (kBins - 1);  // every bin padded to a multiple of 32
#ifndef EF_BANDIT_CTAS
#define EF_BANDIT_CTAS 4
#endif
#ifndef EF_BANDIT_DYNAMIC
#define EF_BANDIT_DYNAMIC 1
#endif
#ifndef EF_BANDIT_ALIGN
#define EF_BANDIT_ALIGN 32u   // A/B switch: 1u packs the bins back to back like the first version
#endif
constexpr int kSortedCtasPerSm = EF_BANDIT_CTAS;  // 64 registers x 256 threads: four resident CTAs per SM
constexpr uint32_t kBinSkip = 15u;                 // out-of-range env or rejected action: not evaluated
constexpr int32_t kSortedMaxArms = 1 << 16;        // the action waits in 16 bits next to rank and bin
constexpr uint32_t kEmptySlot = 0xffffffffu;

__device__ __forceinline__ ef_arm load_arm(const ef_arm* __restrict__ arms, int64_t index) {
    const int4 raw = __ldg(reinterpret_cast<const int4*>(arms) + index);
    ef_arm arm;
    arm.family = raw.x;
    arm.p0 = __int_as_float(raw.y);
    arm.p1 = __int_as_float(raw.z);
    arm.reserved = 0;
    return arm;
}

// Code-path bin of an arm: the family, with the samplers that branch on their parameters split further.
__device__ __forceinline__ uint32_t arm_bin(const ef_arm& arm) {
    const uint32_t f = static_cast<uint32_t>(arm.family);
    if (f == EF_ARM_GAMMA) return arm.p0 > 1.0f ? 3u : (arm.p0 == 1.0f ? 2u : (arm.p0 == 0.0f ? 1u : 9u));
    if (f == EF_ARM_BETA) {
        if (arm.p0 <= 1.0f && arm.p1 <= 1.0f) return 10u;                 // Johnk
        return (arm.p0 > 1.0f && arm.p1 > 1.0f) ? 4u : 11u;               // two Marsaglia-Tsang gammas / mixed
    }
    return f <= 8u ? f : 12u;
}

__global__ void __launch_bounds__(kThreads, kSortedCtasPerSm) bandit_step_sorted_kernel(const BanditArgs a) {
    __shared__ uint32_t s_cnt[kBins], s_base[kBins];
    __shared__ uint32_t s_next;
    __shared__ uint32_t s_sorted[kSortedSlots];   // action << 12 | local index, ordered by bin, bins 32-aligned
    __shared__ float s_rew[kTile];                // phase 1: action << 16 | rank << 4 | bin;  afterwards: rewards
    uint32_t* const s_tmp = reinterpret_cast<uint32_t*>(s_rew);
    EpisodeAcc acc;
    ErrorAcc errs;
    const uint32_t lane = threadIdx.x & 31u;
    const int64_t tiles = (a.n + kTile - 1) / kTile;
    const int64_t units = tiles * a.K;
    for (int64_t unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const int j = static_cast<int>(unit / tiles);
        const int64_t base_i = (unit - static_cast<int64_t>(j) * tiles) * kTile;
        const uint64_t t = a.t0 + static_cast<uint64_t>(j);
        const ef_arm* __restrict__ arms = a.arms + static_cast<int64_t>(j) * a.k;
        if (threadIdx.x < kBins) s_cnt[threadIdx.x] = 0u;
        if (threadIdx.x == 0) s_next = kThreads / 32;   // groups 0..7 are the warps' first ones
        __syncthreads();
        // phase 1: actions in (coalesced; eight loads of a thread issued before any is used), observations out,
        // bin + rank of every env of the tile
#pragma unroll
        for (int half = 0; half < kItemsPerThread; half += 8) {
            int32_t acts[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int64_t i = base_i + static_cast<uint32_t>(half + it) * kThreads + threadIdx.x;
                acts[it] = 0;
                if (i < a.n) {
                    const int64_t o = static_cast<int64_t>(j) * a.n + i;
                    acts[it] = a.action ? __ldg(a.action + o)
                                        : static_cast<int32_t>(__umulhi(group_word1(a.seed, static_cast<uint32_t>(a.env_offset + i), t,
                                                                                    kStreamPolicy), static_cast<uint32_t>(a.k)));
                }
            }
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const uint32_t l = static_cast<uint32_t>(half + it) * kThreads + threadIdx.x;
                const int64_t i = base_i + l;
                uint32_t key = kBinSkip;
                const int32_t act = acts[it];
                if (i < a.n) {
                    const uint32_t g = static_cast<uint32_t>(a.env_offset + i);
                    const int64_t o = static_cast<int64_t>(j) * a.n + i;
                    if (a.obs) a.obs[o] = act;  // observation = action (k_armed_bandit.py:200)
                    if (static_cast<uint32_t>(act) >= static_cast<uint32_t>(a.k)) {
                        errs.record(g, static_cast<uint32_t>(act), EF_STEP_BAD_ACTION);
                    } else {
                        key = arm_bin(load_arm(arms, act));
                    }
                }
                // warp-aggregated counting: lanes with the same bin elect a leader that reserves their slots at once
                const uint32_t peers = __match_any_sync(0xffffffffu, key);
                const uint32_t leader = static_cast<uint32_t>(__ffs(peers) - 1);
                uint32_t first = 0u;
                if (lane == leader && key != kBinSkip) first = atomicAdd(&s_cnt[key], static_cast<uint32_t>(__popc(peers)));
                first = __shfl_sync(0xffffffffu, first, leader);
                const uint32_t rank = first + static_cast<uint32_t>(__popc(peers & ((1u << lane) - 1u)));
                s_tmp[l] = (static_cast<uint32_t>(act) << 16) | ((rank & 0xfffu) << 4) | key;
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {  // exclusive prefix over the padded bin sizes (one warp), sentinels into the gaps
            const uint32_t c = lane < kBins ? s_cnt[lane] : 0u;
            const uint32_t cp = (lane < kBinSkip) ? ((c + (EF_BANDIT_ALIGN - 1u)) & ~(EF_BANDIT_ALIGN - 1u)) : 0u;
            uint32_t incl = cp;
#pragma unroll
            for (int o = 1; o < kBins; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= static_cast<uint32_t>(o)) incl += v;
            }
            if (lane < kBins) s_base[lane] = incl - cp;   // s_base[kBinSkip] = all padded bins together
            if (lane < kBinSkip)
                for (uint32_t q = incl - cp + c; q < incl; ++q) s_sorted[q] = kEmptySlot;
        }
        __syncthreads();
        const uint32_t total = s_base[kBinSkip];  // slots to visit = everything in the bins below the skip bin
#pragma unroll
        for (int it = 0; it < kItemsPerThread; ++it) {
            const uint32_t l = static_cast<uint32_t>(it) * kThreads + threadIdx.x;
            const uint32_t w = s_tmp[l];
            const uint32_t key = w & 15u;
            if (key != kBinSkip)
                s_sorted[s_base[key] + ((w >> 4) & 0xfffu)] = ((w >> 16) << 12) | l;
            else
                s_rew[l] = 0.0f;   // rejected action (or past the end of the batch): reward 0, not evaluated
        }
        __syncthreads();
        // phase 2: evaluate in sorted order -- the 32 lanes of a warp run the same sampler
#if EF_BANDIT_DYNAMIC
        for (uint32_t grp = threadIdx.x >> 5; grp * 32u < total;) {
            const uint32_t rec = s_sorted[grp * 32u + lane];
#else
        for (uint32_t p = threadIdx.x; p < total; p += kThreads) {
            const uint32_t rec = s_sorted[p];
#endif
            if (rec != kEmptySlot) {
                const uint32_t l = rec & 0xfffu, act = rec >> 12;
                const uint32_t g = static_cast<uint32_t>(a.env_offset + base_i + l);
                LazyPrimitives pr(a.rk, g, t);
                const ef_arm arm = load_arm(arms, act);
                float r;
                if (a.means != nullptr && static_cast<uint32_t>(arm.family) >= 8u)   // EF_ARM_DEFAULT_NORMAL, mean looked up
                    r = (__ldg(a.means + (base_i + l) * a.k + act) + arm.p0) + arm.p1 * pr.n();
                else
                    r = sample_arm(arm, pr, a.seed, g, act);
                s_rew[l] = r;
                acc.steps += 1u;
                acc.finish(1, r);  // every bandit step is a finished one-step episode (done is always True, :198)
            }
#if EF_BANDIT_DYNAMIC
            if (lane == 0) grp = atomicAdd(&s_next, 1u);
            grp = __shfl_sync(0xffffffffu, grp, 0);
#endif
        }
        __syncthreads();
        // phase 3: rewards out (coalesced)
        if (a.reward) {
#pragma unroll
            for (int it = 0; it < kItemsPerThread; ++it) {
                const uint32_t l = static_cast<uint32_t>(it) * kThreads + threadIdx.x;
                const int64_t i = base_i + l;
                if (i < a.n) a.reward[static_cast<int64_t>(j) * a.n + i] = s_rew[l];
            }
        }
        __syncthreads();
    }
    cta_epilogue(acc, errs, a.stats, a.err, nullptr, 0ull, 0ull);
}

__global__ void __launch_bounds__(kThreads) bandit_means_kernel(int32_t k, uint64_t seed, uint64_t env_offset,
                                                                float* __restrict__ means, int64_t n) {
    const int64_t total = n * k;
    for (int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < total;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t i = j / k;
        means[j] = default_mean(seed, static_cast<uint32_t>(env_offset + i), static_cast<uint32_t>(j - i * k));
    }
}

}  // namespace ef

using namespace ef;

extern "C" int ef_bandit_step_means(const ef_arm* arms, int32_t k, const int32_t* action, const float* means, uint64_t seed,
                                    uint64_t t0, uint64_t env_offset, int32_t K, int32_t* obs, float* reward, double* stats,
                                    uint32_t* err, int64_t n, void* stream) {
    if (!arms || k <= 0 || K < 0 || n < 0 || !aligned(arms, 16) || !aligned(means, 4)) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0 || K == 0) return EF_OK;
    BanditArgs a{};
    a.arms = arms; a.k = k; a.action = action; a.means = means; a.seed = seed; a.t0 = t0; a.env_offset = env_offset; a.K = K;
    a.obs = obs; a.reward = reward; a.stats = stats; a.err = err; a.n = n;
    a.rk = philox_round_keys(seed);
    const bool unsorted = getenv("ENVFORGE_BANDIT_UNSORTED") != nullptr;  // A/B + parity switch: the plain kernel
    if (k <= kSortedMaxArms && !unsorted) {
        const int64_t units = (n + kTile - 1) / kTile * K;
        const int grid = static_cast<int>(std::min<int64_t>(units, static_cast<int64_t>(sm_count()) * kSortedCtasPerSm));
        bandit_step_sorted_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
    } else {
        bandit_step_kernel<<<grid_for(n, 8), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
    }
    return cuda_status(cudaGetLastError());
}

extern "C" int ef_bandit_step(const ef_arm* arms, int32_t k, const int32_t* action, uint64_t seed, uint64_t t0,
                              uint64_t env_offset, int32_t K, int32_t* obs, float* reward, double* stats,
                              uint32_t* err, int64_t n, void* stream) {
    return ef_bandit_step_means(arms, k, action, nullptr, seed, t0, env_offset, K, obs, reward, stats, err, n, stream);
}

extern "C" int ef_bandit_default_means(int32_t k, uint64_t seed, uint64_t env_offset, float* means, int64_t n,
                                       void* stream) {
    if (k <= 0 || n < 0 || !means) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    bandit_means_kernel<<<grid_for(n * k, 8), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(k, seed, env_offset, means, n);
    return cuda_status(cudaGetLastError());
}
