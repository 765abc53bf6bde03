// envforge_b200 -- KArmedBandit stream-exact step (sm_100a): every bandit owns the generator the reference would build
// for it, numpy's legacy RandomState (MT19937), and samples with numpy's legacy algorithms in float64.
//
// Replaces, for N independent instances and WITHOUT injected draws,
//   KArmedBandit.__init__ / reset   rl_envs_forge/envs/k_armed_bandit/k_armed_bandit.py:145, :215  (RandomState(seed))
//   Arm.sample                      rl_envs_forge/envs/k_armed_bandit/k_armed_bandit.py:66-122
//   KArmedBandit.step               rl_envs_forge/envs/k_armed_bandit/k_armed_bandit.py:175-200
// so that `VecKArmedBandit(n, seed=s, rng="mt19937")` env i reproduces the reward stream of `KArmedBandit(seed=s+i)`
// (the bandit half: 2.5 KB of generator state per env makes this the small-N mode; the Philox
// kernels in bandit.cu stay the throughput path).
//
// The sampling arithmetic lives in a third-party dependency (numpy 1.26.4, legacy-distributions.c / mt19937.c; not under
// /root/reference).  Restated here from the published algorithms, the same restatement oracle/bandit.py pins bit-exactly
// against numpy.random.RandomState:
//   mt19937_next / mt19937_gen        32-bit Mersenne Twister with numpy's tempering
//   next_double                       (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53
//   legacy_gauss                      polar Box-Muller with a one-value cache that PERSISTS across calls (per-env state)
//   standard_exponential              -log(1 - U)
//   standard_gamma                    shape < 1: Ahrens-Dieter rejection;  shape > 1: Marsaglia-Tsang
//   beta                              a, b <= 1: Johnk;  otherwise two gammas
//   logistic / weibull / lognormal / normal / uniform: inverse-cdf / affine forms
// Host state layout: key [624, n] uint32 (word-major: the lanes of a warp read consecutive addresses while their
// generators advance in lock step), pos [n] int32, has_gauss [n] int32, gauss [n] double.  The host fills them from
// numpy's own RandomState(seed).get_state() after the default means were drawn (randn(k), :162), so seeding
// (init_genrand / init_by_array) is numpy's, not a restatement.
//
// float64 throughout: accept/reject comparisons and the number of words consumed must follow numpy, else a stream
// desynchronises for good.  CUDA's double log/exp/pow/sqrt are within 1-2 ulp of glibc's, so rewards agree with the
// reference to ~1e-15 relative before the fp32 cast of the public reward tensor (tests: rtol 1e-6 on fp32, 1e-12 on the
// optional float64 output, zero desynchronised streams).
#include <algorithm>

#include "ef_common.cuh"

namespace ef {

struct BanditMtArgs {
    const ef_arm64* arms;  // [k] parameters in force for this step
    int32_t k;
    const int32_t* action;  // [n]
    uint32_t* key;          // [624, n]
    int32_t* pos;           // [n]
    int32_t* has_gauss;     // [n]
    double* gauss;          // [n]
    const double* means;    // [n, k] per-env default-arm means (randn(k) of that env's generator), nullable
    int32_t* obs;
    float* reward;
    double* reward64;
    double* stats;
    uint32_t* err;
    uint64_t env_offset;
    int64_t n;
};

constexpr int kMtN = 624, kMtM = 397;

struct Mt19937 {
    uint32_t* key;  // this env's column: word j at key[j * n]
    int64_t n;
    int pos;
    bool has_gauss;
    double gauss;

    __device__ __forceinline__ uint32_t& w(int j) { return key[static_cast<int64_t>(j) * n]; }

    __device__ void regenerate() {  // mt19937_gen
        constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;
        int kk = 0;
        uint32_t cur = w(0), first = cur;
        for (; kk < kMtN - kMtM; ++kk) {
            const uint32_t nxt = w(kk + 1);
            const uint32_t y = (cur & kUpper) | (nxt & kLower);
            w(kk) = w(kk + kMtM) ^ (y >> 1) ^ ((y & 1u) ? kMatrixA : 0u);
            cur = nxt;
        }
        first = w(0);  // already regenerated: the wrap-around terms below read NEW words, like the reference loop
        for (; kk < kMtN - 1; ++kk) {
            const uint32_t nxt = w(kk + 1);
            const uint32_t y = (cur & kUpper) | (nxt & kLower);
            w(kk) = w(kk + (kMtM - kMtN)) ^ (y >> 1) ^ ((y & 1u) ? kMatrixA : 0u);
            cur = nxt;
        }
        const uint32_t y = (cur & kUpper) | (first & kLower);
        w(kMtN - 1) = w(kMtM - 1) ^ (y >> 1) ^ ((y & 1u) ? kMatrixA : 0u);
        pos = 0;
    }
    __device__ __forceinline__ uint32_t next_u32() {
        if (pos == kMtN) regenerate();
        uint32_t y = w(pos++);
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    __device__ __forceinline__ double u() {  // mt19937_next_double
        const uint32_t a = next_u32() >> 5, b = next_u32() >> 6;
        return (static_cast<double>(a) * 67108864.0 + static_cast<double>(b)) / 9007199254740992.0;
    }
    __device__ double n01() {  // legacy_gauss
        if (has_gauss) {
            const double t = gauss;
            has_gauss = false;
            gauss = 0.0;
            return t;
        }
        double x1, x2, r2;
        do {
            x1 = 2.0 * u() - 1.0;
            x2 = 2.0 * u() - 1.0;
            r2 = x1 * x1 + x2 * x2;
        } while (r2 >= 1.0 || r2 == 0.0);
        const double f = sqrt(-2.0 * log(r2) / r2);
        gauss = f * x1;
        has_gauss = true;
        return f * x2;
    }
    __device__ __forceinline__ double std_exp() { return -log(1.0 - u()); }
};

// Rejection loops are capped: with finite parameters a loop accepts within a handful of trips (2^16 trips are out of
// reach by hundreds of orders of magnitude), but a NaN shape handed in through the C ABI would never be accepted and
// must not spin the GPU forever.  The capped exit returns NaN, which is also what numpy returns for NaN parameters.
constexpr int kMtMaxTrips = 1 << 16;

__device__ double mt_std_gamma(Mt19937& r, double shape) {  // legacy_standard_gamma
    if (shape == 1.0) return r.std_exp();
    if (shape == 0.0) return 0.0;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    if (shape < 1.0) {
        for (int trip = 0; trip < kMtMaxTrips; ++trip) {
            const double u = r.u();
            const double v = r.std_exp();
            if (u <= 1.0 - shape) {
                const double x = pow(u, 1.0 / shape);
                if (x <= v) return x;
            } else {
                const double y = -log((1.0 - u) / shape);
                const double x = pow(1.0 - shape + shape * y, 1.0 / shape);
                if (x <= v + y) return x;
            }
        }
        return nan;
    }
    const double b = shape - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * b);
    for (int trip = 0; trip < kMtMaxTrips; ++trip) {
        double x, v;
        do {
            x = r.n01();
            v = 1.0 + c * x;
        } while (v <= 0.0);
        v = v * v * v;
        const double u = r.u();
        if (u < 1.0 - 0.0331 * (x * x) * (x * x)) return b * v;
        if (log(u) < 0.5 * x * x + b * (1.0 - v + log(v))) return b * v;
    }
    return nan;
}

__device__ double mt_beta(Mt19937& r, double a, double b) {  // legacy_beta
    if (a <= 1.0 && b <= 1.0) {
        for (int trip = 0; trip <= kMtMaxTrips; ++trip) {  // Johnk
            if (trip == kMtMaxTrips) return __longlong_as_double(0x7ff8000000000000ll);
            const double u = r.u(), v = r.u();
            const double x = pow(u, 1.0 / a), y = pow(v, 1.0 / b);
            const double xy = x + y;
            if (xy <= 1.0 && u + v > 0.0) {
                if (xy > 0.0) return x / xy;
                double lx = log(u) / a, ly = log(v) / b;
                const double lm = lx > ly ? lx : ly;
                lx -= lm;
                ly -= lm;
                return exp(lx - log(exp(lx) + exp(ly)));
            }
        }
    }
    const double ga = mt_std_gamma(r, a);
    const double gb = mt_std_gamma(r, b);
    return ga / (ga + gb);
}

// explicit mul / add: the build runs with -fmad=false, these keep the separate roundings visible in the source too
__device__ __forceinline__ double affine(double loc, double scale, double z) { return __dadd_rn(loc, __dmul_rn(scale, z)); }

__device__ double mt_sample(Mt19937& r, int family, double p0, double p1, double default_mean) {  // Arm.sample :66-122
    switch (family) {
        case EF_ARM_NORMAL: return affine(p0, p1, r.n01());
        case EF_ARM_DEFAULT_NORMAL: return affine(default_mean + p0, p1, r.n01());
        case EF_ARM_UNIFORM: return affine(p0, p1 - p0, r.u());
        case EF_ARM_EXPONENTIAL: return p0 * r.std_exp();
        case EF_ARM_GAMMA: return p1 * mt_std_gamma(r, p0);
        case EF_ARM_BETA: return mt_beta(r, p0, p1);
        case EF_ARM_LOGISTIC: {
            double u = r.u();
            while (u <= 0.0) u = r.u();
            return affine(p0, p1, log(u / (1.0 - u)));
        }
        case EF_ARM_WEIBULL: return p0 == 0.0 ? 0.0 : pow(r.std_exp(), 1.0 / p0);
        default: return exp(affine(p0, p1, r.n01()));  // EF_ARM_LOGNORMAL
    }
}

__global__ void __launch_bounds__(kThreads) bandit_step_mt19937_kernel(const BanditMtArgs a) {
    EpisodeAcc acc;
    ErrorAcc errs;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int32_t act = __ldg(a.action + i);
        if (a.obs) a.obs[i] = act;
        if (act < 0 || act >= a.k) {  // the reference asserts (:188) before sampling: no draw is consumed
            errs.record(static_cast<uint32_t>(a.env_offset + i), static_cast<uint32_t>(act), EF_STEP_BAD_ACTION);
            if (a.reward) a.reward[i] = 0.0f;
            if (a.reward64) a.reward64[i] = 0.0;
            continue;
        }
        Mt19937 r;
        r.key = a.key + i;
        r.n = a.n;
        r.pos = a.pos[i];
        r.has_gauss = a.has_gauss[i] != 0;
        r.gauss = a.gauss[i];
        const ef_arm64 arm = a.arms[act];
        const double mean = (arm.family == EF_ARM_DEFAULT_NORMAL && a.means) ? a.means[i * a.k + act] : 0.0;
        const double v = mt_sample(r, arm.family, arm.p0, arm.p1, mean);
        a.pos[i] = r.pos;
        a.has_gauss[i] = r.has_gauss ? 1 : 0;
        a.gauss[i] = r.gauss;
        const float v32 = static_cast<float>(v);
        if (a.reward) a.reward[i] = v32;
        if (a.reward64) a.reward64[i] = v;
        acc.steps += 1u;
        acc.finish(1, v32);
    }
    cta_epilogue(acc, errs, a.stats, a.err, nullptr, 0ull, 0ull);
}

}  // namespace ef

using namespace ef;

extern "C" int ef_bandit_step_mt19937(const ef_arm64* arms, int32_t k, const int32_t* action, uint32_t* key, int32_t* pos,
                                      int32_t* has_gauss, double* gauss, const double* means, uint64_t env_offset,
                                      int32_t* obs, float* reward, double* reward64, double* stats, uint32_t* err,
                                      int64_t n, void* stream) {
    if (!arms || k <= 0 || n < 0 || !action || !key || !pos || !has_gauss || !gauss) return EF_ERR_INVALID_ARGUMENT;
    if (!aligned(arms, 8) || !aligned(gauss, 8) || !aligned(means, 8) || !aligned(reward64, 8)) return EF_ERR_INVALID_ARGUMENT;
    if (env_offset + static_cast<uint64_t>(n) > (1ull << 32)) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    BanditMtArgs a{};
    a.arms = arms; a.k = k; a.action = action; a.key = key; a.pos = pos; a.has_gauss = has_gauss; a.gauss = gauss;
    a.means = means; a.obs = obs; a.reward = reward; a.reward64 = reward64; a.stats = stats; a.err = err;
    a.env_offset = env_offset; a.n = n;
    bandit_step_mt19937_kernel<<<grid_for(n, 8), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return cuda_status(cudaGetLastError());
}
