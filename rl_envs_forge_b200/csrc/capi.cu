// envforge_b200 -- version / status / device helpers and the in-flight episode statistics reduction.
#include <cstdlib>

#include "ef_common.cuh"

namespace ef {

thread_local cudaError_t g_last_cuda_error = cudaSuccess;

int sm_count() {
    static thread_local int cached_dev = -1, cached_sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return cached_sms;
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) {
            cached_sms = sms;
            cached_dev = dev;
        }
    }
    return cached_sms;
}

bool pdl_enabled() {
    static const bool on = (getenv("ENVFORGE_NO_PDL") == nullptr);
    return on;
}

__global__ void __launch_bounds__(kThreads) episode_stat_reduce_kernel(const int32_t* __restrict__ ep_len,
                                                                       const float* __restrict__ ep_ret,
                                                                       double* __restrict__ out, int64_t n) {
    EpisodeAcc acc;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        acc.finish(ep_len[i], ep_ret[i]);
    cta_epilogue(acc, ErrorAcc(), out, nullptr, nullptr, 0ull, 0ull);
}

}  // namespace ef

using namespace ef;

extern "C" int ef_abi_version(void) { return EF_ABI_VERSION; }

extern "C" const char* ef_status_string(int status) {
    switch (status) {
        case EF_OK: return "EF_OK";
        case EF_ERR_INVALID_ARGUMENT: return "EF_ERR_INVALID_ARGUMENT";
        case EF_ERR_CUDA: return "EF_ERR_CUDA";
        case EF_ERR_UNSUPPORTED: return "EF_ERR_UNSUPPORTED";
        default: return "EF_ERR_UNKNOWN";
    }
}

extern "C" int ef_last_cuda_error(void) { return static_cast<int>(g_last_cuda_error); }

extern "C" const char* ef_last_cuda_error_string(void) { return cudaGetErrorName(g_last_cuda_error); }

extern "C" int ef_device_info(int32_t* sm, int32_t* cc_major, int32_t* cc_minor) {
    int dev = 0;
    if (cudaError_t e = cudaGetDevice(&dev)) return cuda_status(e);
    int v = 0;
    if (sm) { if (cudaError_t e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)) return cuda_status(e); *sm = v; }
    if (cc_major) { if (cudaError_t e = cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)) return cuda_status(e); *cc_major = v; }
    if (cc_minor) { if (cudaError_t e = cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)) return cuda_status(e); *cc_minor = v; }
    return EF_OK;
}

extern "C" int ef_episode_stat_reduce(const int32_t* ep_len, const float* ep_ret, double* out, int64_t n, void* stream) {
    if (n < 0 || !ep_len || !ep_ret || !out) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    episode_stat_reduce_kernel<<<grid_for(n, 8), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(ep_len, ep_ret, out, n);
    return cuda_status(cudaGetLastError());
}
