// envforge_b200 -- host-buffer step entry points: the per-step call a host-side actor makes (actions in pinned host
// memory in, observations / rewards / flags in pinned host memory out), pipelined so that PCIe carries both directions
// at once.
//
// One step over n envs is cut into `chunks` contiguous index ranges.  For chunk c, on the caller's stream:
//     [copy stream s_in ]  H2D of the chunk's actions (only when a device staging buffer is given; otherwise the step
//                          kernel reads the actions straight from the pinned host buffer through its UVA pointer)
//     [caller's stream  ]  the ordinary step kernel on the chunk (global env index = env_offset + chunk start, so the
//                          Philox streams -- and therefore the results -- are those of the unchunked launch)
//     [copy stream s_out]  D2H of the chunk's outputs while the kernel of chunk c+1 runs and reads its actions
// The caller's stream finally waits for the last D2H, so a plain stream synchronise makes the host buffers valid.
#include <algorithm>
#include <new>

#include "ef_common.cuh"

namespace {
constexpr int kMaxChunks = 32;
constexpr int64_t kChunkAlign = 1024;  // chunk starts stay multiples of 1024 envs: quads and Philox groups stay intact
}  // namespace

struct ef_host_pipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t e_start = nullptr, e_done = nullptr;
    cudaEvent_t e_in[kMaxChunks] = {}, e_k[kMaxChunks] = {};
    int device = -1;
};

namespace ef {

struct HostOut {
    void* d;        // device buffer the kernel writes (whole batch)
    void* h;        // pinned host destination (whole batch)
    size_t stride;  // bytes per env
};

#define EF_CUDA(expr)                                                   \
    do {                                                                \
        const cudaError_t e__ = (expr);                                 \
        if (e__ != cudaSuccess) return cuda_status(e__);                \
    } while (0)

template <typename Launch>
int run_host_pipe(ef_host_pipe* p, const void* h_in, void* d_in, size_t in_stride, const HostOut* outs, int n_out,
                  int64_t n, int chunks, cudaStream_t st, Launch launch) {
    if (!p || !h_in || n < 0 || chunks < 1) return EF_ERR_INVALID_ARGUMENT;
    if (n == 0) return EF_OK;
    int dev = 0;
    EF_CUDA(cudaGetDevice(&dev));
    if (dev != p->device) return EF_ERR_INVALID_ARGUMENT;
    chunks = std::min(chunks, kMaxChunks);
    int64_t per = (n + chunks - 1) / chunks;
    per = (per + kChunkAlign - 1) / kChunkAlign * kChunkAlign;
    if (d_in) {  // earlier work on the caller's stream may still read the staging buffer
        EF_CUDA(cudaEventRecord(p->e_start, st));
        EF_CUDA(cudaStreamWaitEvent(p->s_in, p->e_start, 0));
    }
    int c = 0;
    for (int64_t off = 0; off < n; off += per, ++c) {
        const int64_t cnt = std::min(per, n - off);
        const char* in = static_cast<const char*>(h_in) + off * in_stride;
        if (d_in) {
            char* din = static_cast<char*>(d_in) + off * in_stride;
            EF_CUDA(cudaMemcpyAsync(din, in, cnt * in_stride, cudaMemcpyHostToDevice, p->s_in));
            EF_CUDA(cudaEventRecord(p->e_in[c], p->s_in));
            EF_CUDA(cudaStreamWaitEvent(st, p->e_in[c], 0));
            in = din;
        }
        if (const int s = launch(off, cnt, static_cast<const void*>(in))) return s;
        EF_CUDA(cudaEventRecord(p->e_k[c], st));
        EF_CUDA(cudaStreamWaitEvent(p->s_out, p->e_k[c], 0));
        for (int j = 0; j < n_out; ++j) {
            if (!outs[j].d || !outs[j].h) continue;
            EF_CUDA(cudaMemcpyAsync(static_cast<char*>(outs[j].h) + off * outs[j].stride,
                                    static_cast<const char*>(outs[j].d) + off * outs[j].stride, cnt * outs[j].stride,
                                    cudaMemcpyDeviceToHost, p->s_out));
        }
    }
    EF_CUDA(cudaEventRecord(p->e_done, p->s_out));
    EF_CUDA(cudaStreamWaitEvent(st, p->e_done, 0));
    return EF_OK;
}

template <typename T>
inline T* at(T* base, int64_t off, int width = 1) {
    return base ? base + off * width : nullptr;
}

}  // namespace ef

using namespace ef;

extern "C" int ef_host_pipe_create(ef_host_pipe** out) {
    if (!out) return EF_ERR_INVALID_ARGUMENT;
    ef_host_pipe* p = new (std::nothrow) ef_host_pipe();
    if (!p) return EF_ERR_INVALID_ARGUMENT;
    cudaError_t e = cudaGetDevice(&p->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->e_start, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->e_done, cudaEventDisableTiming);
    for (int i = 0; i < kMaxChunks && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&p->e_in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->e_k[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        ef_host_pipe_destroy(p);
        return cuda_status(e);
    }
    *out = p;
    return EF_OK;
}

extern "C" int ef_host_pipe_destroy(ef_host_pipe* p) {
    if (!p) return EF_OK;
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    if (p->e_start) cudaEventDestroy(p->e_start);
    if (p->e_done) cudaEventDestroy(p->e_done);
    for (int i = 0; i < kMaxChunks; ++i) {
        if (p->e_in[i]) cudaEventDestroy(p->e_in[i]);
        if (p->e_k[i]) cudaEventDestroy(p->e_k[i]);
    }
    delete p;
    return EF_OK;
}

extern "C" int ef_gridworld_step_host(ef_host_pipe* pipe, const ef_grid_desc* grid, int32_t* pos, int32_t* ep_len,
                                      float* ep_ret, const int32_t* h_action, int32_t* d_action, uint64_t seed,
                                      uint64_t t, uint64_t env_offset, int32_t* d_obs, float* d_reward,
                                      uint8_t* d_terminated, uint8_t* d_truncated, int32_t* d_final_obs, int32_t* h_obs,
                                      float* h_reward, uint8_t* h_terminated, uint8_t* h_truncated, int32_t* h_final_obs,
                                      double* stats, uint32_t* err, int64_t n, int32_t autoreset, int32_t chunks,
                                      void* stream) {
    const HostOut outs[5] = {{d_obs, h_obs, 8}, {d_reward, h_reward, 4}, {d_terminated, h_terminated, 1},
                             {d_truncated, h_truncated, 1}, {d_final_obs, h_final_obs, 8}};
    return run_host_pipe(pipe, h_action, d_action, 4, outs, 5, n, chunks, static_cast<cudaStream_t>(stream),
                         [&](int64_t off, int64_t cnt, const void* in) {
                             return ef_gridworld_step(grid, pos + off, at(ep_len, off), ep_ret + off,
                                                      static_cast<const int32_t*>(in), nullptr, seed, t, env_offset + off,
                                                      at(d_obs, off, 2), at(d_reward, off), at(d_terminated, off),
                                                      at(d_truncated, off), at(d_final_obs, off, 2), stats, err, nullptr,
                                                      cnt, autoreset, stream);
                         });
}

extern "C" int ef_gridworld_step_u8_host(ef_host_pipe* pipe, const ef_grid_desc* grid, int32_t* pos, int32_t* ep_len,
                                         float* ep_ret, const uint8_t* h_action, uint8_t* d_action, uint64_t seed,
                                         uint64_t t, uint64_t env_offset, uint8_t* d_obs, float* d_reward,
                                         uint8_t* d_terminated, uint8_t* d_truncated, uint8_t* d_final_obs, uint8_t* h_obs,
                                         float* h_reward, uint8_t* h_terminated, uint8_t* h_truncated, uint8_t* h_final_obs,
                                         double* stats, uint32_t* err, int64_t n, int32_t autoreset, int32_t chunks,
                                         void* stream) {
    const HostOut outs[5] = {{d_obs, h_obs, 2}, {d_reward, h_reward, 4}, {d_terminated, h_terminated, 1},
                             {d_truncated, h_truncated, 1}, {d_final_obs, h_final_obs, 2}};
    return run_host_pipe(pipe, h_action, d_action, 1, outs, 5, n, chunks, static_cast<cudaStream_t>(stream),
                         [&](int64_t off, int64_t cnt, const void* in) {
                             return ef_gridworld_step_u8(grid, pos + off, at(ep_len, off), ep_ret + off,
                                                         static_cast<const uint8_t*>(in), nullptr, seed, t, env_offset + off,
                                                         at(d_obs, off, 2), at(d_reward, off), at(d_terminated, off),
                                                         at(d_truncated, off), at(d_final_obs, off, 2), stats, err, nullptr,
                                                         cnt, autoreset, stream);
                         });
}

#define EF_PENDULUM_STEP_HOST(NAME, STEP, DIM, ACCW)                                                                    \
    extern "C" int NAME(ef_host_pipe* pipe, const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret,    \
                        const float* h_action, float* d_action, uint64_t seed, uint64_t t, uint64_t env_offset,          \
                        float* d_obs, float* d_reward, uint8_t* d_terminated, uint8_t* d_truncated, float* d_final_obs,  \
                        float* d_acc, float* h_obs, float* h_reward, uint8_t* h_terminated, uint8_t* h_truncated,        \
                        float* h_final_obs, float* h_acc, double* stats, uint32_t* err, int64_t n, int32_t autoreset,    \
                        int32_t chunks, void* stream) {                                                                  \
        const HostOut outs[6] = {{d_obs, h_obs, 4 * DIM},       {d_reward, h_reward, 4},                                 \
                                 {d_terminated, h_terminated, 1}, {d_truncated, h_truncated, 1},                         \
                                 {d_final_obs, h_final_obs, 4 * DIM}, {d_acc, h_acc, 4 * ACCW}};                         \
        return run_host_pipe(pipe, h_action, d_action, 4, outs, 6, n, chunks, static_cast<cudaStream_t>(stream),         \
                             [&](int64_t off, int64_t cnt, const void* in) {                                             \
                                 return STEP(p, state + off * DIM, ep_len + off, ep_ret + off,                           \
                                             static_cast<const float*>(in), seed, t, env_offset + off,                   \
                                             at(d_obs, off, DIM), at(d_reward, off), at(d_terminated, off),              \
                                             at(d_truncated, off), at(d_final_obs, off, DIM), at(d_acc, off, ACCW),      \
                                             stats, err, nullptr, cnt, autoreset, stream);                               \
                             });                                                                                         \
    }

EF_PENDULUM_STEP_HOST(ef_cartpole_step_host, ef_cartpole_step, 4, 2)
EF_PENDULUM_STEP_HOST(ef_pendulum_disk_step_host, ef_pendulum_disk_step, 2, 1)
#undef EF_PENDULUM_STEP_HOST

extern "C" int ef_bandit_step_host(ef_host_pipe* pipe, const ef_arm* arms, int32_t k, const int32_t* h_action,
                                   int32_t* d_action, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t* d_obs,
                                   float* d_reward, int32_t* h_obs, float* h_reward, double* stats, uint32_t* err,
                                   int64_t n, int32_t chunks, void* stream) {
    const HostOut outs[2] = {{d_obs, h_obs, 4}, {d_reward, h_reward, 4}};
    return run_host_pipe(pipe, h_action, d_action, 4, outs, 2, n, chunks, static_cast<cudaStream_t>(stream),
                         [&](int64_t off, int64_t cnt, const void* in) {
                             return ef_bandit_step(arms, k, static_cast<const int32_t*>(in), seed, t0, env_offset + off,
                                                   1, at(d_obs, off), at(d_reward, off), stats, err, cnt, stream);
                         });
}
