// Measures the FP32 FMA and INT32 (IMAD / LOP3) instruction-issue peaks of the GPU: the roofline denominators for the
// fused rollout kernels, which are bound by instruction issue rather than by HBM (MEASURED_PEAKS.json has no such figure).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_peak tools/issue_peak.cu && tools/issue_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) burn(float* out, int iters, float a, unsigned b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    unsigned y0 = threadIdx.x, y1 = y0 + 1, y2 = y0 + 2, y3 = y0 + 3, y4 = y0 + 4, y5 = y0 + 5, y6 = y0 + 6, y7 = y0 + 7;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {  // 8 independent FFMA chains
            x0 = fmaf(x0, a, 1.0f); x1 = fmaf(x1, a, 1.0f); x2 = fmaf(x2, a, 1.0f); x3 = fmaf(x3, a, 1.0f);
            x4 = fmaf(x4, a, 1.0f); x5 = fmaf(x5, a, 1.0f); x6 = fmaf(x6, a, 1.0f); x7 = fmaf(x7, a, 1.0f);
        } else if (MODE == 1) {  // 8 independent IMAD chains
            y0 = y0 * b + 1u; y1 = y1 * b + 1u; y2 = y2 * b + 1u; y3 = y3 * b + 1u;
            y4 = y4 * b + 1u; y5 = y5 * b + 1u; y6 = y6 * b + 1u; y7 = y7 * b + 1u;
        } else {  // 8 independent LOP3 chains (alu pipe)
            y0 = (y0 ^ b) & (y0 | 0x55u); y1 = (y1 ^ b) & (y1 | 0x55u); y2 = (y2 ^ b) & (y2 | 0x55u); y3 = (y3 ^ b) & (y3 | 0x55u);
            y4 = (y4 ^ b) & (y4 | 0x55u); y5 = (y5 ^ b) & (y5 | 0x55u); y6 = (y6 ^ b) & (y6 | 0x55u); y7 = (y7 ^ b) & (y7 | 0x55u);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + (float)(y0 + y1 + y2 + y3 + y4 + y5 + y6 + y7);
}

template <int MODE>
double run(const char* name, int sms) {
    const int blocks = sms * 8, iters = 1 << 14;
    float* out;
    cudaMalloc(&out, sizeof(float) * blocks * 256);
    burn<MODE><<<blocks, 256>>>(out, 64, 1.0001f, 3u);
    cudaDeviceSynchronize();
    cudaEvent_t s, e;
    cudaEventCreate(&s);
    cudaEventCreate(&e);
    cudaEventRecord(s);
    burn<MODE><<<blocks, 256>>>(out, iters, 1.0001f, 3u);
    cudaEventRecord(e);
    cudaEventSynchronize(e);
    float ms;
    cudaEventElapsedTime(&ms, s, e);
    const double inst = double(blocks) * 256 * iters * 8;  // thread-level instructions of the measured kind
    const double rate = inst / (ms * 1e-3);
    printf("%-6s %.3e thread-instr/s  (%.1f per SM per clock at 1.965 GHz)\n", name, rate, rate / sms / 1.965e9);
    cudaFree(out);
    return rate;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double f = run<0>("FFMA", sms), i = run<1>("IMAD", sms), l = run<2>("LOP3", sms);
    printf("{\"sms\": %d, \"ffma_per_s\": %.4e, \"imad_per_s\": %.4e, \"lop3_per_s\": %.4e, \"fp32_tflops\": %.2f}\n", sms, f, i, l,
           2 * f / 1e12);
    return 0;
}
