/* Plain-C host of the envforge_b200 C ABI: no Python, no torch -- what a cgo / JNI / ctypes-free caller would do.
 * Reads a compiled GridWorld table + actions from a file, steps n envs for `steps` steps through ef_gridworld_step on
 * device buffers it allocates itself with the CUDA runtime, and writes every step's outputs to another file.
 * tests/test_gpu_c_abi_host.py builds the input with the package's host table compiler and compares the output with
 * VecGridWorld.step bit for bit.
 *
 *   input : int32 header[16] = {rows, cols, n_slots, idx_cap, thr0, thr1, thr2, start_cell, episode_limit, table_cells,
 *                               n, steps, seed, packed, 0, 0}
 *           ef_grid_outcome table[table_cells * 4 * n_slots];  int32 actions[steps][n]
 *   output: per step: int32 obs[n][2], float reward[n], uint8 terminated[n], uint8 truncated[n];
 *           then double stats[EF_STATS_LEN] (sum over the replicas), uint32 err[EF_ERR_LEN]
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "envforge_b200.h"

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                  \
            return 2;                                                                 \
        }                                                                             \
    } while (0)
#define STATS_REPLICAS EF_STATS_REPLICAS /* the accumulator block the kernels add into: replicas x stride doubles */
#define STATS_STRIDE EF_STATS_STRIDE

int main(int argc, char** argv) {
    if (argc != 3) {
        fprintf(stderr, "usage: %s input.bin output.bin\n", argv[0]);
        return 1;
    }
    if (ef_abi_version() != EF_ABI_VERSION) {
        fprintf(stderr, "ABI mismatch\n");
        return 1;
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 1;
    int32_t h[16];
    if (fread(h, sizeof(int32_t), 16, f) != 16) return 1;
    const int64_t n = h[10];
    const int steps = h[11], packed = h[13];
    const size_t entries = (size_t)h[9] * 4 * (size_t)h[2];
    ef_grid_outcome* table = (ef_grid_outcome*)malloc(entries * sizeof(ef_grid_outcome));
    int32_t* actions = (int32_t*)malloc((size_t)steps * n * sizeof(int32_t));
    if (fread(table, sizeof(ef_grid_outcome), entries, f) != entries) return 1;
    if (fread(actions, sizeof(int32_t), (size_t)steps * n, f) != (size_t)steps * n) return 1;
    fclose(f);

    ef_grid_outcome* d_table;
    int32_t *d_pos, *d_len = NULL, *d_act, *d_obs;
    float *d_ret, *d_rew;
    uint8_t *d_term, *d_trunc;
    double* d_stats;
    uint32_t* d_err;
    CK(cudaMalloc((void**)&d_table, entries * sizeof(ef_grid_outcome)));
    CK(cudaMemcpy(d_table, table, entries * sizeof(ef_grid_outcome), cudaMemcpyHostToDevice));
    CK(cudaMalloc((void**)&d_pos, n * 4));
    if (!packed) CK(cudaMalloc((void**)&d_len, n * 4));
    CK(cudaMalloc((void**)&d_ret, n * 4));
    CK(cudaMalloc((void**)&d_act, n * 4));
    CK(cudaMalloc((void**)&d_obs, n * 8));
    CK(cudaMalloc((void**)&d_rew, n * 4));
    CK(cudaMalloc((void**)&d_term, n));
    CK(cudaMalloc((void**)&d_trunc, n));
    CK(cudaMalloc((void**)&d_stats, STATS_REPLICAS * STATS_STRIDE * sizeof(double)));
    CK(cudaMalloc((void**)&d_err, EF_ERR_LEN * sizeof(uint32_t)));
    CK(cudaMemset(d_stats, 0, STATS_REPLICAS * STATS_STRIDE * sizeof(double)));
    CK(cudaMemset(d_err, 0, EF_ERR_LEN * sizeof(uint32_t)));

    ef_grid_desc d;
    memset(&d, 0, sizeof d);
    d.rows = h[0]; d.cols = h[1]; d.n_slots = h[2]; d.idx_cap = h[3];
    d.thr[0] = (uint32_t)h[4]; d.thr[1] = (uint32_t)h[5]; d.thr[2] = (uint32_t)h[6];
    d.start_cell = h[7]; d.episode_limit = h[8]; d.table_cells = h[9];
    d.table = d_table;
    const uint64_t seed = (uint64_t)(uint32_t)h[12];

    int rc = ef_gridworld_reset(d.start_cell, d.cols, d_pos, d_len, d_ret, d_obs, NULL, n, NULL);
    if (rc) {
        fprintf(stderr, "reset: %s (%s)\n", ef_status_string(rc), ef_last_cuda_error_string());
        return 3;
    }
    FILE* o = fopen(argv[2], "wb");
    if (!o) return 1;
    int32_t* obs = (int32_t*)malloc(n * 8);
    float* rew = (float*)malloc(n * 4);
    uint8_t* flags = (uint8_t*)malloc(n);
    for (int t = 0; t < steps; ++t) {
        CK(cudaMemcpy(d_act, actions + (size_t)t * n, n * 4, cudaMemcpyHostToDevice));
        rc = ef_gridworld_step(&d, d_pos, d_len, d_ret, d_act, NULL, seed, (uint64_t)t, 0, d_obs, d_rew, d_term, d_trunc,
                               NULL, d_stats, d_err, NULL, n, 1, NULL /* default stream */);
        if (rc) {
            fprintf(stderr, "step %d: %s (%s)\n", t, ef_status_string(rc), ef_last_cuda_error_string());
            return 3;
        }
        CK(cudaMemcpy(obs, d_obs, n * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(rew, d_rew, n * 4, cudaMemcpyDeviceToHost));
        fwrite(obs, 8, n, o);
        fwrite(rew, 4, n, o);
        CK(cudaMemcpy(flags, d_term, n, cudaMemcpyDeviceToHost));
        fwrite(flags, 1, n, o);
        CK(cudaMemcpy(flags, d_trunc, n, cudaMemcpyDeviceToHost));
        fwrite(flags, 1, n, o);
    }
    double raw[STATS_REPLICAS * STATS_STRIDE], stats[EF_STATS_LEN] = {0};
    uint32_t err[EF_ERR_LEN];
    CK(cudaMemcpy(raw, d_stats, sizeof raw, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(err, d_err, sizeof err, cudaMemcpyDeviceToHost));
    for (int r = 0; r < STATS_REPLICAS; ++r)
        for (int k = 0; k < EF_STATS_LEN; ++k) stats[k] += raw[r * STATS_STRIDE + k];
    fwrite(stats, sizeof(double), EF_STATS_LEN, o);
    fwrite(err, sizeof(uint32_t), EF_ERR_LEN, o);
    fclose(o);
    printf("ok: %lld envs x %d steps, %.0f episodes finished\n", (long long)n, steps, stats[0]);
    return 0;
}
