/* A plain C client of libgobblet_b200.so: no Python, no torch -- only include/gobblet_b200.h and the CUDA
 * runtime for device memory.  Plays n envs for T fused random-legal steps, steps once more with action 0
 * through gbl_step, and prints the statistics and byte sums that tests/test_gpu_c_abi.py compares with
 * the oracle.  Build: nvcc -x cu (or gcc + -lcudart) c_abi_client.c -L... -lgobblet_b200 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/gobblet_b200.h"

#define CK(call) do { int rc_ = (call); if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, gbl_last_error()); return 1; } } while (0)
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char **argv) {
    int64_t n = argc > 1 ? atoll(argv[1]) : 1000;
    int32_t T = argc > 2 ? atoi(argv[2]) : 20;
    uint64_t seed = argc > 3 ? strtoull(argv[3], 0, 10) : 5;
    void *state; int8_t *obs, *mask, *rew; uint8_t *term, *trunc, *agent; int64_t *stats, *actions;
    CU(cudaMalloc(&state, n * GBL_STATE_BYTES));
    CU(cudaMalloc((void **)&obs, n * GBL_OBS_BYTES + 16)); CU(cudaMalloc((void **)&mask, n * GBL_MASK_BYTES + 16));
    CU(cudaMalloc((void **)&rew, 2 * n)); CU(cudaMalloc((void **)&term, n)); CU(cudaMalloc((void **)&trunc, n));
    CU(cudaMalloc((void **)&agent, n)); CU(cudaMalloc((void **)&stats, 64)); CU(cudaMalloc((void **)&actions, 8 * n));
    CU(cudaMemset(stats, 0, 64)); CU(cudaMemset(actions, 0, 8 * n));
    if (gbl_abi_version() != GBL_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }
    if (gbl_observe(state, obs + 1, mask, 0, n, 0) != GBL_E_INVALID) { fprintf(stderr, "misaligned obs accepted\n"); return 1; }
    CK(gbl_reset(state, n, 0));
    CK(gbl_rollout_random(state, n, T, seed, 0, 0, 0, obs, mask, 0, 0, 1, 0, 0, 0, 0, 0, 0, stats, GBL_AUTORESET_SAME_STEP, 0));
    CK(gbl_step(state, actions, 8, obs, mask, rew, term, trunc, agent, 0, 0, stats, n, GBL_AUTORESET_SAME_STEP, 0));
    CU(cudaDeviceSynchronize());
    int8_t *h_obs = (int8_t *)malloc(n * GBL_OBS_BYTES), *h_mask = (int8_t *)malloc(n * GBL_MASK_BYTES), *h_rew = (int8_t *)malloc(2 * n);
    uint8_t *h_term = (uint8_t *)malloc(n);
    int64_t h_stats[8];
    CU(cudaMemcpy(h_obs, obs, n * GBL_OBS_BYTES, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_mask, mask, n * GBL_MASK_BYTES, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_rew, rew, 2 * n, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(h_term, term, n, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_stats, stats, 64, cudaMemcpyDeviceToHost));
    long long so = 0, sm = 0, sr = 0, st = 0;
    for (int64_t i = 0; i < n * GBL_OBS_BYTES; ++i) so += (long long)h_obs[i] * (i % 251 + 1);
    for (int64_t i = 0; i < n * GBL_MASK_BYTES; ++i) sm += (long long)h_mask[i] * (i % 251 + 1);
    for (int64_t i = 0; i < 2 * n; ++i) sr += (long long)h_rew[i] * (i % 251 + 1);
    for (int64_t i = 0; i < n; ++i) st += h_term[i];
    printf("stats %lld %lld %lld %lld %lld %lld %lld %lld\n", (long long)h_stats[0], (long long)h_stats[1], (long long)h_stats[2],
           (long long)h_stats[3], (long long)h_stats[4], (long long)h_stats[5], (long long)h_stats[6], (long long)h_stats[7]);
    printf("sums %lld %lld %lld %lld\n", so, sm, sr, st);

    /* one more step through HOST buffers in a single call (gbl_step_host: packed records over PCIe, expanded by
     * the library's thread pool): pinned actions in, reference-shaped host arrays out */
    uint8_t *ha, *da; uint32_t *hrec, *drec; int8_t *xo, *xm, *xr; uint8_t *xt, *xu, *xa;
    CU(cudaMallocHost((void **)&ha, n)); CU(cudaMallocHost((void **)&hrec, 24 * n)); CU(cudaMalloc((void **)&da, n)); CU(cudaMalloc((void **)&drec, 24 * n));
    CU(cudaMallocHost((void **)&xo, n * GBL_OBS_BYTES)); CU(cudaMallocHost((void **)&xm, n * GBL_MASK_BYTES)); CU(cudaMallocHost((void **)&xr, 2 * n));
    CU(cudaMallocHost((void **)&xt, n)); CU(cudaMallocHost((void **)&xu, n)); CU(cudaMallocHost((void **)&xa, n));
    for (int64_t i = 0; i < n; ++i) ha[i] = (uint8_t)(i % 54);
    cudaStream_t streams[2]; cudaEvent_t events[2];
    for (int c = 0; c < 2; ++c) { CU(cudaStreamCreate(&streams[c])); CU(cudaEventCreateWithFlags(&events[c], cudaEventDisableTiming)); }
    int64_t ends[2] = {n / 2 / 2 * 2, n};
    CK(gbl_step_host(state, ha, n, GBL_AUTORESET_SAME_STEP, da, drec, hrec, 2, ends, (void *)streams[0], (void *const *)events,
                     xo, xm, xr, xt, xu, xa, stats, 3));
    so = sm = sr = st = 0;
    long long su = 0, sa = 0;
    for (int64_t i = 0; i < n * GBL_OBS_BYTES; ++i) so += (long long)xo[i] * (i % 251 + 1);
    for (int64_t i = 0; i < n * GBL_MASK_BYTES; ++i) sm += (long long)xm[i] * (i % 251 + 1);
    for (int64_t i = 0; i < 2 * n; ++i) sr += (long long)xr[i] * (i % 251 + 1);
    for (int64_t i = 0; i < n; ++i) { st += xt[i]; su += xu[i]; sa += xa[i]; }
    printf("host %lld %lld %lld %lld %lld %lld\n", so, sm, sr, st, su, sa);
    return 0;
}
