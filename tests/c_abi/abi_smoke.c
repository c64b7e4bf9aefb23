/* Plain-C consumer of include/agv_b200.h (no Python, no torch): what a cgo / JNI / ctypes
 * binding does.  Runs `ticks` fused ticks (intelligent protocol + A* guidance, clip encode)
 * of 2 scenes on the README 4,2,2,2,2 layout with the task list read from argv-free stdin
 * and prints one line per tick: the action log of every AGV.
 *   cc abi_smoke.c -I<repo>/include -I/usr/local/cuda/include -L<pkg> -lagv_b200 -lcudart */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "agv_b200.h"

#define CHECK(x) do { int _rc = (x); if (_rc != AGV_OK) { fprintf(stderr, "%s -> %s\n", #x, agv_strerror(_rc)); return 2; } } while (0)

int main(void) {
    int H, W, E, N, T, ticks;
    if (scanf("%d %d %d %d %d %d", &H, &W, &E, &N, &T, &ticks) != 6) return 1;
    uint8_t *grid = (uint8_t *)malloc((size_t)H * W);
    for (int i = 0; i < H * W; i++) { int v; if (scanf("%d", &v) != 1) return 1; grid[i] = (uint8_t)v; }
    uint8_t *tasks = (uint8_t *)malloc((size_t)E * T * 4);
    for (int i = 0; i < E * T * 4; i++) { int v; if (scanf("%d", &v) != 1) return 1; tasks[i] = (uint8_t)v; }
    AgvConfig cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.H = H; cfg.W = W; cfg.E = E; cfg.N = N; cfg.T = T; cfg.P = 3;
    cfg.max_steps = (int64_t)((W / 2.0 + H) * 3 * ((W * H) / 2.5));
    cfg.device = 0;
    AgvHandle *h = NULL;
    CHECK(agv_create(&cfg, grid, &h));
    CHECK(agv_set_tasks(h, tasks, 0, E, NULL));
    CHECK(agv_reset(h, NULL, NULL));
    int8_t *d_act; float *d_rew; uint8_t *d_done; void *d_obs;
    cudaMalloc((void **)&d_act, (size_t)E * N);
    cudaMalloc((void **)&d_rew, (size_t)E * N * sizeof(float));
    cudaMalloc((void **)&d_done, (size_t)E * N);
    cudaMalloc(&d_obs, (size_t)E * N * 3 * 7 * 7 * 2);
    int8_t *act = (int8_t *)malloc((size_t)E * N);
    float *rew = (float *)malloc((size_t)E * N * sizeof(float));
    for (int t = 0; t < ticks; t++) {
        CHECK(agv_tick(h, AGV_PROTO_INTELLIGENT, AGV_POLICY_GUIDED, AGV_OBS_CLIP, NULL, d_act, d_rew, d_done, NULL,
                       NULL, d_obs, NULL, NULL));
        cudaMemcpy(act, d_act, (size_t)E * N, cudaMemcpyDeviceToHost);
        cudaMemcpy(rew, d_rew, (size_t)E * N * sizeof(float), cudaMemcpyDeviceToHost);
        for (int i = 0; i < E * N; i++) printf("%d:%g ", (int)act[i], (double)rew[i]);
        printf("\n");
    }
    uint64_t cnt[3];
    CHECK(agv_read_counters(h, cnt, NULL));
    printf("turns %llu launches %llu %s\n", (unsigned long long)cnt[0], (unsigned long long)agv_launch_count(),
           agv_version());
    CHECK(agv_destroy(h));
    return 0;
}
