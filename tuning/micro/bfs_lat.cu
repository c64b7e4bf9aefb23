// microbenchmark: latency of warp_bfs per level, one warp alone vs several warps per SM
#include "../../reinforcement-learning-for-multi-agv-pathfinding_b200/csrc/agv_core.cuh"
#include <cstdio>
using namespace agv;
template <int MODE>
__global__ void k(long long *out, int reps) {
    const int lane = threadIdx.x & 31;
    Row<1> v[2];
    // MODE 0: open grid, corner to corner (all phase 1).  MODE 1: serpentine walls (phase 2 heavy)
    for (int s = 0; s < 2; s++) {
        const int r = lane * 2 + s;
        uint64_t row = ~0ull;
        if (MODE == 1 && (r & 1)) row = ((r >> 1) & 1) ? 1ull : (1ull << 63);   // wall rows with one gap
        v[s].w[0] = row;
    }
    long long t0 = clock64();
    int len = 0, acc = 0;
    for (int i = 0; i < reps; i++) {
        int a = warp_bfs<1, 2>(v, MODE == 1 ? 1 : 0, 0, MODE == 1 ? 2 : 63, MODE == 1 ? 62 : 63, 64, 64, lane, len);
        acc += a + len;
        v[0].w[0] |= (acc == -12345);  // keep the loop alive
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = len; out[2] = acc; }
}
int main() {
    long long *d, h[3];
    cudaMalloc(&d, 24);
    for (int mode = 0; mode < 2; mode++)
        for (int warps = 1; warps <= 16; warps *= 2) {
            const int reps = 50;
            if (mode == 0) k<0><<<148, 32 * warps>>>(d, reps); else k<1><<<148, 32 * warps>>>(d, reps);
            cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
            printf("mode %d warps/SM %2d: len %lld  cycles/search %lld  cycles/level %.1f\n", mode, warps, h[1], h[0] / reps, (double)h[0] / reps / (double)h[1]);
        }
    return 0;
}
