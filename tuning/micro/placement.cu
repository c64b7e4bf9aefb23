// where do the warps of co-resident CTAs land?  prints %smid / %warpid per (CTA, warp) for a persistent grid
// nvcc -arch=sm_100a -o placement placement.cu ; ./placement <warps per CTA> <CTAs per SM> <smem KB>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void k(unsigned *out, int nw) {
    extern __shared__ unsigned char sm[];
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    if ((threadIdx.x & 31) == 0) {
        const int w = threadIdx.x >> 5;
        out[(blockIdx.x * nw + w) * 2 + 0] = smid;
        out[(blockIdx.x * nw + w) * 2 + 1] = warpid;
    }
    const long long t0 = clock64();
    while (clock64() - t0 < 400000) { }   // keep every CTA resident
    if (threadIdx.x == 0) sm[0] = 1;
}
int main(int argc, char **argv) {
    const int nw = argc > 1 ? atoi(argv[1]) : 6, cps = argc > 2 ? atoi(argv[2]) : 2, kb = argc > 3 ? atoi(argv[3]) : 59;
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * cps;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024);
    unsigned *d, *h = (unsigned *)malloc(grid * nw * 8);
    cudaMalloc(&d, grid * nw * 8);
    k<<<grid, nw * 32, kb * 1024>>>(d, nw);
    cudaMemcpy(h, d, grid * nw * 8, cudaMemcpyDeviceToHost);
    printf("warps/CTA %d, CTAs/SM %d, smem %d KB: %s\n", nw, cps, kb, cudaGetErrorString(cudaGetLastError()));
    for (int b : {0, 1, 2, 147, 148, 149, 150, 295}) {
        if (b >= grid) continue;
        printf(" CTA %3d: sm %3u warpids", b, h[b * nw * 2]);
        for (int w = 0; w < nw; w++) printf(" %2u", h[(b * nw + w) * 2 + 1]);
        printf("\n");
    }
    // per SM: which CTAs
    for (int s : {0, 1, 2}) {
        printf(" SM %d holds CTAs:", s);
        for (int b = 0; b < grid; b++) if (h[b * nw * 2] == (unsigned)s) printf(" %d(w0 slot %u)", b, h[b * nw * 2 + 1]);
        printf("\n");
    }
    return 0;
}
