// warp_slots.cu -- which hardware warp slot (and so which of the four schedulers, slot % 4) does warp w of a CTA get
// when two CTAs of 352 (or 384) threads share an SM?  Informs the placement of the auxiliary warps in rx_front_kernel.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_slots warp_slots.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* out, int nwarps, long long spin) {
    extern __shared__ char smem[];
    const int w = threadIdx.x >> 5;
    unsigned warpid, smid;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    long long t0 = clock64();
    while (clock64() - t0 < spin) { }
    if ((threadIdx.x & 31) == 0) { out[(blockIdx.x * nwarps + w) * 2] = (int)smid; out[(blockIdx.x * nwarps + w) * 2 + 1] = (int)warpid; }
    smem[threadIdx.x] = 0;
}
int main() {
    for (int threads : {352, 384}) {
        const int nw = threads / 32, grid = 148 * 2 * 3;      // three waves
        int* d; cudaMalloc(&d, grid * nw * 2 * sizeof(int));
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
        k<<<grid, threads, 110 * 1024>>>(d, nw, 200000);
        cudaDeviceSynchronize();
        int* h = new int[grid * nw * 2];
        cudaMemcpy(h, d, grid * nw * 2 * sizeof(int), cudaMemcpyDeviceToHost);
        printf("threads %d\n", threads);
        int shown = 0;
        for (int b = 0; b < grid && shown < 10; b++) {
            if (h[b * nw * 2] != 0 && h[b * nw * 2] != 5) continue;     // CTAs that ran on SM 0 or 5
            printf("  cta %4d sm %3d slots:", b, h[b * nw * 2]);
            for (int w = 0; w < nw; w++) printf(" %2d", h[(b * nw + w) * 2 + 1]);
            printf("   sched:");
            for (int w = 0; w < nw; w++) printf(" %d", h[(b * nw + w) * 2 + 1] & 3);
            printf("\n");
            shown++;
        }
        // histogram over all CTAs: scheduler of warp w
        for (int w = 0; w < nw; w++) {
            int c[4] = {0, 0, 0, 0};
            for (int b = 0; b < grid; b++) c[h[(b * nw + w) * 2 + 1] & 3]++;
            printf("  warp %2d -> sched counts %d %d %d %d\n", w, c[0], c[1], c[2], c[3]);
        }
        delete[] h; cudaFree(d);
    }
    return 0;
}
