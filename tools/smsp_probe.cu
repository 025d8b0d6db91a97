// smsp_probe.cu -- which warps of a CTA share a scheduler (SM sub-partition)?  Warps 0 and j run a dependent-free FFMA
// stream that saturates one scheduler's FP32 pipe; when they sit on the same scheduler the pair takes twice as long.
// Also prints %warpid of every warp, to check "scheduler = %warpid % 4".
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smsp_probe smsp_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int* slots, int wa, int wb, int iters) {
    const int w = threadIdx.x >> 5;
    unsigned slot; asm volatile("mov.u32 %0, %%warpid;" : "=r"(slot));
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) slots[w] = (int)slot;
    if (w != wa && w != wb) return;
    float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f, a4 = 4.f, a5 = 5.f, a6 = 6.f, a7 = 7.f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f);
            a4 = fmaf(a4, 1.0001f, 0.5f); a5 = fmaf(a5, 1.0001f, 0.5f); a6 = fmaf(a6, 1.0001f, 0.5f); a7 = fmaf(a7, 1.0001f, 0.5f);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
    const int threads = 352, nw = threads / 32;
    float* out; int* slots; cudaMalloc(&out, 148 * threads * 4); cudaMalloc(&slots, 64 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int hs[64];
    for (int j = 1; j < nw; j++) {
        k<<<148, threads>>>(out, slots, 0, j, 1000);
        cudaDeviceSynchronize();
        cudaEventRecord(e0); k<<<148, threads>>>(out, slots, 0, j, 20000); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(hs, slots, 64 * 4, cudaMemcpyDeviceToHost);
        printf("warps 0 (slot %2d) + %2d (slot %2d): %.3f ms\n", hs[0], j, hs[j], ms);
    }
    return 0;
}
