// packed_peak_bench.cu -- how many scheduler cycles does one packed FP32x2 instruction of the exact FIR stream
// (FMUL2.FTZ + FADD2 pairs, 16 accumulators) really take, and which companion instruction of the real strip loop
// (LDS.64 sample fetch, LDCU.64 tap fetch, loop branch) moves that number?  Cycles are read with clock64 inside the
// kernel, so the SM clock does not enter.
//   mode 0: registers only (samples and taps in registers), 512 packed per trip
//   mode 1: samples from shared memory (one LDS.64 per 32 packed), taps in registers
//   mode 2: samples in registers, taps from the constant bank with a uniform index (LDCU.64 per 16 packed)
//   mode 3: both, the shape of fir_strip's steady loop
//   mode 4: mode 3 with FFMA2 (the fast mode's stream), half the packed instructions per tap
//   mode 5: the tap walk of rx_front2.cuh (16 samples in registers, one LDS.64 + one LDCU.64 per 16 tap updates); 6 / 7: the same
//           with the products of 8 / 4 outputs grouped in the source (ptxas reorders them anyway)
// Result on B200: 2.00-2.03 cycles per packed instruction from two warps per scheduler up, 2.16 with one -- in every mode
// that has enough registers.  Each LDS costs one more issue cycle, the uniform constant loads none.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o packed_peak_bench packed_peak_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2ftz(u64 a, u64 b) { u64 r; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

struct Taps { float2 t[256]; };

template <int MODE>
__global__ void __launch_bounds__(MODE >= 5 ? 512 : 1024, 1) k(const float2* __restrict__ xin, const __grid_constant__ Taps tb, float2* out, long long* cyc, int iters) {
    constexpr int R = 16;
    __shared__ u64 xs[32][161];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 32 * 160; i += blockDim.x) { float2 v = xin[i & 1023]; xs[i / 160][i % 160] = pk(v.x, v.y); }
    __syncthreads();
    u64 acc[R], xr[R], cr[R];
#pragma unroll
    for (int r = 0; r < R; r++) { acc[r] = 0ull; float2 v = xin[(threadIdx.x + r) & 1023]; xr[r] = pk(v.x, v.y); cr[r] = pk(tb.t[r].x, tb.t[r].x); }
    const u64* xrow = &xs[lane][(w * 16) & 15];
    long long t0 = clock64();
    if (MODE == 6 || MODE == 7) {
        constexpr int G = MODE == 6 ? 8 : 4;
        u64 X[R];
#pragma unroll
        for (int r = 0; r < R; r++) X[r] = xrow[r];
#pragma unroll 1
        for (int it = 0; it < 2 * iters; it++) {
            const int k0 = (it & 7) * 16;
            const u64* xp = xrow + 16 + k0;
#pragma unroll
            for (int j = 0; j < R; j++) {
                const u64 cc = *reinterpret_cast<const u64*>(&tb.t[k0 + j]);
#pragma unroll
                for (int r0 = 0; r0 < R; r0 += G) {
                    u64 p[G];
#pragma unroll
                    for (int g = 0; g < G; g++) asm volatile("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(p[g]) : "l"(X[(r0 + g + j) & 15]), "l"(cc));
#pragma unroll
                    for (int g = 0; g < G; g++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc[r0 + g]) : "l"(p[g]));
                }
                X[j] = xp[j];
            }
        }
    } else if (MODE == 5) {
        u64 X[R];
#pragma unroll
        for (int r = 0; r < R; r++) X[r] = xrow[r];
#pragma unroll 1
        for (int it = 0; it < 2 * iters; it++) {      // 256 pairs per trip: twice the trips for the same packed count
            const int k0 = (it & 7) * 16;
            const u64* xp = xrow + 16 + k0;
#pragma unroll
            for (int j = 0; j < R; j++) {
                const u64 cc = *reinterpret_cast<const u64*>(&tb.t[k0 + j]);
#pragma unroll
                for (int r = 0; r < R; r++) acc[r] = add2(acc[r], mul2ftz(X[(r + j) & 15], cc));
                X[j] = xp[j];
            }
        }
    } else
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        const int d0 = (it & 7) * 16;
#pragma unroll
        for (int e = 0; e < R; e++) {
            const u64 xv = (MODE & 1) || MODE == 4 ? xrow[d0 + e] : xr[e];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const u64 cc = (MODE & 2) || MODE == 4 ? *reinterpret_cast<const u64*>(&tb.t[d0 + e + 16 - r]) : cr[(e + r) & 15];
                if (MODE == 4) acc[r] = fma2(xv, cc, acc[r]);
                else acc[r] = add2(acc[r], mul2ftz(xv, cc));
            }
        }
    }
    long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * 32 + w] = t1 - t0;
#pragma unroll
    for (int r = 0; r < R; r++) { float a, b; unpk(acc[r], a, b); out[((size_t)blockIdx.x * blockDim.x + threadIdx.x) * R + r] = make_float2(a, b); }
}

template <int MODE>
static void run(const float2* x, const Taps& tb, float2* out, long long* cyc, int sms, int warps_per_sched) {
    const int threads = 128 * warps_per_sched, iters = 400;
    k<MODE><<<sms, threads>>>(x, tb, out, cyc, 10);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<MODE><<<sms, threads>>>(x, tb, out, cyc, iters); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    static long long h[148 * 32 * 2];
    cudaMemcpy(h, cyc, sizeof(long long) * sms * 32, cudaMemcpyDeviceToHost);
    double mean = 0, mx = 0; int n = 0;
    for (int s = 0; s < sms; s++) for (int w = 0; w < threads / 32; w++) { double c = (double)h[s * 32 + w]; mean += c; if (c > mx) mx = c; n++; }
    mean /= n;
    const double packed_per_warp = (MODE == 4 ? 256.0 : MODE >= 5 ? 1024.0 : 512.0) * iters;
    // a scheduler runs warps_per_sched warps, so its cycles per packed instruction = elapsed / (warps x packed per warp)
    printf("{\"mode\": %d, \"warps_per_scheduler\": %d, \"ms\": %.4f, \"cycles_per_packed_mean\": %.4f, \"cycles_per_packed_max\": %.4f}\n", MODE, warps_per_sched, ms,
           mean / (packed_per_warp * warps_per_sched), mx / (packed_per_warp * warps_per_sched));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    float2* x; float2* out; long long* cyc;
    cudaMalloc(&x, 1024 * sizeof(float2)); cudaMalloc(&out, (size_t)sms * 1024 * 16 * sizeof(float2)); cudaMalloc(&cyc, sizeof(long long) * sms * 32);
    float2 hx[1024]; Taps tb;
    for (int i = 0; i < 1024; i++) hx[i] = make_float2(0.001f * (i % 97) - 0.04f, 0.002f * (i % 89) - 0.08f);
    for (int i = 0; i < 256; i++) tb.t[i] = make_float2(0.01f * (i % 13) - 0.05f, 0.01f * (i % 13) - 0.05f);
    cudaMemcpy(x, hx, sizeof hx, cudaMemcpyHostToDevice);
    printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, sms);
    for (int wps = 1; wps <= 8; wps *= 2) {
        run<1>(x, tb, out, cyc, sms, wps);
        run<2>(x, tb, out, cyc, sms, wps);
        run<3>(x, tb, out, cyc, sms, wps);
        run<4>(x, tb, out, cyc, sms, wps);
        if (wps <= 4) {
            run<5>(x, tb, out, cyc, sms, wps);
            run<6>(x, tb, out, cyc, sms, wps);
            run<7>(x, tb, out, cyc, sms, wps);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
