// fp32_pipe_bench.cu -- issue-rate microbenchmark for the FIR inner-loop candidates on sm_100a.
//
// The RRC FIR (reference rrc_fir.c:22-26) is one rounded multiply and one rounded add per
// tap per component.  This program measures how many complex tap-updates per clock per SM
// each SASS formulation sustains, so DESIGN.md can quote a measured FP32 ceiling:
//   v0  FMUL,FMUL,FADD,FADD     (scalar, exact)
//   v1  FMUL,FMUL,FADD2         (exact)
//   v2  FMUL2,FADD,FADD         (exact)
//   v3  FMUL2.FTZ,FADD2         (exact unless a product is subnormal; .ftz only stops ptxas
//                                12.9 from contracting mul.f32x2+add.f32x2 into FFMA2)
//   v4  FFMA2                   (fast mode, fused)
//   v5  FFMA,FFMA               (fast mode, scalar fused)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipe_bench fp32_pipe_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2ftz(u64 a, u64 b) { u64 r; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

#define R 8        // independent accumulators per thread
#define TAPS 8     // taps per inner pass (kept in registers)
#define WIN (R + TAPS - 1)

template <int V>
__global__ void __launch_bounds__(256) pipe_kernel(const float2* __restrict__ xin, const float* __restrict__ taps, float2* out, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    constexpr bool PACKED_X = (V == 2 || V == 3 || V == 4);
    constexpr bool PACKED_ACC = (V == 1 || V == 3 || V == 4);
    float c[TAPS]; u64 cc[TAPS];
#pragma unroll
    for (int i = 0; i < TAPS; i++) { c[i] = taps[i]; cc[i] = pk(c[i], c[i]); }
    __shared__ float2 xs[2048 + WIN];
    for (int i = threadIdx.x; i < 2048 + WIN; i += blockDim.x) xs[i] = xin[i & 1023];
    __syncthreads();
    float2 x[WIN]; u64 xp[WIN];
    float2 acc[R]; u64 ap[R];
#pragma unroll
    for (int r = 0; r < R; r++) { acc[r] = make_float2(0.f, 0.f); ap[r] = 0ull; }
    // conflict-free lane stride (17 float2 = 34 words) like the real kernel's padded sample tile
    int base = (threadIdx.x & 31) * 17 + (threadIdx.x >> 5) * 64;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        // a fresh sample window every pass: products are never loop-invariant
        int off = (base + it * TAPS) & 1023;
#pragma unroll
        for (int i = 0; i < WIN; i++) { x[i] = xs[off + i]; xp[i] = pk(x[i].x, x[i].y); }
#pragma unroll
        for (int i = 0; i < TAPS; i++) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (V == 0) {
                    acc[r].x = __fadd_rn(acc[r].x, __fmul_rn(x[r + i].x, c[i]));
                    acc[r].y = __fadd_rn(acc[r].y, __fmul_rn(x[r + i].y, c[i]));
                } else if (V == 1) {
                    ap[r] = add2(ap[r], pk(__fmul_rn(x[r + i].x, c[i]), __fmul_rn(x[r + i].y, c[i])));
                } else if (V == 2) {
                    float pr, pi;
                    unpk(mul2(xp[r + i], cc[i]), pr, pi);
                    acc[r].x = __fadd_rn(acc[r].x, pr);
                    acc[r].y = __fadd_rn(acc[r].y, pi);
                } else if (V == 3) {
                    ap[r] = add2(ap[r], mul2ftz(xp[r + i], cc[i]));
                } else if (V == 4) {
                    ap[r] = fma2(xp[r + i], cc[i], ap[r]);
                } else {
                    acc[r].x = __fmaf_rn(x[r + i].x, c[i], acc[r].x);
                    acc[r].y = __fmaf_rn(x[r + i].y, c[i], acc[r].y);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        if (PACKED_ACC) unpk(ap[r], acc[r].x, acc[r].y);
        out[(size_t)t * R + r] = acc[r];
    }
    (void)PACKED_X;
}

template <int V>
static void run(const char* name, const float2* x, const float* taps, float2* out, int sms, double mhz) {
    const int threads = 256, ctas_per_sm = 4, iters = 2000;
    int grid = sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    pipe_kernel<V><<<grid, threads>>>(x, taps, out, 10);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        pipe_kernel<V><<<grid, threads>>>(x, taps, out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double tapops = (double)grid * threads * (double)iters * TAPS * R;   // complex tap updates
    double per_s = tapops / (best * 1e-3);
    printf("{\"variant\": \"%s\", \"ms\": %.3f, \"complex_tap_updates_per_s\": %.4e, \"per_clk_per_sm_at_%.0fMHz\": %.2f}\n",
           name, best, per_s, mhz, per_s / (mhz * 1e6) / sms);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double mhz = khz / 1000.0;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f}\n", p.name, sms, mhz);
    float2* x; float* taps; float2* out;
    cudaMalloc(&x, 1024 * sizeof(float2)); cudaMalloc(&taps, 64 * sizeof(float));
    cudaMalloc(&out, (size_t)sms * 4 * 256 * R * sizeof(float2));
    float2 hx[1024]; float ht[64];
    for (int i = 0; i < 1024; i++) hx[i] = make_float2(0.001f * (i % 97) - 0.04f, 0.002f * (i % 89) - 0.08f);
    for (int i = 0; i < 64; i++) ht[i] = 0.01f * (i % 13) - 0.05f;
    cudaMemcpy(x, hx, sizeof hx, cudaMemcpyHostToDevice); cudaMemcpy(taps, ht, sizeof ht, cudaMemcpyHostToDevice);
    run<0>("v0_fmul_fmul_fadd_fadd", x, taps, out, sms, mhz);
    run<1>("v1_fmul_fmul_fadd2", x, taps, out, sms, mhz);
    run<2>("v2_fmul2_fadd_fadd", x, taps, out, sms, mhz);
    run<3>("v3_fmul2ftz_fadd2", x, taps, out, sms, mhz);
    run<4>("v4_ffma2", x, taps, out, sms, mhz);
    run<5>("v5_ffma_ffma", x, taps, out, sms, mhz);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
