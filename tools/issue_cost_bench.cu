// issue_cost_bench.cu -- what does one extra instruction of type X cost a warp that otherwise streams packed FP32x2
// mul+add pairs (the exact FIR inner loop)?  Each warp runs the v3 loop of fp32_pipe_bench (FMUL2.FTZ + FADD2, 8x8 tap
// updates per pass, 15 LDS.64) plus NEX extra instructions of one type per pass on an independent dependency chain.
// cost = (t(NEX) - t(0)) * clock * 4 schedulers * SMs / (extra warp-instructions)  [scheduler cycles per warp-instruction]
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_cost_bench issue_cost_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2ftz(u64 a, u64 b) { u64 r; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
#define R 8
#define TAPS 8
#define WIN (R + TAPS - 1)
enum { X_NONE, X_FADD, X_FMUL, X_DMUL, X_F2F_UP, X_F2F_DOWN, X_IADD, X_LDS, X_STS, X_FSETP_SEL, X_GAIN, X_I2F, X_MUFU };

template <int X, int NEX>
__global__ void __launch_bounds__(256) k(const float2* __restrict__ xin, const float* __restrict__ taps, float2* out, int iters, float seed) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    u64 cc[TAPS];
#pragma unroll
    for (int i = 0; i < TAPS; i++) cc[i] = pk(taps[i], taps[i]);
    __shared__ float2 xs[2048 + WIN];
    __shared__ float sink[256];
    for (int i = threadIdx.x; i < 2048 + WIN; i += blockDim.x) xs[i] = xin[i & 1023];
    __syncthreads();
    u64 xp[WIN], ap[R];
#pragma unroll
    for (int r = 0; r < R; r++) ap[r] = 0ull;
    int base = (threadIdx.x & 31) * 17 + (threadIdx.x >> 5) * 64;
    float e0 = seed, e1 = seed * 0.5f, e2 = seed + 1.f, e3 = seed - 1.f;     // four independent chains
    double d0 = seed, d1 = seed + 2.0;
    int i0 = (int)seed, i1 = 3;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        int off = (base + it * TAPS) & 1023;
#pragma unroll
        for (int i = 0; i < WIN; i++) { float2 v = xs[off + i]; xp[i] = pk(v.x, v.y); }
#pragma unroll
        for (int i = 0; i < TAPS; i++) {
#pragma unroll
            for (int r = 0; r < R; r++) ap[r] = add2(ap[r], mul2ftz(xp[r + i], cc[i]));
            // NEX extra instructions spread over the pass
#pragma unroll
            for (int q = 0; q < NEX / TAPS; q++) {
                float& e = (q & 3) == 0 ? e0 : (q & 3) == 1 ? e1 : (q & 3) == 2 ? e2 : e3;
                if (X == X_FADD) e = __fadd_rn(e, 1.25f);
                else if (X == X_FMUL) e = __fmul_rn(e, 1.0000001f);
                else if (X == X_DMUL) { double& d = (q & 1) ? d1 : d0; d = __dmul_rn(d, 1.0000000001); }
                else if (X == X_F2F_UP) { double& d = (q & 1) ? d1 : d0; d = (double)e; e = __fadd_rn(e, 1.0f); }        // counts F2F + FADD
                else if (X == X_F2F_DOWN) { double& d = (q & 1) ? d1 : d0; e = (float)d; d = __dadd_rn(d, 1.0); }          // counts F2F + DADD
                else if (X == X_IADD) { int& ii = (q & 1) ? i1 : i0; ii = ii * 3 + 1; }
                else if (X == X_LDS) e = __fadd_rn(e, sink[(threadIdx.x + q + it) & 255]);                                  // LDS + FADD
                else if (X == X_STS) sink[(threadIdx.x + q) & 255] = e;
                else if (X == X_FSETP_SEL) e = (e > 0.5f) ? e0 : e1;
                else if (X == X_GAIN) e = (float)((double)e * 1.85);                                                       // F2F + DMUL + F2F
                else if (X == X_I2F) { int& ii = (q & 1) ? i1 : i0; e = (float)ii; ii += 1; }                                // I2F + IADD
                else if (X == X_MUFU) e = __frcp_rn(e);
            }
        }
    }
    float s = e0 + e1 + e2 + e3 + (float)(d0 + d1) + (float)(i0 + i1);
#pragma unroll
    for (int r = 0; r < R; r++) { float a, b; unpk(ap[r], a, b); out[(size_t)t * R + r] = make_float2(a + s, b); }
}

template <int X, int NEX>
static double run(const float2* x, const float* taps, float2* out, int sms) {
    const int threads = 256, ctas_per_sm = 4, iters = 2000;
    int grid = sms * ctas_per_sm;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<X, NEX><<<grid, threads>>>(x, taps, out, 10, 1.5f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(a); k<X, NEX><<<grid, threads>>>(x, taps, out, iters, 1.5f); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float2* x; float* taps; float2* out;
    cudaMalloc(&x, 1024 * sizeof(float2)); cudaMalloc(&taps, 64 * sizeof(float)); cudaMalloc(&out, (size_t)sms * 4 * 256 * R * sizeof(float2));
    float2 hx[1024]; float ht[64];
    for (int i = 0; i < 1024; i++) hx[i] = make_float2(0.001f * (i % 97) - 0.04f, 0.002f * (i % 89) - 0.08f);
    for (int i = 0; i < 64; i++) ht[i] = 0.01f * (i % 13) - 0.05f;
    cudaMemcpy(x, hx, sizeof hx, cudaMemcpyHostToDevice); cudaMemcpy(taps, ht, sizeof ht, cudaMemcpyHostToDevice);
    const double t0 = run<X_NONE, 0>(x, taps, out, sms);
    // warp-passes per scheduler: 4 CTAs x 8 warps / 4 schedulers x 2000 iterations
    const double passes = 4.0 * 8 / 4 * 2000;
    printf("{\"device\": \"%s\", \"baseline_ms\": %.3f, \"cycles_per_pass\": %.1f, \"note\": \"128 packed + 15 LDS.64 per pass\"}\n", p.name, t0,
           t0 * 1e-3 * khz * 1e3 / passes);
#define ROW(X, N, name, ninstr) { double t = run<X, N>(x, taps, out, sms); \
        printf("{\"extra\": \"%s\", \"per_pass\": %d, \"ms\": %.3f, \"sched_cycles_per_extra_group\": %.2f, \"instructions_per_group\": %d}\n", name, N, t, \
               (t - t0) * 1e-3 * khz * 1e3 / passes / N, ninstr); }
    ROW(X_FADD, 16, "FADD", 1)
    ROW(X_FMUL, 16, "FMUL", 1)
    ROW(X_IADD, 16, "IMAD", 1)
    ROW(X_FSETP_SEL, 16, "FSETP+FSEL", 2)
    ROW(X_DMUL, 16, "DMUL", 1)
    ROW(X_F2F_UP, 16, "F2F.F64.F32 + FADD", 2)
    ROW(X_F2F_DOWN, 16, "F2F.F32.F64 + DADD", 2)
    ROW(X_GAIN, 16, "gain: F2F + DMUL + F2F", 3)
    ROW(X_I2F, 16, "I2F + IADD", 2)
    ROW(X_MUFU, 16, "MUFU.RCP (+fixup)", 1)
    ROW(X_LDS, 16, "LDS + FADD", 2)
    ROW(X_STS, 16, "STS", 1)
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
