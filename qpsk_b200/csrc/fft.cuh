// fft.cuh -- K2: batched power-of-two FFT in shared memory with a fused |X|^2 / argmax epilogue.
//
// Replaces fft/fftn/ifft/ifftn of the reference (algorithms/fft.c:98-136: recursive radix-2,
// complex double, forward transform scaled by 1/n, inverse unscaled) for many bursts at once, in
// FP32 (north_star tolerance 1e-5), and adds the estimator the reference never wrote: the first
// strict maximum of |X[k]|^2 (mirroring the argmax idiom of qpsk.c:173-180).
//
// Stockham autosort, radix-16 stages for n >= 256 (radix 8 below; the remainder stage is radix 8, 4 or 2),
// 16 points per thread in registers: n = 256 crosses shared memory once, 4096 twice.  Stage 0 reads the burst straight from HBM (coalesced), the last stage leaves its
// outputs in registers for the magnitude/argmax reduction (warp shuffles, then one shared-memory
// hop across warps), so shared memory is crossed stages-1 times and HBM exactly once.
#pragma once

#include "common.cuh"

template <int LOG2N>
struct FftCfg {
    static constexpr int N = 1 << LOG2N;
    static constexpr int P = (N >= 256) ? 16 : ((N >= 8) ? 8 : N); // points per thread = largest radix
    static constexpr int TPF = N / P;                              // threads per transform
    static constexpr int THREADS = (TPF >= 128) ? TPF : 128;
    static constexpr int FPB = THREADS / TPF;                      // transforms per CTA pass
    static constexpr int PTS = FPB * N;
    static constexpr int SKEW_PTS = PTS + PTS / 16;                // float2 elements, one pad slot per 16: unit-stride and stride-8 accesses are conflict-free
    static constexpr int TW = (N >= 2) ? N / 2 : 1;
    static constexpr size_t SMEM = sizeof(float2) * SKEW_PTS + sizeof(float2) * TW + sizeof(float) * 64 + sizeof(int) * 64;
};

// tolerance-mode arithmetic (1e-5): explicit fused multiply-adds, since the TU is built with -fmad=false
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
    return make_float2(__fmaf_rn(a.x, b.x, -(a.y * b.y)), __fmaf_rn(a.x, b.y, a.y * b.x));
}

// forward R-point DFT in registers, natural order in and out
template <int R>
__device__ __forceinline__ void dft_small(float2 (&v)[R]);

template <>
__device__ __forceinline__ void dft_small<2>(float2 (&v)[2]) {
    const float2 a = v[0], b = v[1];
    v[0] = make_float2(a.x + b.x, a.y + b.y);
    v[1] = make_float2(a.x - b.x, a.y - b.y);
}
template <>
__device__ __forceinline__ void dft_small<4>(float2 (&v)[4]) {
    const float2 s02 = make_float2(v[0].x + v[2].x, v[0].y + v[2].y), d02 = make_float2(v[0].x - v[2].x, v[0].y - v[2].y);
    const float2 s13 = make_float2(v[1].x + v[3].x, v[1].y + v[3].y), d13 = make_float2(v[1].x - v[3].x, v[1].y - v[3].y);
    v[0] = make_float2(s02.x + s13.x, s02.y + s13.y);
    v[2] = make_float2(s02.x - s13.x, s02.y - s13.y);
    v[1] = make_float2(d02.x + d13.y, d02.y - d13.x);      // d02 + (-i) d13
    v[3] = make_float2(d02.x - d13.y, d02.y + d13.x);      // d02 + (+i) d13
}
template <>
__device__ __forceinline__ void dft_small<8>(float2 (&v)[8]) {
    const float h = 0.70710678118654752f;
    float2 e[4] = { v[0], v[2], v[4], v[6] }, o[4] = { v[1], v[3], v[5], v[7] };
    dft_small<4>(e);
    dft_small<4>(o);
    // o[q] *= exp(-2 pi i q / 8)
    const float2 o1 = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));
    const float2 o2 = make_float2(o[2].y, -o[2].x);
    const float2 o3 = make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));
    v[0] = make_float2(e[0].x + o[0].x, e[0].y + o[0].y);  v[4] = make_float2(e[0].x - o[0].x, e[0].y - o[0].y);
    v[1] = make_float2(e[1].x + o1.x, e[1].y + o1.y);      v[5] = make_float2(e[1].x - o1.x, e[1].y - o1.y);
    v[2] = make_float2(e[2].x + o2.x, e[2].y + o2.y);      v[6] = make_float2(e[2].x - o2.x, e[2].y - o2.y);
    v[3] = make_float2(e[3].x + o3.x, e[3].y + o3.y);      v[7] = make_float2(e[3].x - o3.x, e[3].y - o3.y);
}

template <>
__device__ __forceinline__ void dft_small<16>(float2 (&v)[16]) {
    // 16 = 4 x 4: four radix-4 transforms over stride-4 subsequences, twiddles W16^(q*r), four radix-4 across
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
    float2 a[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        float2 t[4] = { v[r], v[r + 4], v[r + 8], v[r + 12] };
        dft_small<4>(t);
#pragma unroll
        for (int q = 0; q < 4; q++) a[r][q] = t[q];
    }
    // a[r][q] *= W16^(r*q), W16 = exp(-2 pi i / 16)
    const float2 w[10] = { {1.f, 0.f}, {c1, -s1}, {h, -h}, {s1, -c1}, {0.f, -1.f}, {-s1, -c1}, {-h, -h}, {-c1, -s1}, {-1.f, 0.f}, {-c1, s1} };
#pragma unroll
    for (int r = 1; r < 4; r++)
#pragma unroll
        for (int q = 1; q < 4; q++) {
            const int e = r * q;                      // 1,2,3,2,4,6,3,6,9
            if (e == 4) a[r][q] = make_float2(a[r][q].y, -a[r][q].x);
            else a[r][q] = cmulf(a[r][q], w[e]);
        }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        float2 t[4] = { a[0][q], a[1][q], a[2][q], a[3][q] };
        dft_small<4>(t);
#pragma unroll
        for (int p = 0; p < 4; p++) v[q + 4 * p] = t[p];
    }
}

__device__ __forceinline__ int fft_skew(int i) { return i + (i >> 4); }

struct FftArgs {
    const float2* in;      // [nbursts][N]
    float2* spectrum;      // optional [nbursts][N]
    int* bin;              // optional [nbursts] argmax bin
    float* mag2;           // optional [nbursts] |X[bin]|^2 (after scaling)
    const float2* tw;      // [N/2] exp(-2 pi i t / N), host-computed in double
    int nbursts;
    float im_sign;         // +1 forward, -1 inverse (inverse = conj(FFT(conj(x))))
    float scale;           // 1/N forward (fft.c:105-107), 1 inverse (fft.c:122-128)
};

// one Stockham stage for the P points a thread owns: radix R, sub-transform length NS already done
template <int LOG2N, int R, int NS, bool FIRST, bool LAST>
__device__ __forceinline__ void fft_stage(float2 (&pts)[FftCfg<LOG2N>::P], float2* sdat, const float2* stw,
                                          const float2* gin, int j, int base, bool active, float imsgn) {
    using Cfg = FftCfg<LOG2N>;
    constexpr int N = Cfg::N, P = Cfg::P, TPF = Cfg::TPF, NB = P / R;   // NB butterflies per thread
#pragma unroll
    for (int t = 0; t < NB; t++) {
        const int jj = j + t * TPF;                  // butterfly index in [0, N/R)
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int idx = jj + r * (N / R);
            if (FIRST) { v[r] = active ? gin[idx] : make_float2(0.f, 0.f); v[r].y *= imsgn; }   // inverse = conj(FFT(conj x))
            else v[r] = sdat[fft_skew(base + idx)];
        }
        if (NS > 1) {
            const int k = jj % NS;
            const int ti = k * (N / (NS * R));       // w1 = exp(-2 pi i k / (NS*R))
            const float2 w1 = stw[ti];
            if (R == 2) {
                v[1] = cmulf(v[1], w1);
            } else if (R == 4) {
                const float2 w2 = stw[2 * ti], w3 = cmulf(w1, w2);
                v[1] = cmulf(v[1], w1); v[2] = cmulf(v[2], w2); v[3] = cmulf(v[3], w3);
            } else if (R == 8) {
                const float2 w2 = stw[2 * ti], w4 = stw[4 * ti];
                const float2 w3 = cmulf(w1, w2), w5 = cmulf(w4, w1), w6 = cmulf(w4, w2), w7 = cmulf(w4, w3);
                v[1] = cmulf(v[1], w1); v[2] = cmulf(v[2], w2); v[3] = cmulf(v[3], w3); v[4] = cmulf(v[4], w4);
                v[5] = cmulf(v[5], w5); v[6] = cmulf(v[6], w6); v[7] = cmulf(v[7], w7);
            } else {
                const float2 w2 = stw[2 * ti], w4 = stw[4 * ti], w8 = stw[8 * ti];
                const float2 w3 = cmulf(w1, w2), w5 = cmulf(w4, w1), w6 = cmulf(w4, w2), w7 = cmulf(w4, w3);
                v[1] = cmulf(v[1], w1); v[2] = cmulf(v[2], w2); v[3] = cmulf(v[3], w3); v[4] = cmulf(v[4], w4);
                v[5] = cmulf(v[5], w5); v[6] = cmulf(v[6], w6); v[7] = cmulf(v[7], w7); v[8] = cmulf(v[8], w8);
                v[9] = cmulf(v[9], cmulf(w8, w1)); v[10] = cmulf(v[10], cmulf(w8, w2)); v[11] = cmulf(v[11], cmulf(w8, w3));
                v[12] = cmulf(v[12], cmulf(w8, w4)); v[13] = cmulf(v[13], cmulf(w8, w5)); v[14] = cmulf(v[14], cmulf(w8, w6));
                v[15] = cmulf(v[15], cmulf(w8, w7));
            }
        }
        dft_small<R>(v);
#pragma unroll
        for (int q = 0; q < R; q++) pts[t * R + q] = v[q];
    }
    if (!LAST) {
        if (!FIRST) __syncthreads();                 // everyone has read this stage's inputs
#pragma unroll
        for (int t = 0; t < NB; t++) {
            const int jj = j + t * TPF;
            const int o = (jj / NS) * NS * R + (jj % NS);
#pragma unroll
            for (int q = 0; q < R; q++) {
                sdat[fft_skew(base + o + q * NS)] = pts[t * R + q];
            }
        }
        __syncthreads();
    }
}

template <int LOG2N, int NS, bool FIRST>
__device__ __forceinline__ void fft_stages(float2 (&pts)[FftCfg<LOG2N>::P], float2* sdat, const float2* stw,
                                           const float2* gin, int j, int base, bool active, float imsgn) {
    constexpr int N = FftCfg<LOG2N>::N;
    constexpr int REM = N / NS;
    constexpr int RMAX = FftCfg<LOG2N>::P;
    constexpr int R = (REM >= RMAX) ? RMAX : REM;
    constexpr bool LAST = (NS * R == N);
    fft_stage<LOG2N, R, NS, FIRST, LAST>(pts, sdat, stw, gin, j, base, active, imsgn);
    if constexpr (!LAST) fft_stages<LOG2N, NS * R, false>(pts, sdat, stw, gin, j, base, active, imsgn);
}

// output index of pts[i] after the last stage (radix RLAST, sub-transform length NSL = N / RLAST)
template <int LOG2N>
struct FftLast {
    static constexpr int last_ns() { int ns = 1; while ((1 << LOG2N) / ns > FftCfg<LOG2N>::P) ns *= FftCfg<LOG2N>::P; return ns; }
    static constexpr int NSL = last_ns();
    static constexpr int RLAST = (1 << LOG2N) / NSL;
};
template <int LOG2N>
__device__ __forceinline__ int fft_out_index(int j, int i) {
    constexpr int RLAST = FftLast<LOG2N>::RLAST, NSL = FftLast<LOG2N>::NSL;
    const int t = i / RLAST, q = i % RLAST;
    return (j + t * FftCfg<LOG2N>::TPF) + q * NSL;
}

template <int LOG2N>
__global__ void __launch_bounds__(FftCfg<LOG2N>::THREADS) fft_kernel(const FftArgs a) {
    using Cfg = FftCfg<LOG2N>;
    constexpr int N = Cfg::N, P = Cfg::P, TPF = Cfg::TPF, FPB = Cfg::FPB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sdat = reinterpret_cast<float2*>(smem_raw);
    float2* stw = sdat + Cfg::SKEW_PTS;
    float* red_mag = reinterpret_cast<float*>(stw + Cfg::TW);
    int* red_idx = reinterpret_cast<int*>(red_mag + 64);

    for (int i = threadIdx.x; i < Cfg::TW; i += blockDim.x) stw[i] = a.tw[i];
    __syncthreads();

    const int fl = threadIdx.x / TPF, j = threadIdx.x % TPF;
    const int base = fl * N;
    for (int b0 = blockIdx.x * FPB; b0 < a.nbursts; b0 += gridDim.x * FPB) {
        const int b = b0 + fl;
        const bool active = b < a.nbursts;
        const float2* gin = a.in + (size_t)(active ? b : 0) * N;
        float2 pts[P];
        fft_stages<LOG2N, 1, true>(pts, sdat, stw, gin, j, base, active, a.im_sign);
        // ---- epilogue: scale, optional spectrum store, |X|^2 argmax
        float best = -1.0f;
        int besti = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < P; i++) {
            const int idx = fft_out_index<LOG2N>(j, i);
            const float2 x = make_float2(pts[i].x * a.scale, pts[i].y * a.scale * a.im_sign);
            if (a.spectrum != nullptr && active) a.spectrum[(size_t)b * N + idx] = x;
            const float m = x.x * x.x + x.y * x.y;
            if (m > best || (m == best && idx < besti)) { best = m; besti = idx; }
        }
        if (a.bin != nullptr) {
            // reduce over the TPF threads of this transform: first strict maximum = largest value, lowest index on ties
            constexpr int W = (TPF < 32) ? TPF : 32;
#pragma unroll
            for (int off = W / 2; off > 0; off >>= 1) {
                const float om = __shfl_down_sync(0xffffffffu, best, off, W);
                const int oi = __shfl_down_sync(0xffffffffu, besti, off, W);
                if (om > best || (om == best && oi < besti)) { best = om; besti = oi; }
            }
            if (TPF <= 32) {
                if (j == 0 && active) { a.bin[b] = besti; if (a.mag2) a.mag2[b] = best; }
            } else {
                constexpr int WPF = TPF / 32;            // warps per transform (FPB == 1 whenever TPF > 256; else FPB*WPF == 8)
                const int wib = threadIdx.x >> 5;
                if ((threadIdx.x & 31) == 0) { red_mag[wib] = best; red_idx[wib] = besti; }
                __syncthreads();
                if (j == 0 && active) {
                    const int w0 = fl * WPF;
                    for (int w = 1; w < WPF; w++) {
                        const float om = red_mag[w0 + w];
                        const int oi = red_idx[w0 + w];
                        if (om > best || (om == best && oi < besti)) { best = om; besti = oi; }
                    }
                    a.bin[b] = besti;
                    if (a.mag2) a.mag2[b] = best;
                }
                __syncthreads();
            }
        }
        // the next pass's stage-0 writes must not race with this pass's last-stage reads: the last
        // stage read shared memory before its registers were final, and every thread passed the
        // barrier inside the previous stage's write phase; one more barrier closes the loop
        __syncthreads();
    }
}
