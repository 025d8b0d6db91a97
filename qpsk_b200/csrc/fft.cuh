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

// points per thread = largest radix: 32 where it saves a pass through shared memory (512 = 32 x 16, 1024 = 32 x 32 inside
// one warp, 4096 = 32 x 32 x 4, 8192 = 32 x 32 x 8).  2048 stays 16 x 16 x 8: measured 3.6 % faster than 32 x 8 x 8 and 4 %
// faster than 32 x 32 x 2 (profiles/r01_notes.md).
__host__ __device__ constexpr int qpsk_fft_points_per_thread(int n) {
    return (n >= 512 && n != 2048) ? 32 : ((n >= 256) ? 16 : ((n >= 8) ? 8 : n));
}
// radix of the stage that still has `rem` points to combine: the largest one, except that 2 P is split evenly
// (P/4 x P/4 for P = 32) instead of ending in a radix-2 pass
__host__ __device__ constexpr int qpsk_fft_radix(int rem, int p) {
    return (p >= 32 && rem == 2 * p) ? p / 4 : (rem >= p ? p : rem);
}
#ifndef QPSK_FFT_PREFETCH_MIN_N
#define QPSK_FFT_PREFETCH_MIN_N 1024
#endif

template <int LOG2N>
struct FftCfg {
    static constexpr int N = 1 << LOG2N;
    static constexpr int P = qpsk_fft_points_per_thread(N);
    static constexpr int TPF = N / P;                              // threads per transform
    static constexpr int THREADS = (TPF >= 128) ? TPF : 128;
    static constexpr int FPB = THREADS / TPF;                      // transforms per CTA pass
    static constexpr int PTS = FPB * N;
    static constexpr int SKEW = (P >= 32) ? 32 : 16;               // one pad slot per SKEW points: unit-stride and stride-P accesses are conflict-free
    static constexpr int SKEW_PTS = PTS + PTS / SKEW;              // float2 elements
    // per-stage twiddle tables: stage (NS, R) holds exp(-2 pi i m k / (NS R)) for m = 1..R-1, k < NS, laid out [m-1][k]
    // so that the lanes of a warp (consecutive k) read consecutive words -- no bank conflicts, no products to form
    static constexpr int tw_count() { int ns = 1, tot = 0; while (ns < N) { int rem = N / ns; int r = qpsk_fft_radix(rem, P); if (ns > 1) tot += (r - 1) * ns; ns *= r; } return tot > 0 ? tot : 1; }
    static constexpr int TW = tw_count();
    // transforms of n >= PREFETCH_MIN_N points are staged: while one pass is being transformed the next pass's input
    // lands in a second buffer through cp.async, so a CTA that fills its SM (512 threads x 128 registers at n = 8192)
    // no longer idles through every load
    static constexpr bool PREFETCH = (N >= QPSK_FFT_PREFETCH_MIN_N);
    static constexpr size_t SMEM = sizeof(float2) * SKEW_PTS + sizeof(float2) * TW + sizeof(float) * 64 + sizeof(int) * 64
                                 + (PREFETCH ? sizeof(float2) * PTS : 0);
};

// Tolerance-mode arithmetic (1e-5) on packed FP32 pairs: one complex value per 64-bit register.  FADD2 takes
// per-operand swizzle/negate modifiers, so a +-i rotation folded into an add costs nothing, and a complex
// multiply is FMUL2 + FFMA2 (scalar-broadcast and swapped operands are modifiers too) -- half the issue slots
// of the scalar form, which is what this kernel is short of (profiles/r01_notes.md).
typedef u64 c64;
__device__ __forceinline__ c64 cadd(c64 a, c64 b) { return add2(a, b); }
__device__ __forceinline__ c64 csub(c64 a, c64 b) { c64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// a + (-i) b  and  a - (-i) b = a + i b ;  (-i)(x + iy) = y - ix
__device__ __forceinline__ c64 cadd_mi(c64 a, c64 b) { float x, y; unpack2(b, x, y); return add2(a, pack2(y, -x)); }
__device__ __forceinline__ c64 cadd_pi(c64 a, c64 b) { float x, y; unpack2(b, x, y); return add2(a, pack2(-y, x)); }
__device__ __forceinline__ c64 cmul(c64 a, c64 w) {          // (ax wx - ay wy, ax wy + ay wx)
    float ax, ay, wx, wy, tx, ty;
    unpack2(a, ax, ay);
    unpack2(w, wx, wy);
    unpack2(mul2(pack2(ay, ay), pack2(wy, wx)), tx, ty);
    return fma2(pack2(ax, ax), w, pack2(-tx, ty));
}
__device__ __forceinline__ c64 cmul_c(c64 a, float wx, float wy) { return cmul(a, pack2(wx, wy)); }
__device__ __forceinline__ c64 cscale(c64 a, float s) { return mul2(a, pack2(s, s)); }
__device__ __forceinline__ c64 cconj_if(c64 a, float sgn) { float x, y; unpack2(a, x, y); return pack2(x, y * sgn); }

// forward R-point DFT in registers, natural order in and out
template <int R>
__device__ __forceinline__ void dft_small(c64 (&v)[R]);

template <>
__device__ __forceinline__ void dft_small<2>(c64 (&v)[2]) {
    const c64 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}
template <>
__device__ __forceinline__ void dft_small<4>(c64 (&v)[4]) {
    const c64 s02 = cadd(v[0], v[2]), d02 = csub(v[0], v[2]);
    const c64 s13 = cadd(v[1], v[3]), d13 = csub(v[1], v[3]);
    v[0] = cadd(s02, s13);
    v[2] = csub(s02, s13);
    v[1] = cadd_mi(d02, d13);      // d02 + (-i) d13
    v[3] = cadd_pi(d02, d13);      // d02 + (+i) d13
}
template <>
__device__ __forceinline__ void dft_small<8>(c64 (&v)[8]) {
    const float h = 0.70710678118654752f;
    c64 e[4] = { v[0], v[2], v[4], v[6] }, o[4] = { v[1], v[3], v[5], v[7] };
    dft_small<4>(e);
    dft_small<4>(o);
    // o[q] *= exp(-2 pi i q / 8)
    const c64 o1 = cmul_c(o[1], h, -h);
    const c64 o3 = cmul_c(o[3], -h, -h);
    v[0] = cadd(e[0], o[0]);     v[4] = csub(e[0], o[0]);
    v[1] = cadd(e[1], o1);       v[5] = csub(e[1], o1);
    v[2] = cadd_mi(e[2], o[2]);  v[6] = cadd_pi(e[2], o[2]);
    v[3] = cadd(e[3], o3);       v[7] = csub(e[3], o3);
}
template <>
__device__ __forceinline__ void dft_small<16>(c64 (&v)[16]) {
    // 16 = 4 x 4: four radix-4 transforms over stride-4 subsequences, twiddles W16^(q*r), four radix-4 across
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
    c64 a[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        c64 t[4] = { v[r], v[r + 4], v[r + 8], v[r + 12] };
        dft_small<4>(t);
#pragma unroll
        for (int q = 0; q < 4; q++) a[r][q] = t[q];
    }
    // a[r][q] *= W16^(r*q), W16 = exp(-2 pi i / 16); the e = 4 case (-i) is folded into the second layer
    a[1][1] = cmul_c(a[1][1], c1, -s1);  a[1][2] = cmul_c(a[1][2], h, -h);    a[1][3] = cmul_c(a[1][3], s1, -c1);
    a[2][1] = cmul_c(a[2][1], h, -h);    /* a[2][2] *= -i below */           a[2][3] = cmul_c(a[2][3], -h, -h);
    a[3][1] = cmul_c(a[3][1], s1, -c1);  a[3][2] = cmul_c(a[3][2], -h, -h);   a[3][3] = cmul_c(a[3][3], -c1, s1);
    {
        float x, y;
        unpack2(a[2][2], x, y);
        a[2][2] = pack2(y, -x);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        c64 t[4] = { a[0][q], a[1][q], a[2][q], a[3][q] };
        dft_small<4>(t);
#pragma unroll
        for (int p = 0; p < 4; p++) v[q + 4 * p] = t[p];
    }
}

template <>
__device__ __forceinline__ void dft_small<32>(c64 (&v)[32]) {
    // 32 = 2 x 16, decimation in time: X[k] = E[k] + W32^k O[k], X[k+16] = E[k] - W32^k O[k]
    const float c1 = 0.98078528040323043f, s1 = 0.19509032201612825f, c2 = 0.92387953251128674f, s2 = 0.38268343236508977f,
                c3 = 0.83146961230254524f, s3 = 0.55557023301960218f, h = 0.70710678118654752f;
    c64 e[16], o[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { e[k] = v[2 * k]; o[k] = v[2 * k + 1]; }
    dft_small<16>(e);
    dft_small<16>(o);
    o[1] = cmul_c(o[1], c1, -s1);    o[2] = cmul_c(o[2], c2, -s2);    o[3] = cmul_c(o[3], c3, -s3);    o[4] = cmul_c(o[4], h, -h);
    o[5] = cmul_c(o[5], s3, -c3);    o[6] = cmul_c(o[6], s2, -c2);    o[7] = cmul_c(o[7], s1, -c1);    /* o[8] *= -i below */
    o[9] = cmul_c(o[9], -s1, -c1);   o[10] = cmul_c(o[10], -s2, -c2); o[11] = cmul_c(o[11], -s3, -c3); o[12] = cmul_c(o[12], -h, -h);
    o[13] = cmul_c(o[13], -c3, -s3); o[14] = cmul_c(o[14], -c2, -s2); o[15] = cmul_c(o[15], -c1, -s1);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k == 8) { v[k] = cadd_mi(e[k], o[k]); v[k + 16] = cadd_pi(e[k], o[k]); }
        else { v[k] = cadd(e[k], o[k]); v[k + 16] = csub(e[k], o[k]); }
    }
}

template <int SKEW>
__device__ __forceinline__ int fft_skew(int i) { return i + i / SKEW; }

// the threads of one transform exchange data between passes: a warp-level barrier is enough when a transform lives
// inside one warp (n / points-per-thread <= 32)
template <int TPF>
__device__ __forceinline__ void fft_sync() {
    if (TPF <= 32) __syncwarp();
    else __syncthreads();
}

struct FftArgs {
    const float2* in;      // [nbursts][N]
    float2* spectrum;      // optional [nbursts][N]
    int* bin;              // optional [nbursts] argmax bin
    float* mag2;           // optional [nbursts] |X[bin]|^2 (after scaling)
    const float2* tw;      // per-stage twiddle tables (FftCfg::TW entries), host-computed in double
    int nbursts;
    float im_sign;         // +1 forward, -1 inverse (inverse = conj(FFT(conj(x))))
    float scale;           // 1/N forward (fft.c:105-107), 1 inverse (fft.c:122-128)
};

// one Stockham stage for the P points a thread owns: radix R, sub-transform length NS already done
template <int LOG2N, int R, int NS, bool FIRST, bool LAST>
__device__ __forceinline__ void fft_stage(c64 (&pts)[FftCfg<LOG2N>::P], c64* sdat, const c64* stw,
                                          const float2* gin, int j, int base, bool active, float imsgn) {
    using Cfg = FftCfg<LOG2N>;
    constexpr int N = Cfg::N, P = Cfg::P, TPF = Cfg::TPF, NB = P / R;   // NB butterflies per thread
#pragma unroll
    for (int t = 0; t < NB; t++) {
        const int jj = j + t * TPF;                  // butterfly index in [0, N/R)
        c64 v[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int idx = jj + r * (N / R);
            if (FIRST && Cfg::PREFETCH) v[r] = cconj_if(reinterpret_cast<const c64*>(gin)[idx], imsgn);   // gin = this transform in the staging buffer (zero-filled when inactive)
            else if (FIRST) v[r] = cconj_if(active ? reinterpret_cast<const c64*>(gin)[idx] : 0ull, imsgn);   // inverse = conj(FFT(conj x))
            else v[r] = sdat[fft_skew<Cfg::SKEW>(base + idx)];
        }
        if (NS > 1) {
            const int k = jj % NS;
#pragma unroll
            for (int m = 1; m < R; m++) v[m] = cmul(v[m], stw[(m - 1) * NS + k]);
        }
        dft_small<R>(v);
#pragma unroll
        for (int q = 0; q < R; q++) pts[t * R + q] = v[q];
    }
    if (!LAST) {
        if (!FIRST) fft_sync<TPF>();                 // everyone has read this stage's inputs
#pragma unroll
        for (int t = 0; t < NB; t++) {
            const int jj = j + t * TPF;
            const int o = (jj / NS) * NS * R + (jj % NS);
#pragma unroll
            for (int q = 0; q < R; q++) {
                sdat[fft_skew<Cfg::SKEW>(base + o + q * NS)] = pts[t * R + q];
            }
        }
        fft_sync<TPF>();
    }
}

template <int LOG2N, int NS, bool FIRST>
__device__ __forceinline__ void fft_stages(c64 (&pts)[FftCfg<LOG2N>::P], c64* sdat, const c64* stw,
                                           const float2* gin, int j, int base, bool active, float imsgn) {
    constexpr int N = FftCfg<LOG2N>::N;
    constexpr int REM = N / NS;
    constexpr int RMAX = FftCfg<LOG2N>::P;
    constexpr int R = qpsk_fft_radix(REM, RMAX);
    constexpr bool LAST = (NS * R == N);
    fft_stage<LOG2N, R, NS, FIRST, LAST>(pts, sdat, stw, gin, j, base, active, imsgn);
    if constexpr (!LAST) fft_stages<LOG2N, NS * R, false>(pts, sdat, stw + (NS > 1 ? (R - 1) * NS : 0), gin, j, base, active, imsgn);
}

// output index of pts[i] after the last stage (radix RLAST, sub-transform length NSL = N / RLAST)
template <int LOG2N>
struct FftLast {
    static constexpr int last_ns() {
        int ns = 1;
        while (true) { const int rem = (1 << LOG2N) / ns, r = qpsk_fft_radix(rem, FftCfg<LOG2N>::P); if (ns * r == (1 << LOG2N)) return ns; ns *= r; }
    }
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
    c64* sdat = reinterpret_cast<c64*>(smem_raw);
    c64* stw = sdat + Cfg::SKEW_PTS;
    float* red_mag = reinterpret_cast<float*>(stw + Cfg::TW);
    int* red_idx = reinterpret_cast<int*>(red_mag + 64);

    for (int i = threadIdx.x; i < Cfg::TW; i += blockDim.x) stw[i] = reinterpret_cast<const c64*>(a.tw)[i];
    __syncthreads();

    const int fl = threadIdx.x / TPF, j = threadIdx.x % TPF;
    const int base = fl * N;
    c64* stage = reinterpret_cast<c64*>(red_idx + 64);      // [FPB][N], only with Cfg::PREFETCH
    // the threads of a transform fetch that transform: N/2 16-byte pieces over TPF threads = P/2 pieces each, so the
    // staging buffer of a transform is only ever touched by its own threads (a warp-level barrier orders it when TPF <= 32)
    auto prefetch = [&](int pass_b0) {
        const int pb = pass_b0 + fl;
        const bool in_range = pb < a.nbursts;
        const float2* src = a.in + (in_range ? (size_t)pb * N : 0);
#pragma unroll
        for (int i = 0; i < P / 2; i++) {
            const int piece = j + i * TPF;                                // 2 points each
            const unsigned d = (unsigned)__cvta_generic_to_shared(stage + base + 2 * piece);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src + (in_range ? 2 * piece : 0)), "r"(in_range ? 16 : 0) : "memory");
        }
    };
    if (Cfg::PREFETCH && blockIdx.x * FPB < a.nbursts) prefetch(blockIdx.x * FPB);
    for (int b0 = blockIdx.x * FPB; b0 < a.nbursts; b0 += gridDim.x * FPB) {
        const int b = b0 + fl;
        const bool active = b < a.nbursts;
        c64 pts[P];
        if constexpr (Cfg::PREFETCH) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            fft_sync<TPF>();                                             // this transform's input has landed for all its threads
            constexpr int R0 = P;                                        // N >= P * P here: the first stage is a full radix-P one
            fft_stage<LOG2N, R0, 1, true, false>(pts, sdat, stw, reinterpret_cast<const float2*>(stage + base), j, base, active, a.im_sign);
            // the stage ended with a barrier after everyone's reads of the staging buffer: refill it
            if (b0 + gridDim.x * FPB < a.nbursts) prefetch(b0 + gridDim.x * FPB);
            fft_stages<LOG2N, R0, false>(pts, sdat, stw, nullptr, j, base, active, a.im_sign);
        } else {
            const float2* gin = a.in + (size_t)(active ? b : 0) * N;
            fft_stages<LOG2N, 1, true>(pts, sdat, stw, gin, j, base, active, a.im_sign);
        }
        // ---- epilogue: scale, optional spectrum store, |X|^2 argmax
        float best = -1.0f;
        int besti = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < P; i++) {
            const int idx = fft_out_index<LOG2N>(j, i);
            float2 x;
            unpack2(pts[i], x.x, x.y);
            x.x *= a.scale;
            x.y *= a.scale * a.im_sign;
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
        fft_sync<TPF>();
    }
}
