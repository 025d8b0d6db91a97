// fft.cuh -- K2: batched power-of-two FFT in shared memory with a fused |X|^2 / argmax epilogue.
//
// Replaces fft/fftn/ifft/ifftn of the reference (algorithms/fft.c:98-136: recursive radix-2,
// complex double, forward transform scaled by 1/n, inverse unscaled) for many bursts at once, in
// FP32 (north_star tolerance 1e-5), and adds the estimator the reference never wrote: the first
// strict maximum of |X[k]|^2 (mirroring the argmax idiom of qpsk.c:173-180).
//
// Stockham autosort, 16 or 32 points per thread in registers (radix-16 / radix-32 stages, the
// remainder stage is radix 16, 8, 4 or 2): n = 256 crosses shared memory once, 4096 and 8192 twice.
// Stage 0 reads the burst straight from HBM (coalesced), the last stage leaves its outputs in
// registers for the magnitude/argmax reduction (warp shuffles, then one shared-memory hop across
// warps), so HBM is crossed exactly once.
//
// The kernel is issue-bound (profiles/r01_notes.md: ~68 instructions per point in round 1, of which
// 19 are the packed butterflies), so round 2 is about the other 49:
//   * every shared-memory access is `per-thread base + compile-time offset`: the skew (one pad slot per
//     SKEW points) is linear in the unrolled indices, see fft_off();
//   * the estimator entry point (forward, argmax only) is its own instantiation: no conjugation, no
//     scaling, no spectrum store, |X|^2 as one FMUL2 + one FADD per point, the maximum as an FMNMX tree
//     and the bin recovered only by the thread(s) that hold the maximum;
//   * the twiddles of a remainder stage are factored, w^(m (j + t TPF)) = w^(m j) W_P^(m t) with the second
//     factor a compile-time constant, which shrinks the largest table from ~n to ~n/P entries: 8192 points
//     fit two CTAs per SM and 4096 four, without the staging buffer of round 1.
#pragma once

#include "common.cuh"

// points per thread = largest radix: 32 where it saves a pass through shared memory (512 = 32 x 16, 1024 = 32 x 32 inside
// one warp).  2048, 4096 and 8192 take 64 points per thread (2048 = 64 x 32 inside one warp, 4096 = 64 x 64, 8192 = 64 x 64 x 2):
// one crossing of shared memory instead of two for 2048 and 4096, and every thread has its 64 loads in flight at once.
#ifndef QPSK_FFT_P2048
#define QPSK_FFT_P2048 64
#endif
#ifndef QPSK_FFT_P4096
#define QPSK_FFT_P4096 64
#endif
#ifndef QPSK_FFT_P8192
#define QPSK_FFT_P8192 64
#endif
__host__ __device__ constexpr int qpsk_fft_points_per_thread(int n) {
    return n >= 8192 ? QPSK_FFT_P8192 : (n >= 4096 ? QPSK_FFT_P4096 : (n == 2048 ? QPSK_FFT_P2048 : ((n >= 512) ? 32 : ((n >= 256) ? 16 : ((n >= 8) ? 8 : n)))));
}
// radix of the stage that still has `rem` points to combine: the largest one, except that 2 P is split evenly
// (P/4 x P/4) for P = 32 instead of ending in a radix-2 pass; P = 64 keeps 64 x 2 (the radix-2 pass is the cheap one)
__host__ __device__ constexpr int qpsk_fft_radix(int rem, int p) {
    return (p == 32 && rem == 2 * p) ? p / 4 : (rem >= p ? p : rem);
}
// entries per twiddle index m of the stage (ns, r) of an n-point transform with tpf threads of p points: a remainder
// stage whose sub-transform length exceeds the thread count keeps only the first tpf columns (the rest are constant
// multiples: powers of W32, so only for p <= 32; p = 64 keeps whole rows)
__host__ __device__ constexpr int qpsk_fft_tw_cols(int ns, int tpf, int p) { return (ns < tpf || p > 32) ? ns : tpf; }
// p = 64: the twiddles of the first crossing are applied by the producer, to the outputs of stage 0 as they are stored
// (output q of butterfly jj, which stage 1 reads as input r' = jj R R' / n of its butterfly q, takes w^(q r')): the loads
// and multiplies then overlap the tail of the first 64-point DFT instead of sitting between the barrier and the second.
// That stage's table is laid out [q - 1][r'], (ns - 1) x r entries.
#ifndef QPSK_FFT_TW_ON_STORE
#define QPSK_FFT_TW_ON_STORE 0      // measured slower on B200 (4096: 0.60 -> 0.55 of HBM; registers hit the cap): off
#endif
__host__ __device__ constexpr bool qpsk_fft_tw_on_store(int ns, int p) { return QPSK_FFT_TW_ON_STORE && p >= 64 && ns == p; }
// p = 64: the 63 twiddles a thread needs for a radix-64 stage, z^m with z = w^k, come from 14 table entries, z^(8a) and
// z^b (a, b = 1..7), as z^(8a + b) = z^(8a) z^b: the whole 64 x 64 table (32 KB) does not fit what six resident CTAs
// leave of the L1, and 63 L2 round trips per thread and burst were the kernel's largest stall; 14 x 64 entries (7 KB) do fit.
__host__ __device__ constexpr bool qpsk_fft_tw_two_level(int ns, int r, int p) { return p >= 64 && ns == 64 && r == 64 && !qpsk_fft_tw_on_store(ns, p); }
// ... and the single twiddle of a final radix-2 stage, w^(j + t tpf), is w^j times a compile-time power of W64
#ifndef QPSK_FFT_FACT64
#define QPSK_FFT_FACT64 1
#endif
__host__ __device__ constexpr bool qpsk_fft_tw_fact64(int ns, int r, int tpf, int p) { return QPSK_FFT_FACT64 && p >= 64 && r == 2 && ns > tpf && (2 * ns) / tpf == 64; }
__host__ __device__ constexpr int qpsk_fft_tw_entries(int ns, int r, int tpf, int p) {
    return ns <= 1 ? 0 : (qpsk_fft_tw_on_store(ns, p) ? (ns - 1) * r : (qpsk_fft_tw_two_level(ns, r, p) ? 14 * 64
                        : (qpsk_fft_tw_fact64(ns, r, tpf, p) ? tpf : (r - 1) * qpsk_fft_tw_cols(ns, tpf, p))));
}
// n = 8192 with 64 points per thread is two 4096-point transforms side by side in one CTA (decimation in frequency:
// even bins = FFT(x[n] + x[n + 4096]), odd bins = FFT((x[n] - x[n + 4096]) w^n)), each with ONE crossing of shared memory,
// instead of 64 x 64 x 2 with two.  The radix-2 step happens as stage 0 reads the burst; of the odd half's factor
// w^n = w^j W128^r (n = j + 64 r) the constant part multiplies the inputs and w^j moves into the second stage's twiddle
// base, z = w8192^(2k + 1) instead of w8192^(2k): a second 14 x 64 table.
#ifndef QPSK_FFT_SPLIT8192
#define QPSK_FFT_SPLIT8192 1
#endif
__host__ __device__ constexpr bool qpsk_fft_split(int n) { return QPSK_FFT_SPLIT8192 && n == 8192 && qpsk_fft_points_per_thread(n) == 64 && qpsk_fft_points_per_thread(4096) == 64; }
__host__ __device__ constexpr int qpsk_fft_tw_count(int n) {
    if (qpsk_fft_split(n)) return 2 * 14 * 64;
    const int p = qpsk_fft_points_per_thread(n), tpf = n / p;
    int ns = 1, tot = 0;
    while (ns < n) {
        const int r = qpsk_fft_radix(n / ns, p);
        tot += qpsk_fft_tw_entries(ns, r, tpf, p);
        ns *= r;
    }
    return tot > 0 ? tot : 1;
}

// Per-stage twiddles exp(-2 pi i m k / (ns r)), m = 1..r-1, k < qpsk_fft_tw_cols(ns, tpf), laid out [m-1][k] so that
// the lanes of a warp (consecutive k) read consecutive words; evaluated in double on the host (fft.c:55-56) and
// rounded once.  `tw` holds qpsk_fft_tw_count(n) entries.
inline void qpsk_fft_make_twiddles(int n, float2* tw) {
    const int p = qpsk_fft_points_per_thread(n), tpf = n / p;
    tw[0] = make_float2(1.0f, 0.0f);
    int pos = 0;
    if (qpsk_fft_split(n)) {
        for (int half = 0; half < 2; half++)               // z = w_n^(2k + half), entries z^(8a) (a = 1..7) then z^b (b = 1..7)
            for (int i = 0; i < 14; i++)
                for (int k = 0; k < 64; k++) {
                    const int e = (i < 7) ? 8 * (i + 1) : (i - 6);
                    const double ang = 2.0 * 3.14159265358979323846 * (double)e * (double)(2 * k + half) / (double)n;
                    tw[pos++] = make_float2((float)cos(ang), (float)(-sin(ang)));
                }
        return;
    }
    for (int ns = 1; ns < n;) {
        const int r = qpsk_fft_radix(n / ns, p);
        if (qpsk_fft_tw_on_store(ns, p)) {
            for (int q = 1; q < ns; q++)
                for (int k = 0; k < r; k++) {
                    const double ang = 2.0 * 3.14159265358979323846 * (double)q * (double)k / ((double)ns * (double)r);
                    tw[pos++] = make_float2((float)cos(ang), (float)(-sin(ang)));
                }
        } else if (qpsk_fft_tw_two_level(ns, r, p)) {
            for (int i = 0; i < 14; i++)
                for (int k = 0; k < 64; k++) {
                    const int e = (i < 7) ? 8 * (i + 1) : (i - 6);
                    const double ang = 2.0 * 3.14159265358979323846 * (double)e * (double)k / ((double)ns * (double)r);
                    tw[pos++] = make_float2((float)cos(ang), (float)(-sin(ang)));
                }
        } else if (ns > 1) {
            const int cols = qpsk_fft_tw_fact64(ns, r, tpf, p) ? tpf : qpsk_fft_tw_cols(ns, tpf, p);
            for (int m = 1; m < r; m++)
                for (int k = 0; k < cols; k++) {
                    const double ang = 2.0 * 3.14159265358979323846 * (double)m * (double)k / ((double)ns * (double)r);
                    tw[pos++] = make_float2((float)cos(ang), (float)(-sin(ang)));
                }
        }
        ns *= r;
    }
}

// registers per thread each kernel family is capped at (through __launch_bounds__' resident-CTA argument);
// tools/fft_bench.cu sweeps them
#ifndef QPSK_FFT_REGS_P32
#define QPSK_FFT_REGS_P32 102
#endif
#ifndef QPSK_FFT_REGS_P16
#define QPSK_FFT_REGS_P16 72
#endif
#ifndef QPSK_FFT_REGS_P64
#define QPSK_FFT_REGS_P64 168
#endif
#ifndef QPSK_FFT_L2_AHEAD
#define QPSK_FFT_L2_AHEAD 1
#endif
#ifndef QPSK_FFT_FMA_SMALL
#define QPSK_FFT_FMA_SMALL 1     // radix-32 / radix-16 stages in the FMA form too (dft32_f / dft16_f)
#endif
#ifndef QPSK_FFT_TW_SMEM_MAX
#define QPSK_FFT_TW_SMEM_MAX 4096
#endif
#define QPSK_FFT_MINB(threads, p) ((p) >= 64 ? 65536 / ((threads) * QPSK_FFT_REGS_P64) : (p) >= 32 ? 65536 / ((threads) * QPSK_FFT_REGS_P32) : ((p) >= 16 ? 65536 / ((threads) * QPSK_FFT_REGS_P16) : 4))

template <int LOG2N>
struct FftCfg {
    static constexpr int N = 1 << LOG2N;
    static constexpr int P = qpsk_fft_points_per_thread(N);
    static constexpr int TPF = N / P;                              // threads per transform
    // a transform wider than a warp is a CTA of its own where that leaves >= 64 threads (its barriers are then plain
    // __syncthreads: barriers named per transform inside a larger CTA measured 25 % slower at n = 4096)
    static constexpr int THREADS = (P >= 64 && TPF >= 64) ? TPF : ((TPF >= 128) ? TPF : 128);
    static constexpr int FPB = THREADS / TPF;                      // transforms per CTA pass
    static constexpr int PTS = FPB * N;
    static constexpr int SKEW = (P >= 64) ? 64 : ((P >= 32) ? 32 : 16);   // one pad slot per SKEW points: unit-stride and stride-P accesses are conflict-free
    static constexpr int SKEW_PTS = PTS + PTS / SKEW;              // float2 elements
    // 64 points per thread: the burst comes in through the bulk-copy engine (one cp.async.bulk of the whole burst per
    // transform, completion on an mbarrier) into the transform's work buffer, unpadded -- stage 0 reads it with unit
    // stride across lanes, so it needs no skew, and overwrites it in the skewed layout afterwards.  The copy is issued
    // as soon as the previous burst's last stage has read the buffer, so it runs under that stage's arithmetic and the
    // epilogue, and no thread ever waits for HBM with its registers full.
#ifndef QPSK_FFT_TMA_MINP
#define QPSK_FFT_TMA_MINP 64
#endif
    static constexpr bool TMA_IN = (P >= QPSK_FFT_TMA_MINP);
    static constexpr int TW = qpsk_fft_tw_count(N);
    static constexpr bool LIN = (N >= 256);                        // linear skew offsets (static_asserted per stage)
    // resident CTAs per SM the register allocation is capped for
    static constexpr int MINB = QPSK_FFT_MINB(THREADS, P) > 0 ? QPSK_FFT_MINB(THREADS, P) : 1;
    // large twiddle tables stay in global memory and are read through L1 (same latency as shared memory, and they no
    // longer cost resident CTAs)
    static constexpr bool TW_SMEM = (TW <= QPSK_FFT_TW_SMEM_MAX) && (P < 64);
    static constexpr size_t SMEM = sizeof(float2) * SKEW_PTS + (TW_SMEM ? sizeof(float2) * TW : 0) + sizeof(float) * 64 + sizeof(int) * 64 + 8 * 8;
};

// Tolerance-mode arithmetic (1e-5) on packed FP32 pairs: one complex value per 64-bit register.  FADD2/FMUL2/FFMA2 take
// per-operand swap / negate / scalar-broadcast modifiers, so a +-i rotation folded into an add costs nothing and a
// complex multiply is exactly FMUL2 + FFMA2 (checked in SASS: no MOVs).
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
__device__ __forceinline__ c64 cconj_if(c64 a, float sgn) { float x, y; unpack2(a, x, y); return pack2(x, y * sgn); }

// cos / sin of 2 pi k / 32 as compile-time constants
__host__ __device__ constexpr float qpsk_cos32(int k) {
    constexpr float C[9] = { 1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f,
                             0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f, 0.0f };
    k &= 31;
    if (k > 16) k = 32 - k;
    return k <= 8 ? C[k] : -C[16 - k];
}
__host__ __device__ constexpr float qpsk_sin32(int k) { return qpsk_cos32(k - 8); }
// Every constant twiddle of the small DFTs is a power of W32 = exp(-2 pi i / 32), and every such power is +-1 times an
// optionally swapped copy of (c, s) or (c, -s), (c, s) = (cos, sin)(2 pi r / 32), r = 1..4.  Those eight pairs live in
// registers for the whole kernel (FftConsts); swap and a common sign fold into the operand modifiers of FMUL2 / FFMA2
// (a sign on ONE half of a broadcast operand does not, hence both (c, s) and (c, -s)), so a constant complex multiply is
// two instructions.  As immediates every use cost two extra MOVs to assemble the 64-bit operand (checked in SASS).
struct FftConsts {
    c64 bp[4];     // (c, s)
    c64 bm[4];     // (c, -s)
    c64 bp64[4];   // the same for the odd powers of W64: (cos, sin)(2 pi r / 64), r = 1, 3, 5, 7 (used by the 64-point DFT only)
    c64 bm64[4];
    c64 two;       // (2, 2): a - W b = 2 a - (a + W b)
};
// the host fills FftArgs::kbase with (c, s) x 4 then (c, -s) x 4; arriving as kernel arguments the pairs are whole
// 64-bit uniform-register operands, opaque to the optimiser (which would otherwise fold them back into immediates)
inline void fft_two_host(float2& t) { t = make_float2(2.0f, 2.0f); }
// Stage 0 of the split kernel computes the 64-point DFT of v[m] W128^(h m), m = n1 + 8 n2, h = 0 for the even-bin half
// and 1 for the odd-bin half, through ONE instruction stream (two streams measured 20 % slower: they halve the
// instruction cache's reach), so the factors are data: [h][n2] = W16^(h n2) rides in the first-level butterflies,
// [h][8 + 8 k2 + n1] = W64^(n1 k2) W128^(h n1) in the second level's.
inline void fft_wsplit_host(float2 (&w)[2][72]) {
    const double tau = 2.0 * 3.14159265358979323846;
    for (int h = 0; h < 2; h++) {
        for (int n2 = 0; n2 < 8; n2++) w[h][n2] = make_float2((float)cos(tau * h * n2 / 16.0), (float)(-sin(tau * h * n2 / 16.0)));
        for (int k2 = 0; k2 < 8; k2++)
            for (int n1 = 0; n1 < 8; n1++) {
                const double ang = tau * (double)(n1 * (2 * k2 + h)) / 128.0;
                w[h][8 + 8 * k2 + n1] = make_float2((float)cos(ang), (float)(-sin(ang)));
            }
    }
}
inline void fft_consts_host(float2 (&kb)[16]) {
    for (int r = 1; r <= 4; r++) {
        kb[r - 1] = make_float2(qpsk_cos32(r), qpsk_sin32(r));
        kb[4 + r - 1] = make_float2(qpsk_cos32(r), -qpsk_sin32(r));
        const double ang = 2.0 * 3.14159265358979323846 * (double)(2 * r - 1) / 64.0;
        kb[8 + r - 1] = make_float2((float)cos(ang), (float)sin(ang));
        kb[12 + r - 1] = make_float2((float)cos(ang), (float)(-sin(ang)));
    }
}
// a * W, W = SGN * (SW ? swap(base) : base)
template <int SW, int SGN>
__device__ __forceinline__ c64 cmul_base(c64 a, c64 base) {
    float ax, ay, bx, by, tx, ty;
    unpack2(a, ax, ay);
    unpack2(base, bx, by);
    const float u = SW ? by : bx, v = SW ? bx : by;              // W = SGN (u, v)
    unpack2(mul2(pack2(ay, ay), pack2(v, u)), tx, ty);           // (ay v, ay u)
    return fma2(pack2(SGN > 0 ? ax : -ax, SGN > 0 ? ax : -ax), pack2(u, v), pack2(SGN > 0 ? -tx : tx, SGN > 0 ? ty : -ty));
}
// v * W32^K, K a compile-time constant
template <int K>
__device__ __forceinline__ c64 cmul_w32(c64 v, const FftConsts& kc) {
    constexpr int k = K & 31, q = k / 8, r = k % 8;
    if constexpr (r == 0) {
        float x, y;
        unpack2(v, x, y);
        if constexpr (q == 0) return v;
        else if constexpr (q == 1) return pack2(y, -x);
        else if constexpr (q == 2) return pack2(-x, -y);
        else return pack2(-y, x);
    } else if constexpr (r <= 4) {                    // W32^r = (c, -s); times (-i)^q
        constexpr int i = r - 1;
        if constexpr (q == 0) return cmul_base<0, 1>(v, kc.bm[i]);         // ( c, -s)
        else if constexpr (q == 1) return cmul_base<1, -1>(v, kc.bp[i]);   // (-s, -c)
        else if constexpr (q == 2) return cmul_base<0, -1>(v, kc.bm[i]);   // (-c,  s)
        else return cmul_base<1, 1>(v, kc.bp[i]);                          // ( s,  c)
    } else {                                          // W32^r = (s', -c') with (c', s') the pair of 8 - r
        constexpr int i = 8 - r - 1;
        if constexpr (q == 0) return cmul_base<1, -1>(v, kc.bm[i]);        // ( s, -c)
        else if constexpr (q == 1) return cmul_base<0, -1>(v, kc.bp[i]);   // (-c, -s)
        else if constexpr (q == 2) return cmul_base<1, 1>(v, kc.bm[i]);    // (-s,  c)
        else return cmul_base<0, 1>(v, kc.bp[i]);                          // ( c,  s)
    }
}

// v * W64^K: even powers are powers of W32, odd ones use the second constant set the same way
template <int K>
__device__ __forceinline__ c64 cmul_w64(c64 v, const FftConsts& kc) {
    constexpr int k = K & 63;
    if constexpr ((k & 1) == 0) {
        return cmul_w32<k / 2>(v, kc);
    } else {
        constexpr int q = k / 16, r = k % 16;
        if constexpr (r < 8) {                        // W64^r = (c, -s); times (-i)^q
            constexpr int i = (r - 1) / 2;
            if constexpr (q == 0) return cmul_base<0, 1>(v, kc.bm64[i]);
            else if constexpr (q == 1) return cmul_base<1, -1>(v, kc.bp64[i]);
            else if constexpr (q == 2) return cmul_base<0, -1>(v, kc.bm64[i]);
            else return cmul_base<1, 1>(v, kc.bp64[i]);
        } else {                                      // W64^r = (s', -c') with (c', s') the pair of 16 - r
            constexpr int i = (16 - r - 1) / 2;
            if constexpr (q == 0) return cmul_base<1, -1>(v, kc.bm64[i]);
            else if constexpr (q == 1) return cmul_base<0, -1>(v, kc.bp64[i]);
            else if constexpr (q == 2) return cmul_base<1, 1>(v, kc.bm64[i]);
            else return cmul_base<0, 1>(v, kc.bp64[i]);
        }
    }
}

// forward R-point DFT in registers, natural order in and out
template <int R>
__device__ __forceinline__ void dft_small(c64 (&v)[R], const FftConsts& kc);

template <>
__device__ __forceinline__ void dft_small<2>(c64 (&v)[2], const FftConsts&) {
    const c64 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}
template <>
__device__ __forceinline__ void dft_small<4>(c64 (&v)[4], const FftConsts&) {
    const c64 s02 = cadd(v[0], v[2]), d02 = csub(v[0], v[2]);
    const c64 s13 = cadd(v[1], v[3]), d13 = csub(v[1], v[3]);
    v[0] = cadd(s02, s13);
    v[2] = csub(s02, s13);
    v[1] = cadd_mi(d02, d13);      // d02 + (-i) d13
    v[3] = cadd_pi(d02, d13);      // d02 + (+i) d13
}
template <>
__device__ __forceinline__ void dft_small<8>(c64 (&v)[8], const FftConsts& kc) {
    c64 e[4] = { v[0], v[2], v[4], v[6] }, o[4] = { v[1], v[3], v[5], v[7] };
    dft_small<4>(e, kc);
    dft_small<4>(o, kc);
    // o[q] *= exp(-2 pi i q / 8)
    const c64 o1 = cmul_w32<4>(o[1], kc);
    const c64 o3 = cmul_w32<12>(o[3], kc);
    v[0] = cadd(e[0], o[0]);     v[4] = csub(e[0], o[0]);
    v[1] = cadd(e[1], o1);       v[5] = csub(e[1], o1);
    v[2] = cadd_mi(e[2], o[2]);  v[6] = cadd_pi(e[2], o[2]);
    v[3] = cadd(e[3], o3);       v[7] = csub(e[3], o3);
}
template <>
__device__ __forceinline__ void dft_small<16>(c64 (&v)[16], const FftConsts& kc) {
    // 16 = 4 x 4: four radix-4 transforms over stride-4 subsequences, twiddles W16^(q*r), four radix-4 across
    c64 a[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        c64 t[4] = { v[r], v[r + 4], v[r + 8], v[r + 12] };
        dft_small<4>(t, kc);
#pragma unroll
        for (int q = 0; q < 4; q++) a[r][q] = t[q];
    }
    // a[r][q] *= W16^(r*q), W16 = exp(-2 pi i / 16); the e = 4 case (-i) is folded into the second layer
    a[1][1] = cmul_w32<2>(a[1][1], kc);  a[1][2] = cmul_w32<4>(a[1][2], kc);   a[1][3] = cmul_w32<6>(a[1][3], kc);
    a[2][1] = cmul_w32<4>(a[2][1], kc);  a[2][2] = cmul_w32<8>(a[2][2], kc);   a[2][3] = cmul_w32<12>(a[2][3], kc);
    a[3][1] = cmul_w32<6>(a[3][1], kc);  a[3][2] = cmul_w32<12>(a[3][2], kc);  a[3][3] = cmul_w32<18>(a[3][3], kc);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        c64 t[4] = { a[0][q], a[1][q], a[2][q], a[3][q] };
        dft_small<4>(t, kc);
#pragma unroll
        for (int p = 0; p < 4; p++) v[q + 4 * p] = t[p];
    }
}

template <>
__device__ __forceinline__ void dft_small<32>(c64 (&v)[32], const FftConsts& kc) {
    // 32 = 2 x 16, decimation in time: X[k] = E[k] + W32^k O[k], X[k+16] = E[k] - W32^k O[k]
    c64 e[16], o[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { e[k] = v[2 * k]; o[k] = v[2 * k + 1]; }
    dft_small<16>(e, kc);
    dft_small<16>(o, kc);
    o[1] = cmul_w32<1>(o[1], kc);    o[2] = cmul_w32<2>(o[2], kc);    o[3] = cmul_w32<3>(o[3], kc);    o[4] = cmul_w32<4>(o[4], kc);
    o[5] = cmul_w32<5>(o[5], kc);    o[6] = cmul_w32<6>(o[6], kc);    o[7] = cmul_w32<7>(o[7], kc);    /* o[8] *= -i below */
    o[9] = cmul_w32<9>(o[9], kc);    o[10] = cmul_w32<10>(o[10], kc); o[11] = cmul_w32<11>(o[11], kc); o[12] = cmul_w32<12>(o[12], kc);
    o[13] = cmul_w32<13>(o[13], kc); o[14] = cmul_w32<14>(o[14], kc); o[15] = cmul_w32<15>(o[15], kc);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k == 8) { v[k] = cadd_mi(e[k], o[k]); v[k + 16] = cadd_pi(e[k], o[k]); }
        else { v[k] = cadd(e[k], o[k]); v[k + 16] = csub(e[k], o[k]); }
    }
}

template <int N1, int K2>
__device__ __forceinline__ void dft64_twiddle_row(c64 (&a)[8][8], const FftConsts& kc) {
    if constexpr (K2 < 8) {
        a[N1][K2] = cmul_w64<N1 * K2>(a[N1][K2], kc);
        dft64_twiddle_row<N1, K2 + 1>(a, kc);
    }
}
template <int N1>
__device__ __forceinline__ void dft64_twiddle(c64 (&a)[8][8], const FftConsts& kc) {
    if constexpr (N1 < 8) {
        dft64_twiddle_row<N1, 1>(a, kc);
        dft64_twiddle<N1 + 1>(a, kc);
    }
}
template <>
__device__ __forceinline__ void dft_small<64>(c64 (&v)[64], const FftConsts& kc) {
    // 64 = 8 x 8 with n = n1 + 8 n2, k = 8 k1 + k2:  W64^(n k) = W8^(n2 k2) W64^(n1 k2) W8^(n1 k1)
    c64 a[8][8];
#pragma unroll
    for (int n1 = 0; n1 < 8; n1++) {
        c64 t[8];
#pragma unroll
        for (int n2 = 0; n2 < 8; n2++) t[n2] = v[n1 + 8 * n2];
        dft_small<8>(t, kc);
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) a[n1][k2] = t[k2];
    }
    dft64_twiddle<1>(a, kc);
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++) {
        c64 t[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; n1++) t[n1] = a[n1][k2];
        dft_small<8>(t, kc);
#pragma unroll
        for (int k1 = 0; k1 < 8; k1++) v[8 * k1 + k2] = t[k1];
    }
}

// ---- butterflies with the twiddle folded into fused multiply-adds (the 64-point DFT of the large transforms, which are
// bound by the FP32 pipe):  x = a + W b costs two FFMA2 -- t = b_im (w_im, w_re) + (-a_re, a_im), x = b_re (w_re, w_im) +
// (-t_re, t_im) -- and y = a - W b = 2 a - x a third, against FMUL2 + FFMA2 + FADD2 + FADD2 for multiply-then-butterfly.
__device__ __forceinline__ c64 cneg(c64 a) { float x, y; unpack2(a, x, y); return pack2(-x, -y); }
// W a runtime twiddle (register or uniform register)
__device__ __forceinline__ void bfly_tw(c64 a, c64 b, c64 w, const FftConsts& kc, c64& x, c64& y) {
    float ar, ai, br, bi, wx, wy, tx, ty;
    unpack2(a, ar, ai);
    unpack2(b, br, bi);
    unpack2(w, wx, wy);
    unpack2(fma2(pack2(bi, bi), pack2(wy, wx), pack2(-ar, ai)), tx, ty);
    x = fma2(pack2(br, br), w, pack2(-tx, ty));
    y = fma2(a, kc.two, cneg(x));
}
// W = SGN * (SW ? swap(base) : base), base one of the constant pairs of FftConsts
template <int SW, int SGN>
__device__ __forceinline__ void bfly_base(c64 a, c64 b, c64 base, const FftConsts& kc, c64& x, c64& y) {
    float ar, ai, br, bi, bx, by, tx, ty;
    unpack2(a, ar, ai);
    unpack2(b, br, bi);
    unpack2(base, bx, by);
    const float u = SW ? by : bx, v = SW ? bx : by;              // W = SGN (u + i v)
    const float sbi = SGN > 0 ? bi : -bi, sbr = SGN > 0 ? br : -br;
    unpack2(fma2(pack2(sbi, sbi), pack2(v, u), pack2(-ar, ai)), tx, ty);
    x = fma2(pack2(sbr, sbr), pack2(u, v), pack2(-tx, ty));
    y = fma2(a, kc.two, cneg(x));
}
// (a + W64^K b, a - W64^K b), K a compile-time constant
template <int K>
__device__ __forceinline__ void bfly_w64(c64 a, c64 b, const FftConsts& kc, c64& x, c64& y) {
    constexpr int k = K & 63;
    if constexpr (k % 16 == 0) {
        constexpr int q = k / 16;
        if constexpr (q == 0) { x = cadd(a, b); y = csub(a, b); }
        else if constexpr (q == 1) { x = cadd_mi(a, b); y = cadd_pi(a, b); }
        else if constexpr (q == 2) { x = csub(a, b); y = cadd(a, b); }
        else { x = cadd_pi(a, b); y = cadd_mi(a, b); }
    } else if constexpr ((k & 1) == 0) {              // a power of W32, decoded as in cmul_w32
        constexpr int h = k / 2, q = h / 8, r = h % 8;
        if constexpr (r <= 4) {
            constexpr int i = r - 1;
            if constexpr (q == 0) bfly_base<0, 1>(a, b, kc.bm[i], kc, x, y);
            else if constexpr (q == 1) bfly_base<1, -1>(a, b, kc.bp[i], kc, x, y);
            else if constexpr (q == 2) bfly_base<0, -1>(a, b, kc.bm[i], kc, x, y);
            else bfly_base<1, 1>(a, b, kc.bp[i], kc, x, y);
        } else {
            constexpr int i = 8 - r - 1;
            if constexpr (q == 0) bfly_base<1, -1>(a, b, kc.bm[i], kc, x, y);
            else if constexpr (q == 1) bfly_base<0, -1>(a, b, kc.bp[i], kc, x, y);
            else if constexpr (q == 2) bfly_base<1, 1>(a, b, kc.bm[i], kc, x, y);
            else bfly_base<0, 1>(a, b, kc.bp[i], kc, x, y);
        }
    } else {                                          // an odd power of W64, decoded as in cmul_w64
        constexpr int q = k / 16, r = k % 16;
        if constexpr (r < 8) {
            constexpr int i = (r - 1) / 2;
            if constexpr (q == 0) bfly_base<0, 1>(a, b, kc.bm64[i], kc, x, y);
            else if constexpr (q == 1) bfly_base<1, -1>(a, b, kc.bp64[i], kc, x, y);
            else if constexpr (q == 2) bfly_base<0, -1>(a, b, kc.bm64[i], kc, x, y);
            else bfly_base<1, 1>(a, b, kc.bp64[i], kc, x, y);
        } else {
            constexpr int i = (16 - r - 1) / 2;
            if constexpr (q == 0) bfly_base<1, -1>(a, b, kc.bm64[i], kc, x, y);
            else if constexpr (q == 1) bfly_base<0, -1>(a, b, kc.bp64[i], kc, x, y);
            else if constexpr (q == 2) bfly_base<1, 1>(a, b, kc.bm64[i], kc, x, y);
            else bfly_base<0, 1>(a, b, kc.bp64[i], kc, x, y);
        }
    }
}

// first layer of the 8-point DFT on inputs that still carry a twiddle each: the pair (t_a w_a, t_b w_b) -> sum, difference.
// MODE 0: no twiddles.  MODE 1: runtime twiddles w[1..7] (w[0] is not read).  MODE 2: input i carries W64^(KSTEP i).
// MODE 3: runtime twiddles w[0..7], input 0 included.
template <int MODE, int KSTEP, int IA, int IB>
__device__ __forceinline__ void dft8_pair(const c64 (&t)[8], const c64* w, const FftConsts& kc, c64& x, c64& y) {
    if constexpr (MODE == 0) {
        x = cadd(t[IA], t[IB]);
        y = csub(t[IA], t[IB]);
    } else if constexpr (MODE == 1 || MODE == 3) {
        const c64 p = (IA == 0 && MODE == 1) ? t[IA] : cmul(t[IA], w[IA]);
        bfly_tw(p, t[IB], w[IB], kc, x, y);
    } else {
        const c64 p = (IA == 0) ? t[IA] : cmul_w64<KSTEP * IA>(t[IA], kc);
        bfly_w64<KSTEP * IB>(p, t[IB], kc, x, y);
    }
}
// forward 8-point DFT of (t_i w_i), natural order in and out: 26 packed instructions plain, 36 with seven twiddles
template <int MODE, int KSTEP>
__device__ __forceinline__ void dft8_tw(c64 (&t)[8], const c64* w, const FftConsts& kc) {
    c64 s04, d04, s26, d26, s15, d15, s37, d37;
    dft8_pair<MODE, KSTEP, 0, 4>(t, w, kc, s04, d04);
    dft8_pair<MODE, KSTEP, 2, 6>(t, w, kc, s26, d26);
    dft8_pair<MODE, KSTEP, 1, 5>(t, w, kc, s15, d15);
    dft8_pair<MODE, KSTEP, 3, 7>(t, w, kc, s37, d37);
    const c64 e0 = cadd(s04, s26), e2 = csub(s04, s26), e1 = cadd_mi(d04, d26), e3 = cadd_pi(d04, d26);
    const c64 o0 = cadd(s15, s37), o2 = csub(s15, s37), o1 = cadd_mi(d15, d37), o3 = cadd_pi(d15, d37);
    t[0] = cadd(e0, o0);
    t[4] = csub(e0, o0);
    bfly_w64<8>(e1, o1, kc, t[1], t[5]);            // W8
    t[2] = cadd_mi(e2, o2);
    t[6] = cadd_pi(e2, o2);
    bfly_w64<24>(e3, o3, kc, t[3], t[7]);           // W8^3
}

// second level of the 64-point DFT: column K2 of a[n1][k2] carries W64^(n1 K2) (TWMODE 0, 1) or the uniform factors
// wu[(n1 (2 K2 + 1)) & 127] (TWMODE 2: W64^(n1 K2) times the W128^(n1) left over from the split kernel's input twiddle)
template <int TWMODE, int K2>
__device__ __forceinline__ void dft64_cols(c64 (&a)[8][8], c64 (&v)[64], const FftConsts& kc, const float2* wu) {
    if constexpr (K2 < 8) {
        c64 t[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; n1++) t[n1] = a[n1][K2];
        if constexpr (TWMODE == 2) {
            c64 w[8];
#pragma unroll
            for (int n1 = 1; n1 < 8; n1++) w[n1] = *reinterpret_cast<const c64*>(&wu[8 + 8 * K2 + n1]);
            dft8_tw<1, 0>(t, w, kc);
        } else if constexpr (K2 == 0) {
            dft8_tw<0, 0>(t, nullptr, kc);
        } else {
            dft8_tw<2, K2>(t, nullptr, kc);
        }
#pragma unroll
        for (int k1 = 0; k1 < 8; k1++) v[8 * k1 + K2] = t[k1];
        dft64_cols<TWMODE, K2 + 1>(a, v, kc, wu);
    }
}
// forward 64-point DFT, 64 = 8 x 8 with n = n1 + 8 n2, k = 8 k1 + k2:  W64^(n k) = W8^(n2 k2) W64^(n1 k2) W8^(n1 k1).
// TWMODE 0: of v.  TWMODE 1: of v[m] z^m, given za[n2] = z^(8 n2) and zb[n1] = z^(n1) (n1, n2 = 1..7; index 0 unused):
// z^(8 n2) rides in the first level's butterflies, z^(n1) multiplies that level's outputs.  TWMODE 2: of v[m] W128^m, the
// odd half of a split burst: W128^(n1 + 8 n2) = W128^(n1) W64^(4 n2), the second factor in the first level's butterflies,
// the first merged into the second level's twiddles (wu = the 128 powers of W128, uniform).
template <int TWMODE>
__device__ __forceinline__ void dft64_f(c64 (&v)[64], const FftConsts& kc, const c64* za, const c64* zb, const float2* wu) {
    c64 a[8][8];
#pragma unroll
    for (int n1 = 0; n1 < 8; n1++) {
        c64 t[8];
#pragma unroll
        for (int n2 = 0; n2 < 8; n2++) t[n2] = v[n1 + 8 * n2];
        if constexpr (TWMODE == 1) dft8_tw<1, 0>(t, za, kc);
        else if constexpr (TWMODE == 2) {
            c64 w[8];
#pragma unroll
            for (int n2 = 1; n2 < 8; n2++) w[n2] = *reinterpret_cast<const c64*>(&wu[n2]);
            dft8_tw<1, 0>(t, w, kc);
        }
        else dft8_tw<0, 0>(t, nullptr, kc);
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) a[n1][k2] = (TWMODE == 1 && n1 > 0) ? cmul(t[k2], zb[n1]) : t[k2];
    }
    dft64_cols<TWMODE, 0>(a, v, kc, wu);
}

// 32 = 4 x 8 with n = n1 + 4 n2, k = 8 k1 + k2:  W32^(n k) = W8^(n2 k2) W32^(n1 k2) W4^(n1 k1); the same FMA-form butterflies.
// TW: input m carries the runtime twiddle w[m] (m = 1..31), folded into the first level.
template <int K2>
__device__ __forceinline__ void dft32_cols(c64 (&a)[4][8], c64 (&v)[32], const FftConsts& kc) {
    if constexpr (K2 < 8) {
        c64 s02, d02, s13, d13;
        bfly_w64<4 * K2>(a[0][K2], a[2][K2], kc, s02, d02);            // W32^(2 K2)
        const c64 p = cmul_w64<2 * K2>(a[1][K2], kc);                    // W32^(K2)
        bfly_w64<6 * K2>(p, a[3][K2], kc, s13, d13);                    // W32^(3 K2)
        v[K2] = cadd(s02, s13);
        v[16 + K2] = csub(s02, s13);
        v[8 + K2] = cadd_mi(d02, d13);
        v[24 + K2] = cadd_pi(d02, d13);
        dft32_cols<K2 + 1>(a, v, kc);
    }
}
template <bool TW>
__device__ __forceinline__ void dft32_f(c64 (&v)[32], const c64* w, const FftConsts& kc) {
    c64 a[4][8];
#pragma unroll
    for (int n1 = 0; n1 < 4; n1++) {
        c64 t[8], wt[8];
#pragma unroll
        for (int n2 = 0; n2 < 8; n2++) { t[n2] = v[n1 + 4 * n2]; wt[n2] = TW ? w[n1 + 4 * n2] : 0ull; }
        if constexpr (!TW) dft8_tw<0, 0>(t, nullptr, kc);
        else if (n1 == 0) dft8_tw<1, 0>(t, wt, kc);
        else dft8_tw<3, 0>(t, wt, kc);
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) a[n1][k2] = t[k2];
    }
    dft32_cols<0>(a, v, kc);
}
// 16 = 2 x 8 with n = n1 + 2 n2, k = 8 k1 + k2:  W16^(n k) = W8^(n2 k2) W16^(n1 k2) W2^(n1 k1)
template <int K2>
__device__ __forceinline__ void dft16_cols(c64 (&a)[2][8], c64 (&v)[16], const FftConsts& kc) {
    if constexpr (K2 < 8) {
        bfly_w64<4 * K2>(a[0][K2], a[1][K2], kc, v[K2], v[8 + K2]);     // W16^(K2)
        dft16_cols<K2 + 1>(a, v, kc);
    }
}
template <bool TW>
__device__ __forceinline__ void dft16_f(c64 (&v)[16], const c64* w, const FftConsts& kc) {
    c64 a[2][8];
#pragma unroll
    for (int n1 = 0; n1 < 2; n1++) {
        c64 t[8], wt[8];
#pragma unroll
        for (int n2 = 0; n2 < 8; n2++) { t[n2] = v[n1 + 2 * n2]; wt[n2] = TW ? w[n1 + 2 * n2] : 0ull; }
        if constexpr (!TW) dft8_tw<0, 0>(t, nullptr, kc);
        else if (n1 == 0) dft8_tw<1, 0>(t, wt, kc);
        else dft8_tw<3, 0>(t, wt, kc);
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) a[n1][k2] = t[k2];
    }
    dft16_cols<0>(a, v, kc);
}

template <int SKEW>
__host__ __device__ constexpr int fft_skew(int i) { return i + i / SKEW; }

// Offset of a compile-time displacement c from a per-thread base whose skewed address is already known:
// skew(base + c) = skew(base) + fft_off(c), provided the low parts never carry into another pad slot
// (proved per stage at compile time by fft_stage_linear_ok).
template <int SKEW>
__host__ __device__ constexpr int fft_off(int c) { return c + c / SKEW; }

// ---- bulk-copy engine + mbarrier (the burst's way into shared memory for the 64-points-per-thread kernels)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, int bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, int parity) {
    asm volatile("{\n .reg .pred p;\n LAB_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra LAB_DONE;\n bra LAB_WAIT;\n LAB_DONE:\n}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// what a transform needs to fetch its next burst
struct FftFeed {
    unsigned long long* bar;    // the transform's mbarrier
    c64* dst;                   // the transform's work buffer
    const c64* next_src;        // the next burst, nullptr when there is none
    int parity;                 // phase of the current burst
    int bytes;
};
// one thread of the transform
__device__ __forceinline__ void fft_feed_issue(const FftFeed& fd, const c64* src) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the buffer's earlier generic-proxy accesses come first
    mbar_arrive_expect_tx(fd.bar, fd.bytes);
    bulk_copy_g2s(fd.dst, src, fd.bytes, fd.bar);
}

// the threads of one transform exchange data between passes: a warp-level barrier is enough when a transform lives
// inside one warp (n / points-per-thread <= 32)
template <int TPF, int THREADS>
__device__ __forceinline__ void fft_sync(int fl) {
    if (TPF <= 32) __syncwarp();
    else if (TPF < THREADS) asm volatile("bar.sync %0, %1;" ::"r"(fl + 1), "n"(TPF) : "memory");      // the transform's own warps only
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
    float2 kbase[16];      // (cos, sin) and (cos, -sin) of 2 pi r / 32, r = 1..4, then of 2 pi r / 64, r = 1, 3, 5, 7: see FftConsts
    float2 two;            // (2, 2)
    float2 wsplit[2][72];  // the split 8192-point kernel's stage-0 factors per half, see fft_wsplit_host
};

// One Stockham stage for the P points a thread owns: radix R, sub-transform length NS already done.
// GEN: the general entry points (inverse through conjugation); the estimator instantiation loads plain.
// SPLIT: this transform is one half (fl = 0 even bins, 1 odd bins) of a burst of 2 N points, see qpsk_fft_split().
template <int LOG2N, int R, int NS, bool FIRST, bool LAST, bool GEN, bool SPLIT>
__device__ __forceinline__ void fft_stage(c64 (&pts)[FftCfg<LOG2N>::P], c64* sdat, const c64* stw, const c64* gin, int j, int base, int fl, float imsgn,
                                          const FftConsts& kc, const FftFeed& fd, const float2* w128) {
    using Cfg = FftCfg<LOG2N>;
    constexpr int N = Cfg::N, P = Cfg::P, TPF = Cfg::TPF, NB = P / R, S = Cfg::SKEW;   // NB butterflies per thread
    constexpr int STRIDE = N / R;
    constexpr int KT = qpsk_fft_tw_cols(NS, TPF, P);     // twiddle columns of this stage
    constexpr bool FACT = LAST && NS > TPF && P <= 32;   // w^(m k), k = j + t TPF  =  table[m][j] * W_P^(m t)
    constexpr bool FACT64 = qpsk_fft_tw_fact64(NS, R, TPF, P);      // the same for a radix-2 remainder stage of the 64-point kernels
    constexpr bool TW2L = qpsk_fft_tw_two_level(NS, R, P);          // z^(8a + b) = z^(8a) z^b from 14 entries per thread
    constexpr bool PERBF = NS > TPF && !FACT && !FACT64; // whole rows in the table: butterfly t reads column j + t TPF
    constexpr bool PRETW = qpsk_fft_tw_on_store(NS, P);  // this stage's inputs were twiddled by their producer
    constexpr bool TWOUT = FIRST && !LAST && qpsk_fft_tw_on_store(R, P);   // ... and this is that producer
    constexpr int RNEXT = LAST ? 1 : qpsk_fft_radix(N / (NS * R), P);
    static_assert(!TWOUT || (NB == 1 && NS == 1 && (N / RNEXT) % R == 0), "twiddle-on-store assumes one first-stage butterfly per thread");
    constexpr bool PRELOAD = NS > 1 && !PRETW && !PERBF && !TW2L && NB > 1; // the same R - 1 twiddles serve every butterfly of the thread
    constexpr bool SMEM_IN = !FIRST || Cfg::TMA_IN;      // inputs come from the work buffer
    constexpr bool RELEASE = LAST && Cfg::TMA_IN;        // the last reader of the buffer hands it to the next burst's copy
    static_assert(!SPLIT || (Cfg::TMA_IN && R == 64 && NB == 1), "the split kernel is built from the bulk-fed 64 x 64 transform");
    static_assert(LAST || NS <= TPF, "only a remainder stage may have more sub-transform columns than threads");
    static_assert(!FACT || (N / TPF == P && 32 % P == 0), "factored twiddles assume n / tpf == P, a divisor of 32");

    const c64* rd = sdat + fft_skew<S>(base + j);
    constexpr int KTE = FACT64 ? TPF : KT;
    const c64* twp = stw + ((PERBF || FACT64) ? j : j % KT);
    c64 tw[PRELOAD ? R - 1 : 1];
    c64 tw2[TW2L ? 14 : 1];
    if (TW2L) {
#pragma unroll
        for (int i = 0; i < 14; i++) tw2[i] = Cfg::TW_SMEM ? stw[i * 64 + j % 64] : __ldg(stw + i * 64 + j % 64);
    }
    if (PRELOAD) {
#pragma unroll
        for (int m = 1; m < R; m++) tw[m - 1] = Cfg::TW_SMEM ? twp[(m - 1) * KTE] : __ldg(twp + (m - 1) * KTE);
    }
    if (FIRST && Cfg::TMA_IN) mbar_wait(fd.bar, fd.parity);       // the burst has landed
    if (RELEASE) {
        // all of the thread's inputs first, so that the buffer can be given away before the arithmetic starts
#pragma unroll
        for (int t = 0; t < NB; t++)
#pragma unroll
            for (int r = 0; r < R; r++) pts[t * R + r] = rd[fft_off<S>(t * TPF + r * STRIDE)];
#pragma unroll
        for (int i = 0; i < P; i++) asm volatile("" ::"l"(pts[i]));      // the loads have completed ...
        fft_sync<TPF, Cfg::THREADS>(fl);                                 // ... in every thread of the transform
        if (j == 0 && fd.next_src != nullptr) fft_feed_issue(fd, fd.next_src);
    }
#pragma unroll
    for (int t = 0; t < NB; t++) {
        c64 v[R];
        if (FIRST && SPLIT) {
            // the radix-2 step of the decimation in frequency, as the burst is read: x[n] +- x[n + N], n = j + r STRIDE
            // x[n] + x[n + N] (even-bin half) or x[n] - x[n + N] (odd-bin half) through one instruction stream: x0 + s x1
            const float sg = fl ? -1.0f : 1.0f;
            const c64 sgn = pack2(sg, sg);
#pragma unroll
            for (int r = 0; r < R; r++) {
                v[r] = fma2(fd.dst[j + r * STRIDE + N], sgn, fd.dst[j + r * STRIDE]);
                if (GEN) v[r] = cconj_if(v[r], imsgn);               // the odd half's factor w^n = w^j W128^r: W128^r inside the DFT below, w^j in the next stage's twiddles
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int c = t * TPF + r * STRIDE;          // butterfly j + t TPF, input r
            if (FIRST && SPLIT) {
            } else if (RELEASE) {
                v[r] = pts[t * R + r];
            } else if (!SMEM_IN) {
                v[r] = gin[j + c];
            } else if (FIRST) {
                v[r] = fd.dst[j + c];                     // the burst as the bulk copy left it: unpadded
            } else if (Cfg::LIN) {
                v[r] = rd[fft_off<S>(c)];
            } else {
                v[r] = sdat[fft_skew<S>(base + j + c)];
            }
            if (FIRST && GEN && !SPLIT) v[r] = cconj_if(v[r], imsgn);    // inverse = conj(FFT(conj x))
        }
        // radix 32 with plain per-input twiddles: the FMA-form DFT folds them into its first level
        // (radix 16 measured 2 % slower in this form at n = 256 and stays on dft_small<16>)
        constexpr bool FMAF = QPSK_FFT_FMA_SMALL && R == 32 && !FACT && !FACT64;
        c64 wf[FMAF ? R : 1];
        if (FMAF && NS > 1 && !PRETW) {
#pragma unroll
            for (int m = 1; m < R; m++) {
                if (PRELOAD) wf[m] = tw[m - 1];
                else wf[m] = Cfg::TW_SMEM ? twp[(m - 1) * KTE + (PERBF ? t * TPF : 0)] : __ldg(twp + (m - 1) * KTE + (PERBF ? t * TPF : 0));
            }
        }
        if (TW2L || FMAF) {
            // applied inside the DFT below
        } else if (NS > 1 && !PRETW) {
#pragma unroll
            for (int m = 1; m < R; m++) {
                c64 w;
                if (PRELOAD) w = tw[m - 1];
                else w = Cfg::TW_SMEM ? twp[(m - 1) * KTE + (PERBF ? t * TPF : 0)] : __ldg(twp + (m - 1) * KTE + (PERBF ? t * TPF : 0));
                v[m] = cmul(v[m], w);
                if (FACT64 && t > 0) {
                    switch (t & 63) {
#define QPSK_W64_CASE(K) case K: v[m] = cmul_w64<K>(v[m], kc); break;
                        QPSK_W64_CASE(1) QPSK_W64_CASE(2) QPSK_W64_CASE(3) QPSK_W64_CASE(4) QPSK_W64_CASE(5) QPSK_W64_CASE(6) QPSK_W64_CASE(7)
                        QPSK_W64_CASE(8) QPSK_W64_CASE(9) QPSK_W64_CASE(10) QPSK_W64_CASE(11) QPSK_W64_CASE(12) QPSK_W64_CASE(13) QPSK_W64_CASE(14)
                        QPSK_W64_CASE(15) QPSK_W64_CASE(16) QPSK_W64_CASE(17) QPSK_W64_CASE(18) QPSK_W64_CASE(19) QPSK_W64_CASE(20) QPSK_W64_CASE(21)
                        QPSK_W64_CASE(22) QPSK_W64_CASE(23) QPSK_W64_CASE(24) QPSK_W64_CASE(25) QPSK_W64_CASE(26) QPSK_W64_CASE(27) QPSK_W64_CASE(28)
                        QPSK_W64_CASE(29) QPSK_W64_CASE(30) QPSK_W64_CASE(31)
#undef QPSK_W64_CASE
                        default: break;
                    }
                }
                if (FACT && t > 0) {
                    // W_P^(m t) as a power of W32
                    switch ((m * t * (32 / (P <= 32 ? P : 32))) & 31) {
#define QPSK_W32_CASE(K) case K: v[m] = cmul_w32<K>(v[m], kc); break;
                        QPSK_W32_CASE(1) QPSK_W32_CASE(2) QPSK_W32_CASE(3) QPSK_W32_CASE(4) QPSK_W32_CASE(5) QPSK_W32_CASE(6) QPSK_W32_CASE(7)
                        QPSK_W32_CASE(8) QPSK_W32_CASE(9) QPSK_W32_CASE(10) QPSK_W32_CASE(11) QPSK_W32_CASE(12) QPSK_W32_CASE(13) QPSK_W32_CASE(14)
                        QPSK_W32_CASE(15) QPSK_W32_CASE(16) QPSK_W32_CASE(17) QPSK_W32_CASE(18) QPSK_W32_CASE(19) QPSK_W32_CASE(20) QPSK_W32_CASE(21)
                        QPSK_W32_CASE(22) QPSK_W32_CASE(23) QPSK_W32_CASE(24) QPSK_W32_CASE(25) QPSK_W32_CASE(26) QPSK_W32_CASE(27) QPSK_W32_CASE(28)
                        QPSK_W32_CASE(29) QPSK_W32_CASE(30) QPSK_W32_CASE(31)
#undef QPSK_W32_CASE
                        default: break;
                    }
                }
            }
        }
        if constexpr (R == 64) {
            if constexpr (TW2L) {
                c64 za[8], zb[8];
#pragma unroll
                for (int i = 1; i < 8; i++) { za[i] = tw2[i - 1]; zb[i] = tw2[7 + i - 1]; }
                dft64_f<1>(v, kc, za, zb, nullptr);
            } else if constexpr (FIRST && SPLIT) {
                dft64_f<2>(v, kc, nullptr, nullptr, w128 + fl * 72);      // w128: the half's row of FftArgs::wsplit
            } else {
                dft64_f<0>(v, kc, nullptr, nullptr, nullptr);
            }
        } else if constexpr (FMAF && R == 32) {
            if constexpr (NS > 1 && !PRETW) dft32_f<true>(v, wf, kc);
            else dft32_f<false>(v, nullptr, kc);
        } else if constexpr (FMAF && R == 16) {
            if constexpr (NS > 1 && !PRETW) dft16_f<true>(v, wf, kc);
            else dft16_f<false>(v, nullptr, kc);
        } else {
            dft_small<R>(v, kc);
        }
        if (TWOUT) {
            // output q of butterfly j is input r' = j R RNEXT / N of the next stage's butterfly q: times w^(q r'), table [q - 1][r']
            const c64* two = stw + (j * (R * RNEXT)) / N;
#pragma unroll
            for (int q = 1; q < R; q++) v[q] = cmul(v[q], Cfg::TW_SMEM ? two[(q - 1) * RNEXT] : __ldg(two + (q - 1) * RNEXT));
        }
#pragma unroll
        for (int q = 0; q < R; q++) pts[t * R + q] = v[q];
    }
    if (!LAST) {
        if (SMEM_IN) fft_sync<TPF, Cfg::THREADS>(fl);  // everyone has read this stage's inputs
        // butterfly jj = j + t TPF writes o + q NS, o = (jj / NS) NS R + jj % NS
        const int dyn = (NS == 1) ? j * R : (j / NS) * NS * R + (j % NS);
        c64* wr = sdat + fft_skew<S>(base + dyn);
#pragma unroll
        for (int t = 0; t < NB; t++) {
#pragma unroll
            for (int q = 0; q < R; q++) {
                const int c = t * TPF * R + q * NS;
                if (Cfg::LIN) wr[fft_off<S>(c)] = pts[t * R + q];
                else sdat[fft_skew<S>(base + dyn + c)] = pts[t * R + q];
            }
        }
        fft_sync<TPF, Cfg::THREADS>(fl);
    }
}

// compile-time proof of the linear-offset claims for one stage (evaluated by static_assert in fft_stages)
template <int LOG2N, int R, int NS, bool FIRST, bool LAST>
__host__ __device__ constexpr bool fft_stage_linear_ok() {
    using Cfg = FftCfg<LOG2N>;
    constexpr int N = Cfg::N, P = Cfg::P, TPF = Cfg::TPF, NB = P / R, S = Cfg::SKEW;
    if (!Cfg::LIN) return true;
    if ((Cfg::N % S) != 0) return false;                               // transform bases are multiples of S
    if (!FIRST) {                                                      // reads: base + j + c (a first stage reads unpadded data)
        for (int t = 0; t < NB; t++)
            for (int r = 0; r < R; r++) {
                const int c = t * TPF + r * (N / R);
                if (TPF >= S) { if (c % S != 0) return false; }        // j spans whole pad periods: c must not disturb them
                else if ((c % S) + TPF > S) return false;              // j < TPF < S: no carry into the next pad slot
            }
    }
    if (!LAST) {                                                       // writes: base + dyn + c
        if (NS == 1) { if (R != S) return false; }                     // dyn = j R is a multiple of S, c = t TPF R + q, q < S
        else if (NS % S != 0 || TPF % NS != 0) return false;           // dyn = (j/NS) NS R + j % NS, c a multiple of S
        for (int t = 0; t < NB; t++)
            for (int q = 0; q < R; q++) {
                const int c = t * TPF * R + q * NS;
                if (NS == 1) { if ((c - q) % S != 0 || q >= S) return false; }
                else if (c % S != 0) return false;
            }
    }
    return true;
}

template <int LOG2N, int NS, bool FIRST, bool GEN, bool SPLIT>
__device__ __forceinline__ void fft_stages(c64 (&pts)[FftCfg<LOG2N>::P], c64* sdat, const c64* stw, const c64* gin, int j, int base, int fl, float imsgn,
                                           const FftConsts& kc, const FftFeed& fd, const float2* w128) {
    constexpr int N = FftCfg<LOG2N>::N;
    constexpr int REM = N / NS;
    constexpr int RMAX = FftCfg<LOG2N>::P;
    constexpr int R = qpsk_fft_radix(REM, RMAX);
    constexpr bool LAST = (NS * R == N);
    static_assert(fft_stage_linear_ok<LOG2N, R, NS, FIRST, LAST>(), "skewed shared-memory offsets of this stage are not linear");
    fft_stage<LOG2N, R, NS, FIRST, LAST, GEN, SPLIT>(pts, sdat, stw, gin, j, base, fl, imsgn, kc, fd, w128);
    if constexpr (!LAST)
        fft_stages<LOG2N, NS * R, false, GEN, SPLIT>(pts, sdat, stw + qpsk_fft_tw_entries(NS, R, FftCfg<LOG2N>::TPF, FftCfg<LOG2N>::P), gin, j, base, fl, imsgn, kc, fd, w128);
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

// GEN = false: forward transform, argmax only (a.bin required; a.spectrum, a.im_sign ignored) -- the estimator hot path.
// GEN = true : forward or inverse, optional spectrum store, optional argmax.
template <int LOG2N, bool GEN>
__global__ void __launch_bounds__(FftCfg<LOG2N>::THREADS, FftCfg<LOG2N>::MINB) fft_kernel(const __grid_constant__ FftArgs a) {
    using Cfg = FftCfg<LOG2N>;
    constexpr int N = Cfg::N, P = Cfg::P, TPF = Cfg::TPF, FPB = Cfg::FPB;
    constexpr bool SPLIT = qpsk_fft_split(N);           // the burst is two half-length transforms side by side
    constexpr int SL = SPLIT ? LOG2N - 1 : LOG2N;       // log2 of the length the stages work on
    using Sub = FftCfg<SL>;
    static_assert(!SPLIT || (Sub::P == P && Sub::TPF * 2 == TPF && FPB == 1 && !Cfg::TW_SMEM), "split kernel layout");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c64* sdat = reinterpret_cast<c64*>(smem_raw);
    c64* stw_s = sdat + Cfg::SKEW_PTS;
    float* red_mag = reinterpret_cast<float*>(stw_s + (Cfg::TW_SMEM ? Cfg::TW : 0));
    int* red_idx = reinterpret_cast<int*>(red_mag + 64);
    const c64* stw = Cfg::TW_SMEM ? stw_s : reinterpret_cast<const c64*>(a.tw);

    if (Cfg::TW_SMEM)
        for (int i = threadIdx.x; i < Cfg::TW; i += blockDim.x) stw_s[i] = reinterpret_cast<const c64*>(a.tw)[i];
    __syncthreads();

    const int fl = threadIdx.x / TPF, j = threadIdx.x % TPF;
    const int base = fl * N;
    const int half = SPLIT ? j / Sub::TPF : 0, js = SPLIT ? j % Sub::TPF : j;      // which half, thread within it
    const float scale2 = a.scale * a.scale;
    FftConsts kc;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        kc.bp[r] = *reinterpret_cast<const c64*>(&a.kbase[r]);
        kc.bm[r] = *reinterpret_cast<const c64*>(&a.kbase[4 + r]);
        kc.bp64[r] = *reinterpret_cast<const c64*>(&a.kbase[8 + r]);
        kc.bm64[r] = *reinterpret_cast<const c64*>(&a.kbase[12 + r]);
    }
    kc.two = *reinterpret_cast<const c64*>(&a.two);
    constexpr int WPF = (TPF > 32) ? TPF / 32 : 1;      // warps per transform
    const int wib = threadIdx.x >> 5;
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(red_idx + 64) + fl;      // one per transform of the CTA
    FftFeed fd;
    fd.bar = mbar; fd.dst = sdat + fft_skew<Cfg::SKEW>(base); fd.next_src = nullptr; fd.parity = 0; fd.bytes = N * 8;
    if (Cfg::TMA_IN) {
        if (j == 0) mbar_init(mbar, 1);                  // one thread of the transform arrives per burst, with the byte count
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
        const int bfirst = blockIdx.x * FPB + fl;
        if (j == 0 && blockIdx.x * FPB < a.nbursts)
            fft_feed_issue(fd, reinterpret_cast<const c64*>(a.in) + (size_t)(bfirst < a.nbursts ? bfirst : a.nbursts - 1) * N);
    }
    int pass = 0;
    for (int b0 = blockIdx.x * FPB; b0 < a.nbursts; b0 += gridDim.x * FPB) {
        const int b = b0 + fl;
        const bool active = b < a.nbursts;
        // an inactive slot of the last pass recomputes the last burst; only its stores are masked
        const c64* gin = reinterpret_cast<const c64*>(a.in) + (size_t)(active ? b : a.nbursts - 1) * N;
        if (Cfg::TMA_IN) {
            const long long bn0 = (long long)b0 + (long long)gridDim.x * FPB;
            const long long bn = bn0 + fl;
            fd.next_src = (bn0 < a.nbursts && half == 0) ? reinterpret_cast<const c64*>(a.in) + (size_t)(bn < a.nbursts ? bn : a.nbursts - 1) * N : nullptr;
            fd.parity = pass & 1;
        }
        // The burst this slot transforms QPSK_FFT_L2_AHEAD passes from now starts its way from HBM to L2 here: two
        // instructions per thread and no registers, and the loads of that pass then wait for L2 instead of DRAM
        // (waiting for stage 0's loads was the largest single stall: profiles/r02_fft4096_v1.summary.csv).
        if (QPSK_FFT_L2_AHEAD > 0 && !Cfg::TMA_IN) {
            const long long bn = (long long)b + (long long)QPSK_FFT_L2_AHEAD * gridDim.x * FPB;
            if (bn < a.nbursts) {
                const char* nx = reinterpret_cast<const char*>(a.in + (size_t)bn * N);
#pragma unroll
                for (int l = 0; l < (N * 8 / 128 + TPF - 1) / TPF; l++) {
                    const int line = j + l * TPF;
                    if (N * 8 / 128 >= TPF || line < N * 8 / 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)line * 128));
                }
            }
        }
        c64 pts[P];
        fft_stages<SL, 1, true, GEN, SPLIT>(pts, sdat, stw + (SPLIT ? half * (14 * 64) : 0), gin, js, base + half * Sub::N, SPLIT ? half : fl, a.im_sign, kc, fd, &a.wsplit[0][0]);
        // bin of pts[i]: a half of a split burst holds every other bin
        auto out_index = [&](int i) { return SPLIT ? 2 * fft_out_index<SL>(js, i) + half : fft_out_index<SL>(js, i); };

        if constexpr (GEN) {
            // ---- general epilogue: scale, optional spectrum store, |X|^2 argmax
            float best = -1.0f;
            int besti = 0x7fffffff;
#pragma unroll
            for (int i = 0; i < P; i++) {
                const int idx = out_index(i);
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
                    if ((threadIdx.x & 31) == 0) { red_mag[wib] = best; red_idx[32 + wib] = besti; }
                    __syncthreads();
                    if (j == 0 && active) {
                        const int w0 = fl * WPF;
                        for (int w = 1; w < WPF; w++) {
                            const float om = red_mag[w0 + w];
                            const int oi = red_idx[32 + w0 + w];
                            if (om > best || (om == best && oi < besti)) { best = om; besti = oi; }
                        }
                        a.bin[b] = besti;
                        if (a.mag2) a.mag2[b] = best;
                    }
                    __syncthreads();
                }
            }
        } else {
            // ---- estimator epilogue: |X|^2 of the unscaled outputs (the 1/n is a power of two: it commutes with every
            // rounding and is applied once to the winner), maximum first, bin second.  The maxima of groups of eight points
            // stay in registers; only the thread(s) holding the transform's maximum go back for the bin, and only into
            // the group(s) that hold it, recomputing the (bitwise identical) magnitudes there.
            constexpr int NG = (P >= 8) ? P / 8 : 1, GS = P / NG;
            float gm[NG];
#pragma unroll
            for (int gI = 0; gI < NG; gI++) {
                float mx = -1.0f;
#pragma unroll
                for (int i = 0; i < GS; i++) {
                    float x, y;
                    unpack2(mul2(pts[gI * GS + i], pts[gI * GS + i]), x, y);
                    mx = fmaxf(mx, __fadd_rn(x, y));
                }
                gm[gI] = mx;
            }
            float lmax = gm[0];
#pragma unroll
            for (int gI = 1; gI < NG; gI++) lmax = fmaxf(lmax, gm[gI]);
            // per warp (or per transform when it is narrower than a warp): the maximum, then the lowest bin that holds it
            constexpr int W = (TPF < 32) ? TPF : 32;
            float g = lmax;
            if (W == 32) {
                // |X|^2 >= 0: the bit patterns order like the values, so the warp maximum is one integer REDUX
                g = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(lmax)));
            } else {
#pragma unroll
                for (int off = W / 2; off > 0; off >>= 1) g = fmaxf(g, __shfl_xor_sync(0xffffffffu, g, off, W));
            }
            int cand = 0x7fffffff;
            if (lmax == g) {
#pragma unroll
                for (int gI = 0; gI < NG; gI++) {
                    if (gm[gI] == g) {
#pragma unroll
                        for (int i = 0; i < GS; i++) {
                            float x, y;
                            unpack2(mul2(pts[gI * GS + i], pts[gI * GS + i]), x, y);
                            if (__fadd_rn(x, y) == g) cand = min(cand, out_index(gI * GS + i));
                        }
                    }
                }
            }
            if (W == 32) {
                cand = __reduce_min_sync(0xffffffffu, cand);
            } else {
#pragma unroll
                for (int off = W / 2; off > 0; off >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, off, W));
            }
            if (TPF <= 32) {
                if (j == 0 && active) { a.bin[b] = cand; if (a.mag2) a.mag2[b] = g * scale2; }
            } else {
                // one shared-memory hop across the warps of the transform; the slots alternate between passes, so the only
                // barrier is the one that also closes the pass
                float* rm = red_mag + 32 * (pass & 1);
                int* ri = red_idx + 32 * (pass & 1);
                if ((threadIdx.x & 31) == 0) { rm[wib] = g; ri[wib] = cand; }
                fft_sync<TPF, Cfg::THREADS>(fl);
                if (j == 0 && active) {
                    const int w0 = fl * WPF;
                    float bm = rm[w0];
                    int bi = ri[w0];
#pragma unroll
                    for (int w = 1; w < WPF; w++) {
                        const float om = rm[w0 + w];
                        const int oi = ri[w0 + w];
                        if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; }
                    }
                    a.bin[b] = bi;
                    if (a.mag2) a.mag2[b] = bm * scale2;
                }
            }
        }
        // The next pass's stage-0 writes must not race with this pass's last-stage reads of shared memory.  With more than
        // one warp per transform the estimator epilogue's barrier came after those reads; otherwise close the pass here.
        if (GEN || TPF <= 32) fft_sync<TPF, Cfg::THREADS>(fl);
        pass++;
    }
}
