// common.cuh -- shared device helpers for the sm_100a QPSK receiver kernels.
//
// Arithmetic contract ("exact mode"): every float operation of the reference's Makefile build
// (-std=c11 => -ffp-contract=off, SSE scalar; /root/reference Makefile:7) is one IEEE-754
// binary32 operation.  Device code therefore uses explicit _rn intrinsics / PTX and the TU is
// compiled with -fmad=false.  See SURVEY.md Appendix A for the per-expression semantics.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;

// ---- packed FP32x2 (Blackwell FMUL2 / FADD2 / FFMA2): one complex sample {re,im} per 64-bit register
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// Rounded product.  The .ftz is deliberate: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2
// into FFMA2 at -O1 and above even with --fmad=false (scalar mul.rn/add.rn are never fused);
// a flush-mode mismatch between the two instructions is what blocks the contraction.  .ftz only
// changes results when an input or the product is subnormal (< 2^-126); receiver samples are
// >= 2^-14 * |phasor component| and taps >= ~1e-8, so that never happens on this path.
__device__ __forceinline__ u64 mul2_exact(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// rrc_fir.c:28  sample[j] = y * GAIN with GAIN a double literal: (float)((double)y * 1.85)
__device__ __forceinline__ float gain_exact(float y) { return __double2float_rn(__dmul_rn((double)y, 1.85)); }

// complex float multiply as GCC evaluates it on finite data:
// re = ar*br - ai*bi, im = ar*bi + ai*br, four rounded products and two rounded sums
__device__ __forceinline__ float2 cmul_exact(float2 a, float2 b) {
    const float ac = __fmul_rn(a.x, b.x), bd = __fmul_rn(a.y, b.y);
    const float ad = __fmul_rn(a.x, b.y), bc = __fmul_rn(a.y, b.x);
    return make_float2(__fsub_rn(ac, bd), __fadd_rn(ad, bc));
}

// the same four rounded products and two rounded sums as two packed multiplies and one packed add (3 FP32-pipe slots
// instead of 6).  The sign flip between them is an integer XOR, and the product that feeds the add directly carries
// .ftz, so ptxas cannot contract anything into an FFMA2 (see mul2_exact).
__device__ __forceinline__ float2 cmul_exact_packed(float2 a, float2 b) {
    const u64 p1 = mul2_exact(pack2(a.x, a.x), pack2(b.x, b.y));             // (ac, ad)
    const u64 p2 = mul2(pack2(a.y, a.y), pack2(b.y, b.x)) ^ 0x80000000ull;   // (-bd, bc)
    float2 r;
    unpack2(add2(p1, p2), r.x, r.y);
    return r;
}

// ---- glibc 2.39 sinf/cosf (x86-64 FMA ifunc variant) restated in FP64: reduce by pi/2, then
// one sine-type and one cosine-type polynomial with every a+b*c fused (sincosf.h: reduce_fast,
// sinf_poly).  Valid for |y| < 120.  Bit-identical to the host libm the reference links against
// (checked exhaustively on [-7,7] for the CPU restatement oracle/qpsk_oracle.c:orc_glibc_*).
__device__ __forceinline__ void sincosf_glibc(float y, float& s_out, float& c_out) {
    const double HPI_INV = 0x1.45F306DC9C883p+23, HPI = 0x1.921FB54442D18p0;
    const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5,
                 C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    const double x0 = (double)y;
    const double r = __dmul_rn(x0, HPI_INV);
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
    const double x = __fma_rn(-(double)n, HPI, x0);
    const double x2 = __dmul_rn(x, x);
    // sine-type polynomial on x * sign[n & 3], sign = {+,-,-,+}
    const int q = n & 3;
    const double xs = (q == 1 || q == 2) ? -x : x;
    const double x3 = __dmul_rn(xs, x2);
    const double s1 = __fma_rn(x2, S3, S2);
    const double x7 = __dmul_rn(x3, x2);
    const double sp = __fma_rn(x3, S1, xs);
    const float sv = __double2float_rn(__fma_rn(x7, s1, sp));
    // cosine-type polynomial; table row 1 (n & 2) has every C coefficient negated
    const double x4 = __dmul_rn(x2, x2);
    const double c2 = __fma_rn(x2, C4, C3);
    const double c1 = __fma_rn(x2, C1, C0);
    const double x6 = __dmul_rn(x4, x2);
    const double cp = __fma_rn(x4, C2, c1);
    float cv = __double2float_rn(__fma_rn(x6, c2, cp));
    if (n & 2) cv = -cv;
    // sinf: even n -> sine poly, odd n -> cosine poly; cosf the other way round
    s_out = (n & 1) ? cv : sv;
    c_out = (n & 1) ? sv : cv;
    // |y| < 2^-12 (abstop12 test of s_sinf.c / s_cosf.c): sinf returns y itself (keeps -0), cosf returns 1
    if (((__float_as_uint(y) >> 20) & 0x7ffu) < 0x398u) { s_out = y; c_out = 1.0f; }
}

#define QPSK_CHUNK 128          // samples per time tile of the front-end kernel
#define QPSK_GROUP 32           // channels per CTA (one per lane)
#define QPSK_MAX_TAPS 512
