// probe.cuh -- measured FP32-pipe ceiling of the device the library runs on (bench.py's roofline denominator).
//
// The matched filter (rrc_fir.c:22-26) costs one rounded multiply and one rounded add per tap and component.  This
// kernel issues nothing but that pair, in the two SASS formulations the filters use -- FMUL2.FTZ + FADD2 (exact mode,
// the reference's unfused arithmetic) and FFMA2 (fast mode) -- in the shape of the filters' steady loop: sixteen
// accumulators per thread, one sample fetched from shared memory per sixteen tap updates, taps from the constant bank
// (uniform-datapath loads).  Eight warps per scheduler.  What it reports is the rate the scheduler + pipe sustain for this
// instruction mix: complex tap-updates per second, machine-wide.  A kernel of its own, not a call into the filter: the
// filter's fill phases, barriers and auxiliary warps are what the roofline fraction is meant to expose.
//
// Round 2: the round-1 probe (eight accumulators, a 15-sample window re-read every 128 packed instructions) reported
// 29.7 of the nominal 32 tap-updates per clock and SM.  tools/packed_peak_bench.cu showed why: a packed FP32x2
// instruction holds the scheduler's issue port for two cycles and EVERY other instruction except the uniform constant
// loads takes one more -- 256 + 15 loads + 5 loop instructions = 276 cycles per pass, exactly what it measured.  With the
// filters' real ratio (one load per 32 packed) the same pipe delivers 2.00-2.03 cycles per packed instruction, i.e. the
// nominal rate, and that is what the probe reports now.
#pragma once

#include "common.cuh"

#define QPSK_PROBE_R 16       // independent accumulators per thread
#define QPSK_PROBE_TAPS 16    // samples per trip; each meets QPSK_PROBE_R taps

struct ProbeTaps { float2 t[128 + 2 * QPSK_PROBE_R]; };

template <int FUSED>
__global__ void __launch_bounds__(256) fp32_pipe_probe_kernel(const float2* __restrict__ xin, const __grid_constant__ ProbeTaps tb, float2* __restrict__ out, int iters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ u64 xs[32][161];                                   // odd row stride: conflict-free 64-bit reads, lane = row as in the filters
    for (int i = threadIdx.x; i < 32 * 160; i += blockDim.x) { const float2 v = xin[i & 1023]; xs[i / 160][i % 160] = pack2(v.x, v.y); }
    __syncthreads();
    u64 ap[QPSK_PROBE_R];
#pragma unroll
    for (int r = 0; r < QPSK_PROBE_R; r++) ap[r] = 0ull;
    const u64* xrow = &xs[threadIdx.x & 31][0];
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        const int d0 = (it & 7) * QPSK_PROBE_TAPS;                // a fresh window and fresh taps every trip: nothing is loop-invariant
#pragma unroll
        for (int e = 0; e < QPSK_PROBE_TAPS; e++) {
            const u64 xv = xrow[d0 + e];
#pragma unroll
            for (int r = 0; r < QPSK_PROBE_R; r++) {
                const u64 cc = *reinterpret_cast<const u64*>(&tb.t[d0 + e + QPSK_PROBE_R - r]);
                ap[r] = FUSED ? fma2(xv, cc, ap[r]) : add2(ap[r], mul2_exact(xv, cc));
            }
        }
    }
#pragma unroll
    for (int r = 0; r < QPSK_PROBE_R; r++) {
        float2 v;
        unpack2(ap[r], v.x, v.y);
        out[(size_t)t * QPSK_PROBE_R + r] = v;
    }
}
