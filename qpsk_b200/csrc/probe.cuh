// probe.cuh -- measured FP32-pipe ceiling of the device the library runs on (bench.py's roofline denominator).
//
// The matched filter (rrc_fir.c:22-26) costs one rounded multiply and one rounded add per tap and component.  This
// kernel issues nothing but that pair, in the two SASS formulations the filters use -- FMUL2.FTZ + FADD2 (exact mode,
// the reference's unfused arithmetic) and FFMA2 (fast mode) -- on eight independent accumulators per thread fed from a
// sample window in shared memory, so what it reports is the rate the pipe itself sustains: complex tap-updates per second,
// machine-wide.  A kernel of its own, not a call into the filter: the filter's loads, barriers and auxiliary warps are what
// the roofline fraction is meant to expose.
#pragma once

#include "common.cuh"

#define QPSK_PROBE_R 8        // independent accumulators per thread
#define QPSK_PROBE_TAPS 8     // taps per inner pass (in registers)
#define QPSK_PROBE_WIN (QPSK_PROBE_R + QPSK_PROBE_TAPS - 1)

template <int FUSED>
__global__ void __launch_bounds__(256) fp32_pipe_probe_kernel(const float2* __restrict__ xin, const float* __restrict__ taps, float2* __restrict__ out, int iters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    u64 cc[QPSK_PROBE_TAPS];
#pragma unroll
    for (int i = 0; i < QPSK_PROBE_TAPS; i++) cc[i] = pack2(taps[i], taps[i]);
    __shared__ float2 xs[2048 + QPSK_PROBE_WIN];
    for (int i = threadIdx.x; i < 2048 + QPSK_PROBE_WIN; i += blockDim.x) xs[i] = xin[i & 1023];
    __syncthreads();
    u64 xp[QPSK_PROBE_WIN], ap[QPSK_PROBE_R];
#pragma unroll
    for (int r = 0; r < QPSK_PROBE_R; r++) ap[r] = 0ull;
    const int base = (threadIdx.x & 31) * 17 + (threadIdx.x >> 5) * 64;      // conflict-free lane stride, as in the filter's sample tile
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        const int off = (base + it * QPSK_PROBE_TAPS) & 1023;               // a fresh window every pass: nothing is loop-invariant
#pragma unroll
        for (int i = 0; i < QPSK_PROBE_WIN; i++) xp[i] = pack2(xs[off + i].x, xs[off + i].y);
#pragma unroll
        for (int i = 0; i < QPSK_PROBE_TAPS; i++)
#pragma unroll
            for (int r = 0; r < QPSK_PROBE_R; r++)
                ap[r] = FUSED ? fma2(xp[r + i], cc[i], ap[r]) : add2(ap[r], mul2_exact(xp[r + i], cc[i]));
    }
#pragma unroll
    for (int r = 0; r < QPSK_PROBE_R; r++) {
        float2 v;
        unpack2(ap[r], v.x, v.y);
        out[(size_t)t * QPSK_PROBE_R + r] = v;
    }
}
