// fir.cuh -- channel-batched rrc_fir(): complex-float samples x real taps, in place, with the
// caller-visible delay line of the reference (rrc_fir.c:17-30, `memory[NTAPS]`).
//
// Same strip decomposition as the receiver front end (rx_front.cuh): lane = channel, warp =
// 16-sample strip of a 128-sample tile, halo tiles kept in shared memory, taps from the constant
// bank.  One CTA walks its 32 channels through the whole block of samples, so filtering in
// place is safe (a tile is staged to shared memory before its outputs overwrite it).
#pragma once

#include "common.cuh"
#include "rx_front.cuh"   // c_taps2, fir_strip

struct FirArgs {
    float2* data;     // [C][T] complex samples, filtered in place
    float2* state;    // [C][NTAPS] delay line: the last NTAPS inputs, oldest first (rrc_fir's `memory`)
    int C, T;
};

template <int NTAPS>
struct FirSmem {
    static constexpr int HT = (NTAPS - 1 + QPSK_CHUNK - 1) / QPSK_CHUNK;   // halo tiles
    static constexpr int XS = (HT + 1) * QPSK_CHUNK + 1;                    // odd stride: conflict-free 64-bit access
    u64 x[QPSK_GROUP][XS];
};

template <int NTAPS, int MODE>
__global__ void __launch_bounds__(256, 1) fir_kernel(const FirArgs a) {
    constexpr int R = 16;
    constexpr int HT = FirSmem<NTAPS>::HT;
    constexpr int CUR = HT * QPSK_CHUNK;             // first slot of the current tile
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FirSmem<NTAPS>& sm = *reinterpret_cast<FirSmem<NTAPS>*>(smem_raw);

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ch = blockIdx.x * QPSK_GROUP + lane;
    const bool live = ch < a.C;
    const int chl = live ? ch : a.C - 1;
    const int strip = w * R;
    u64* xrow = &sm.x[lane][0];
    u64* data = reinterpret_cast<u64*>(a.data) + (size_t)chl * a.T;
    u64* state = reinterpret_cast<u64*>(a.state) + (size_t)chl * NTAPS;

    // halo <- memory[1 .. NTAPS-1] (memory[0] is never read again by rrc_fir)
    for (int i = w; i < NTAPS - 1; i += 8) xrow[CUR - (NTAPS - 1) + i] = state[1 + i];

    const int ntiles = (a.T + QPSK_CHUNK - 1) / QPSK_CHUNK;
    u64 nx[R];
#pragma unroll
    for (int e = 0; e < R; e++) nx[e] = (strip + e < a.T) ? data[strip + e] : 0ull;

    for (int k = 0; k < ntiles; k++) {
        const int t0 = k * QPSK_CHUNK + strip;
#pragma unroll
        for (int e = 0; e < R; e++) xrow[CUR + strip + e] = nx[e];
        __syncthreads();
        if (k + 1 < ntiles) {
#pragma unroll
            for (int e = 0; e < R; e++) nx[e] = (t0 + QPSK_CHUNK + e < a.T) ? data[t0 + QPSK_CHUNK + e] : 0ull;
        }
        u64 acc[R];
        fir_strip<NTAPS, R, MODE>(xrow + CUR + strip - (NTAPS - 1), acc);
        if (live) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (t0 + r < a.T) {
                    float yr, yi;
                    unpack2(acc[r], yr, yi);
                    data[t0 + r] = pack2(gain_exact(yr), gain_exact(yi));     // rrc_fir.c:28
                }
            }
        }
        __syncthreads();
        if (k + 1 == ntiles) {
            // delay line out: the last NTAPS inputs, memory[i] = x[T - NTAPS + i]
            if (live) {
                const int last = CUR + (a.T - 1 - k * QPSK_CHUNK);         // slot of input T-1
                for (int i = w; i < NTAPS; i += 8) {
                    const int slot = last - (NTAPS - 1) + i;
                    // slots below the restored halo exist only when T < 1: unreachable (T >= 1); the oldest entry
                    // for T < NTAPS comes from the previous delay line shifted by T
                    state[i] = (slot >= CUR - (NTAPS - 1)) ? xrow[slot] : 0ull;
                }
            }
        } else {
            // shift every tile one tile to the left; each thread moves its own strips
#pragma unroll
            for (int h = 0; h < HT; h++)
#pragma unroll
                for (int e = 0; e < R; e++) xrow[h * QPSK_CHUNK + strip + e] = xrow[(h + 1) * QPSK_CHUNK + strip + e];
        }
    }
}
