// fir.cuh -- channel-batched rrc_fir(): complex-float samples x real taps, in place, with the
// caller-visible delay line of the reference (rrc_fir.c:17-30, `memory[NTAPS]`).
//
// Same strip decomposition as the receiver front end (rx_front.cuh): lane = channel, warp =
// 16-sample strip of a tile, halo tiles kept in shared memory, taps from the constant bank.
// Tiles enter and leave through a row-contiguous staging buffer (cp.async in, 128-bit stores out),
// so global traffic is whole 512-byte runs of one channel instead of 32 rows per instruction.
// One CTA walks its 32 channels through the whole block of samples, so filtering in place is
// safe (tile k+1 is staged before tile k's outputs are stored, and tiles do not overlap).
#pragma once

#include "common.cuh"
#include "rx_front.cuh"   // TapBank, fir_strip

struct FirArgs {
    float2* data;     // [C][T] complex samples, filtered in place
    float2* state;    // [C][NTAPS] delay line: the last NTAPS inputs, oldest first (rrc_fir's `memory`)
    const float2* halo;  // [C][nblocks][NTAPS-1] inputs preceding every time block, saved by fir_save_halo_kernel (nblocks > 1)
    int C, T;
    int nblocks, tiles_per_block;   // blockIdx.y = time block: the call is cut so the grid fills whole waves of CTAs
};

// The filter works in place, so the NTAPS-1 inputs in front of every time block are copied aside before the main
// kernel overwrites them; for block 0 they are memory[1 .. NTAPS-1] of the delay line, which the CTA of the last
// time block rewrites (no CTA of the main kernel reads `state` when the call is cut into blocks).
__global__ void __launch_bounds__(128) fir_save_halo_kernel(const float2* __restrict__ data, const float2* __restrict__ state,
                                                            float2* __restrict__ halo, int T, int halo_len, int nblocks, int block_len) {
    const int ch = blockIdx.x, b = blockIdx.y;
    const float2* src = b == 0 ? state + (size_t)ch * (halo_len + 1) + 1 : data + (size_t)ch * T + (size_t)b * block_len - halo_len;
    float2* dst = halo + ((size_t)ch * nblocks + b) * halo_len;
    for (int i = threadIdx.x; i < halo_len; i += blockDim.x) dst[i] = src[i];
}

// NW filter warps per CTA, each a 16-sample strip of a tile of TILE = 16 * NW samples.
template <int NTAPS, int NW>
struct FirSmem {
    static constexpr int TILE = 16 * NW;
    static constexpr int HT = (NTAPS - 1 + TILE - 1) / TILE;    // halo tiles
    static constexpr int XS = (HT + 1) * TILE + 1;               // odd stride: conflict-free 64-bit access
    static constexpr int IOS = TILE + 2;                         // rows of an odd number of 16-byte units: aligned, conflict-free 128-bit access
    u64 x[QPSK_GROUP][XS];
    u64 io[QPSK_GROUP][IOS];     // landing zone of the next tile's inputs, then this tile's outputs on their way out
};
// 127 taps, NW = 8: 99 KB, two CTAs of 8 warps per SM (one filters while the other sits at a barrier or moves a tile).
// 256 taps, NW = 16: 198 KB, one CTA of 16 warps: the same 4 filter warps per scheduler, which is what keeps the
// FP32 pipe fed across the LDS/constant-fetch latency at the top of each trip.

// zero-filling forms: bytes past `nbytes` are written as zero, nothing is read when nbytes == 0
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, int nbytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async8_zfill(void* smem_dst, const void* gmem_src, int nbytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gmem_src), "r"(nbytes) : "memory");
}

// Tile `k`, this warp's rows (32 / NW of them) -> io, row-contiguous: one warp instruction moves 512 contiguous bytes
// of one channel.  vec16: rows are 16-byte aligned (T even and an aligned base), else 8-byte pieces.
template <int NTAPS, int NW>
__device__ __forceinline__ void fir_load_tile(FirSmem<NTAPS, NW>& sm, const FirArgs& a, const int k, const int w, const int lane, const bool vec16) {
    constexpr int TILE = 16 * NW, RPW = QPSK_GROUP / NW;
    const u64* base = reinterpret_cast<const u64*>(a.data);
#pragma unroll
    for (int rr = 0; rr < RPW; rr++) {
        const int row = RPW * w + rr;
        const int ch = min(blockIdx.x * QPSK_GROUP + row, a.C - 1);
        const u64* src = base + (size_t)ch * a.T;
        if (vec16) {
#pragma unroll
            for (int h = 0; h < TILE / 64; h++) {
                const int j = 2 * (lane + 32 * h), t = k * TILE + j;      // two samples per piece
                const int left = a.T - t;
                cp_async16_zfill(&sm.io[row][j], src + (left > 0 ? t : 0), left > 0 ? 16 : 0);
            }
        } else {
#pragma unroll
            for (int h = 0; h < TILE / 32; h++) {
                const int j = lane + 32 * h, t = k * TILE + j;
                const int left = a.T - t;
                cp_async8_zfill(&sm.io[row][j], src + (left > 0 ? t : 0), left > 0 ? 8 : 0);
            }
        }
    }
}

// outputs of tile `k`, this warp's rows, io -> global, row-contiguous
template <int NTAPS, int NW>
__device__ __forceinline__ void fir_store_tile(FirSmem<NTAPS, NW>& sm, const FirArgs& a, const int k, const int w, const int lane, const bool vec16) {
    constexpr int TILE = 16 * NW, RPW = QPSK_GROUP / NW;
    u64* base = reinterpret_cast<u64*>(a.data);
#pragma unroll
    for (int rr = 0; rr < RPW; rr++) {
        const int row = RPW * w + rr;
        const int ch = blockIdx.x * QPSK_GROUP + row;
        if (ch >= a.C) continue;
        u64* dst = base + (size_t)ch * a.T;
        if (vec16) {
#pragma unroll
            for (int h = 0; h < TILE / 64; h++) {
                const int j = 2 * (lane + 32 * h), t = k * TILE + j;
                if (t < a.T) *reinterpret_cast<uint4*>(dst + t) = *reinterpret_cast<const uint4*>(&sm.io[row][j]);
            }
        } else {
#pragma unroll
            for (int h = 0; h < TILE / 32; h++) {
                const int j = lane + 32 * h, t = k * TILE + j;
                if (t < a.T) dst[t] = sm.io[row][j];
            }
        }
    }
}

template <int NTAPS, int MODE, int NW>
__global__ void __launch_bounds__(32 * NW, (sizeof(FirSmem<NTAPS, NW>) <= 112 * 1024) ? 2 : 1) fir_kernel(const __grid_constant__ FirArgs a, const __grid_constant__ TapBank<NTAPS> tb) {
    constexpr int R = 16, TILE = 16 * NW;
    constexpr int HT = FirSmem<NTAPS, NW>::HT;
    constexpr int CUR = HT * TILE;                   // first slot of the current tile
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FirSmem<NTAPS, NW>& sm = *reinterpret_cast<FirSmem<NTAPS, NW>*>(smem_raw);

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ch = blockIdx.x * QPSK_GROUP + lane;
    const bool live = ch < a.C;
    const int chl = live ? ch : a.C - 1;
    const int strip = w * R;
    const bool vec16 = ((a.T & 1) == 0) && ((reinterpret_cast<size_t>(a.data) & 15) == 0);
    u64* xrow = &sm.x[lane][0];
    u64* iorow = &sm.io[lane][0];
    u64* state = reinterpret_cast<u64*>(a.state) + (size_t)chl * NTAPS;
    const int ntiles = (a.T + TILE - 1) / TILE;
    const int k0 = blockIdx.y * a.tiles_per_block, k1 = min(ntiles, k0 + a.tiles_per_block);

    // first tile on its way; halo <- memory[1 .. NTAPS-1] (memory[0] is never read again by rrc_fir) or, for a
    // later time block, the inputs saved by the pre-pass
    fir_load_tile<NTAPS, NW>(sm, a, k0, w, lane, vec16);
    const u64* before = a.nblocks == 1 ? state + 1
        : reinterpret_cast<const u64*>(a.halo) + ((size_t)chl * a.nblocks + blockIdx.y) * (NTAPS - 1);
    for (int i = w; i < NTAPS - 1; i += NW) xrow[CUR - (NTAPS - 1) + i] = before[i];
    cp_async_wait_all();
    __syncthreads();
#pragma unroll
    for (int e = 0; e < R; e += 2) {
        const uint4 v = *reinterpret_cast<const uint4*>(&iorow[strip + e]);
        xrow[CUR + strip + e] = ((u64)v.y << 32) | v.x;
        xrow[CUR + strip + e + 1] = ((u64)v.w << 32) | v.z;
    }
    __syncthreads();

    for (int k = k0; k < k1; k++) {
        const bool last = k + 1 == k1;
        // x holds tile k (and its halo); io is free: the next tile starts to land while this one is filtered.  The rows a
        // warp loads are the rows it stored at the end of the previous trip, so a warp-level fence orders the two.
        __syncwarp();
        if (!last) fir_load_tile<NTAPS, NW>(sm, a, k + 1, w, lane, vec16);
        u64 acc[R];
        fir_strip<NTAPS, R, MODE>(xrow + CUR + strip - (NTAPS - 1), tb.t, acc);
        cp_async_wait_all();
        __syncthreads();                               // tile k+1 has landed; nobody reads x any more
        if (last) {
            // delay line out: the last NTAPS inputs, memory[i] = x[T - NTAPS + i]
            if (live && k1 == ntiles) {
                const int lastslot = CUR + (a.T - 1 - k * TILE);             // slot of input T-1
                for (int i = w; i < NTAPS; i += NW) {
                    const int slot = lastslot - (NTAPS - 1) + i;
                    state[i] = (slot >= CUR - (NTAPS - 1)) ? xrow[slot] : 0ull;
                }
            }
        } else {
            // every tile moves one tile to the left, the landed tile becomes the current one; own strips only
#pragma unroll
            for (int h = 0; h < HT; h++)
#pragma unroll
                for (int e = 0; e < R; e++) xrow[h * TILE + strip + e] = xrow[(h + 1) * TILE + strip + e];
#pragma unroll
            for (int e = 0; e < R; e += 2) {
                const uint4 v = *reinterpret_cast<const uint4*>(&iorow[strip + e]);
                xrow[CUR + strip + e] = ((u64)v.y << 32) | v.x;
                xrow[CUR + strip + e + 1] = ((u64)v.w << 32) | v.z;
            }
        }
        __syncthreads();                               // io has been consumed (and x is complete for the next trip)
#pragma unroll
        for (int r = 0; r < R; r += 2) {               // rrc_fir.c:28
            float y0r, y0i, y1r, y1i;
            unpack2(acc[r], y0r, y0i);
            unpack2(acc[r + 1], y1r, y1i);
            uint4 v;
            v.x = __float_as_uint(gain_exact(y0r)); v.y = __float_as_uint(gain_exact(y0i));
            v.z = __float_as_uint(gain_exact(y1r)); v.w = __float_as_uint(gain_exact(y1i));
            *reinterpret_cast<uint4*>(&iorow[strip + r]) = v;
        }
        __syncthreads();                               // the tile's outputs are complete in io
        fir_store_tile<NTAPS, NW>(sm, a, k, w, lane, vec16);
    }
}
