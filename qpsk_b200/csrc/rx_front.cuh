// rx_front.cuh -- K0 (mixer phasor table) and K1 (mixer + RRC matched filter + timing
// histogram + decimation) of the batched receiver.
//
// Replaces, per channel, the first four stages of rx_frame (reference qpsk.c:114-191) and the
// inner loop of rrc_fir (rrc_fir.c:17-30).  Bit-exact in QPSK_MODE_EXACT.
//
// Work decomposition: one CTA = 32 channels (one per lane) x a block of consecutive frames;
// 8 warps split each 128-sample time tile into 16-sample strips, so a thread owns 16 consecutive
// outputs of one channel and slides over the 126-sample halo kept in shared memory.  Taps are
// uniform across the warp and come from the constant bank as (c,c) pairs for FMUL2/FFMA2.
#pragma once

#include "common.cuh"

enum { QPSK_MODE_EXACT = 0, QPSK_MODE_FAST = 1 };
enum { QPSK_UB_ALIAS = 0, QPSK_UB_CLAMP = 1 };

// taps duplicated into both halves of a 64-bit constant: the packed-FP32 multiplier operand
__constant__ float2 c_taps2[QPSK_MAX_TAPS];

// --------------------------------------------------------------------------------------------
// K0: the mixer phasor sequence of qpsk.c:115,120.  It is a data-independent recurrence
// (phase *= rect per sample, renormalised once per frame), identical for every channel of a
// profile, so one thread evaluates it once per launch and all channels share the table.
//   table[QPSK_CHUNK + j] = fbb_rx_phase used for sample j of this launch (j = 0 .. F*N-1)
//   table[0 .. QPSK_CHUNK-1] = the entries of the 128 samples before it (carried in `tail`)
// --------------------------------------------------------------------------------------------
__global__ void phasor_table_kernel(float2* __restrict__ table, float2* __restrict__ tail,
                                    float2* __restrict__ phase_state, float2 rect, int nframes, int frame_size) {
    const int t = threadIdx.x;
    if (t < QPSK_CHUNK) table[t] = tail[t];
    __syncthreads();
    if (t == 0) {
        float2 ph = *phase_state;
        float2* out = table + QPSK_CHUNK;
        for (int f = 0; f < nframes; f++) {
            for (int i = 0; i < frame_size; i++) {
                ph = cmul_exact(ph, rect);                       // qpsk.c:115
                out[(size_t)f * frame_size + i] = ph;
            }
            // qpsk.c:120  phase /= cabsf(phase); glibc hypotf == (float)sqrt(re^2 + im^2) in double
            const double dr = (double)ph.x, di = (double)ph.y;
            const float mag = __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(di, di))));
            ph.x = __fdiv_rn(ph.x, mag);
            ph.y = __fdiv_rn(ph.y, mag);
        }
        *phase_state = ph;
    }
    __syncthreads();
    if (t < QPSK_CHUNK) tail[t] = table[(size_t)nframes * frame_size + t];
}

// --------------------------------------------------------------------------------------------
// FIR strip: R consecutive outputs of one channel.  `x` points at the input sample that meets
// tap 0 of output 0, i.e. sample (t0 - (NTAPS-1)).  Input d contributes to output r through tap
// d - r, so walking d upwards accumulates every output oldest-tap-first from +0, exactly the
// order of rrc_fir.c:22-26.
// --------------------------------------------------------------------------------------------
template <int NTAPS, int R, int MODE>
__device__ __forceinline__ void fir_strip(const u64* __restrict__ x, u64 (&acc)[R]) {
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = 0ull;   // (+0, +0)
#pragma unroll
    for (int d = 0; d < NTAPS - 1 + R; d++) {
        const u64 xv = x[d];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = d - r;
            if (i >= 0 && i < NTAPS) {
                const u64 cc = *reinterpret_cast<const u64*>(&c_taps2[i]);
                if (MODE == QPSK_MODE_EXACT) acc[r] = add2(acc[r], mul2_exact(xv, cc));
                else acc[r] = fma2(xv, cc, acc[r]);
            }
        }
    }
}

struct RxFrontArgs {
    const int16_t* pcm;        // [C][F*N] s16 PCM, channel-major
    const int16_t* pcm_tail;   // [C][QPSK_CHUNK] the 128 samples before frame 0 (zeros at stream start)
    const float2*  phasor;     // [QPSK_CHUNK + F*N] from phasor_table_kernel
    float2* dec_ring;          // [nslots][nsym][Cpad] decimated symbols, channel-fastest
    int*    index_t;           // [F][Cpad] timing index per frame
    float2* fir_dbg;           // optional [C][F*N] matched-filter output (parity taps), may be null
    int C, Cpad, F, N;
    int frames_per_block;      // frames handled by one CTA
    int slot_base, nslots;     // frame f goes to ring slot (slot_base + 1 + f) % nslots
    int ub_mode;
};

template <int SPS>
struct RxFrontSmem {
    static constexpr int XS = 2 * QPSK_CHUNK + 1;   // odd row stride (in float2) => conflict-free 64-bit access
    static constexpr int OS = 512 + 1;
    u64 x[QPSK_GROUP][XS];        // [0,128) previous tile (halo), [128,256) current tile
    u64 out[QPSK_GROUP][OS];      // matched-filter output of the current frame
    float2 ph[2][QPSK_CHUNK];     // mixer phasors of the current / next tile
    u64 hist[2][QPSK_GROUP];      // 7 x 8-bit amplitude-bin counters for I and for Q
    int index[QPSK_GROUP];
};

// mix 16 PCM samples with their phasors and store them as the current tile: qpsk.c:117
__device__ __forceinline__ void mix_store(u64* __restrict__ xrow_cur, const uint4& p0, const uint4& p1,
                                          const float2* __restrict__ ph) {
    const unsigned w[8] = { p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w };
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const int v = (int)(short)((e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xffffu));
        const float s = __fmul_rn((float)v, 6.103515625e-05f);   // (float)in / 16384.0f, exact
        const float2 p = ph[e];
        xrow_cur[e] = pack2(__fmul_rn(p.x, s), __fmul_rn(p.y, s));
    }
}

template <int NTAPS, int SPS, int MODE>
__global__ void __launch_bounds__(256, 1) rx_front_kernel(const RxFrontArgs a) {
    static_assert(NTAPS - 1 <= QPSK_CHUNK - 2, "halo must fit in one previous tile");
    constexpr int R = 16;
    constexpr int NSYM_MAX = 512 / SPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RxFrontSmem<SPS>& sm = *reinterpret_cast<RxFrontSmem<SPS>*>(smem_raw);

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ngroups = a.Cpad / QPSK_GROUP;
    const int g = blockIdx.x % ngroups, fb = blockIdx.x / ngroups;
    const int f0 = fb * a.frames_per_block;
    const int f1 = min(a.F, f0 + a.frames_per_block);
    if (f0 >= f1) return;
    const int ch = g * QPSK_GROUP + lane;
    const bool live = ch < a.C;
    const int chl = live ? ch : a.C - 1;            // padded lanes recompute the last channel, stores are masked
    const int N = a.N, nsym = N / SPS;
    const int tiles_per_frame = N / QPSK_CHUNK;
    const size_t row = (size_t)a.F * N;
    const int16_t* pcm_row = a.pcm + (size_t)chl * row;
    const int strip = w * R;                         // this thread's 16 samples inside a tile

    u64* xrow = &sm.x[lane][0];
    u64* xcur = xrow + QPSK_CHUNK + strip;

    // ---- prologue: the tile before frame f0 becomes the halo
    {
        const int16_t* src = (f0 == 0) ? a.pcm_tail + (size_t)chl * QPSK_CHUNK + strip
                                       : pcm_row + (size_t)f0 * N - QPSK_CHUNK + strip;
        const uint4 p0 = *reinterpret_cast<const uint4*>(src);
        const uint4 p1 = *reinterpret_cast<const uint4*>(src + 8);
        const float2* ph = a.phasor + (size_t)f0 * N + strip;    // table index QPSK_CHUNK + (f0*N - 128 + strip + e)
        float2 phr[16];
#pragma unroll
        for (int e = 0; e < 16; e++) phr[e] = ph[e];
        mix_store(xcur, p0, p1, phr);
    }
    uint4 n0, n1;   // PCM of the next tile, prefetched across the FIR loop
    {
        const int16_t* src = pcm_row + (size_t)f0 * N + strip;
        n0 = *reinterpret_cast<const uint4*>(src);
        n1 = *reinterpret_cast<const uint4*>(src + 8);
    }
    if (threadIdx.x < QPSK_CHUNK) sm.ph[0][threadIdx.x] = a.phasor[QPSK_CHUNK + (size_t)f0 * N + threadIdx.x];
    __syncthreads();

    const int ntiles = (f1 - f0) * tiles_per_frame;
    for (int k = 0; k < ntiles; k++) {
        const size_t tbase = (size_t)f0 * N + (size_t)k * QPSK_CHUNK;   // first sample of this tile in the launch
        // ---- shift own strip: current -> halo, then mix the prefetched PCM in as the new current tile
#pragma unroll
        for (int e = 0; e < R; e++) xrow[strip + e] = xcur[e];
        mix_store(xcur, n0, n1, &sm.ph[k & 1][strip]);
        __syncthreads();

        // ---- prefetch the next tile's PCM and phasors while the FIR runs
        float2 phn = make_float2(0.f, 0.f);
        if (k + 1 < ntiles) {
            const int16_t* src = pcm_row + tbase + QPSK_CHUNK + strip;
            n0 = *reinterpret_cast<const uint4*>(src);
            n1 = *reinterpret_cast<const uint4*>(src + 8);
            if (threadIdx.x < QPSK_CHUNK) phn = a.phasor[QPSK_CHUNK + tbase + QPSK_CHUNK + threadIdx.x];
        }

        // ---- matched filter: rrc_fir.c:22-28
        u64 acc[R];
        fir_strip<NTAPS, R, MODE>(xcur - (NTAPS - 1), acc);
        const int tf = (k % tiles_per_frame) * QPSK_CHUNK + strip;     // sample index inside the frame
        u64* orow = &sm.out[lane][tf];
#pragma unroll
        for (int r = 0; r < R; r++) {
            float yr, yi;
            unpack2(acc[r], yr, yi);
            orow[r] = pack2(gain_exact(yr), gain_exact(yi));
        }
        if (a.fir_dbg != nullptr && live) {
            u64* dst = reinterpret_cast<u64*>(a.fir_dbg) + (size_t)ch * row + tbase + strip;
#pragma unroll
            for (int r = 0; r < R; r++) dst[r] = orow[r];
        }
        if (threadIdx.x < QPSK_CHUNK) sm.ph[(k + 1) & 1][threadIdx.x] = phn;
        __syncthreads();

        if ((k + 1) % tiles_per_frame != 0) continue;

        // ---- frame complete: amplitude histograms, qpsk.c:131-167.  Warp 0 = I, warp 1 = Q, lane = channel.
        const int f = f0 + k / tiles_per_frame;
        if (w < 2) {
            const float* o = reinterpret_cast<const float*>(&sm.out[lane][0]) + w;
            float av = 0.0f, mx = 0.0f;
            u64 hist = 0ull;
#pragma unroll 4
            for (int s = 0; s < NSYM_MAX; s++) {
#pragma unroll
                for (int j = 0; j < SPS; j++) av = __fadd_rn(av, fabsf(o[2 * (s * SPS + j)]));
                av = __fmul_rn(av, 1.0f / SPS);              // av /= CYCLES, exact for a power of two
                if (av > mx) mx = av;
                const float hv = __fmul_rn(mx, 0.125f);      // max / 8.0f
                int bin = 0;                                   // first k in 1..7 with av <= hv*k (hv*k is monotone in k)
#pragma unroll
                for (int kk = 7; kk >= 1; kk--) bin = (av <= __fmul_rn(hv, (float)kk)) ? kk : bin;
                hist += 1ull << (8 * bin);                     // byte 0 collects "no bin"; counts <= 128 fit a byte
            }
            sm.hist[w][lane] = hist;
        }
        __syncthreads();
        if (w == 0) {                                          // qpsk.c:173-180 first strict maximum
            const u64 hi = sm.hist[0][lane], hq = sm.hist[1][lane];
            int hmax = 0, index = 0;
#pragma unroll
            for (int kk = 1; kk < 8; kk++) {
                const int h = (int)((hi >> (8 * kk)) & 0xff) + (int)((hq >> (8 * kk)) & 0xff);
                if (h > hmax) { hmax = h; index = kk; }
            }
            sm.index[lane] = index;
            if (live) a.index_t[(size_t)f * a.Cpad + ch] = index;
        }
        __syncthreads();
        // ---- decimate, qpsk.c:186-191: symbol i = sample i*SPS + index, stored channel-fastest
        {
            const int index = sm.index[lane];
            const int slot = (a.slot_base + 1 + f) % a.nslots;
            u64* dst = reinterpret_cast<u64*>(a.dec_ring) + (size_t)slot * nsym * a.Cpad + ch;
            for (int i = w; i < nsym; i += 8) {
                const int j = i * SPS + index;
                u64 v;
                if (j < N) v = sm.out[lane][j];
                else if (a.ub_mode == QPSK_UB_CLAMP) v = sm.out[lane][N - 1];
                else v = 0ull;   // aliasing read of decimated_frame[j-N]: patched by the Costas kernel
                if (live) dst[(size_t)i * a.Cpad] = v;
            }
        }
        __syncthreads();
    }
}
