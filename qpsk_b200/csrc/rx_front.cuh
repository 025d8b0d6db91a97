// rx_front.cuh -- K0 (mixer phasor table) and K1 (mixer + RRC matched filter + timing
// histogram + decimation) of the batched receiver.
//
// Replaces, per channel, the first four stages of rx_frame (reference qpsk.c:114-191) and the
// inner loop of rrc_fir (rrc_fir.c:17-30).  Bit-exact in QPSK_MODE_EXACT.
//
// Work decomposition: one CTA = 32 channels (one per lane) x a block of consecutive frames, two CTAs
// per SM.  8 filter warps split each 128-sample time tile into 16-sample strips, so a thread owns 16
// consecutive outputs of one channel and slides over the 126-sample halo kept in shared memory; two
// timing warps and one Costas warp follow behind (see rx_front_kernel).  Taps are uniform across the
// warp and come from the constant bank as (c,c) pairs for FMUL2/FFMA2.
#pragma once

#include "common.cuh"
#include "rx_costas.cuh"

enum { QPSK_MODE_EXACT = 0, QPSK_MODE_FAST = 1, QPSK_MODE_IEEE = 2 };
enum { QPSK_UB_ALIAS = 0, QPSK_UB_CLAMP = 1, QPSK_UB_PHASE = 2, QPSK_UB_TAU = 3 };

// Taps duplicated into both halves of a 64-bit constant: the packed-FP32 multiplier operand.  The bank is a
// __grid_constant__ kernel parameter, i.e. it lives in the launch's own slice of the constant bank: every
// context (receiver, filter bank, transmitter) launches with its own taps, on any device and any stream, and
// nothing process-wide has to be re-uploaded or ordered against kernels still in flight.
template <int NTAPS>
struct TapBank {
    float2 t[NTAPS];
    float2 one;        // (1, 1): the operand that makes an FFMA2 a plain rounded add, see fir_tap<QPSK_MODE_IEEE>
};

// --------------------------------------------------------------------------------------------
// K0: the mixer phasor sequence of qpsk.c:115,120.  It is a data-independent recurrence
// (phase *= rect per sample, renormalised once per frame), identical for every channel of a
// profile, so one thread evaluates it once per launch and all channels share the table.
//   table[QPSK_CHUNK + j] = fbb_rx_phase used for sample j of this launch (j = 0 .. F*N-1)
//   table[0 .. QPSK_CHUNK-1] = the entries of the 128 samples before it (the end of the previous call's table)
// Two tables alternate, so the table of the next call can be evaluated on a side stream while this call runs.
// --------------------------------------------------------------------------------------------
__global__ void phasor_table_kernel(const float2* __restrict__ prev_table, int prev_frames, const float2* __restrict__ state_in,
                                    float2* __restrict__ table, float2* __restrict__ state_out, float2 rect, int nframes, int frame_size) {
    const int t = threadIdx.x;
    // the 128 entries before sample 0 are the last 128 entries of the previous call's table
    if (t < QPSK_CHUNK) table[t] = prev_table[(size_t)prev_frames * frame_size + t];
    if (t == 0) {
        float2 ph = *state_in;
        float2* out = table + QPSK_CHUNK;
        for (int f = 0; f < nframes; f++) {
            float2* row = out + (size_t)f * frame_size;
#pragma unroll 16
            for (int i = 0; i < frame_size; i++) {
                ph = cmul_exact(ph, rect);                       // qpsk.c:115
                row[i] = ph;
            }
            // qpsk.c:120  phase /= cabsf(phase); glibc hypotf == (float)sqrt(re^2 + im^2) in double
            const double dr = (double)ph.x, di = (double)ph.y;
            const float mag = __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(di, di))));
            ph.x = __fdiv_rn(ph.x, mag);
            ph.y = __fdiv_rn(ph.y, mag);
        }
        *state_out = ph;
    }
}

// --------------------------------------------------------------------------------------------
// FIR strip: R consecutive outputs of one channel.  `x` points at the input sample that meets
// tap 0 of output 0, i.e. sample (t0 - (NTAPS-1)).  Input d contributes to output r through tap
// d - r, so walking d upwards accumulates every output oldest-tap-first from +0, exactly the
// order of rrc_fir.c:22-26.
// --------------------------------------------------------------------------------------------
// QPSK_MODE_EXACT: FMUL2.FTZ + FADD2, the receiver's form (its products are never subnormal, see mul2_exact).
// QPSK_MODE_IEEE : FMUL2 + FFMA2(acc, (1, 1), product) -- acc * 1 + p rounds once, exactly like acc + p, for every input
//                  including subnormal ones, and a multiply cannot be contracted into the ADDEND of an FMA: the form of
//                  the general rrc_fir entry points, which must not flush (rrc_fir.c:24-26 does not).  Same two packed
//                  instructions per tap.  `one` arrives as a kernel argument so that it stays opaque to the optimiser.
// QPSK_MODE_FAST : FFMA2, fused (not bit-exact).
template <int MODE>
__device__ __forceinline__ void fir_tap(u64& acc, const u64 xv, const float2* __restrict__ taps2, const int i, const u64 one) {
    const u64 cc = *reinterpret_cast<const u64*>(&taps2[i]);
    if (MODE == QPSK_MODE_EXACT) acc = add2(acc, mul2_exact(xv, cc));
    else if (MODE == QPSK_MODE_IEEE) acc = fma2(acc, one, mul2(xv, cc));
    else acc = fma2(xv, cc, acc);
}

// The walk over d has a triangular head (d < R-1: only outputs 0..d are reached), a steady part
// where every output takes a tap, and a triangular tail.  The steady part is a rolled loop of R
// steps per trip: accumulators stay put, the sample address is base + immediate and the tap index
// is (uniform loop base) + immediate, so nothing rotates and the body (R*R tap updates, ~8 KB of
// code at R = 16) stays resident in the instruction cache -- the fully unrolled 70 KB version spent
// 9 % of its issue slots waiting for instruction fetch (profiles/r01_rx_front_v2).
template <int NTAPS, int R, int MODE>
__device__ __forceinline__ void fir_strip(const u64* __restrict__ x, const float2* __restrict__ taps2, u64 (&acc)[R]) {
    constexpr int STEADY = NTAPS - R + 1;            // d = R-1 .. NTAPS-1
    constexpr int TRIPS = STEADY / R, REM = STEADY % R;
    const u64 one = *reinterpret_cast<const u64*>(&taps2[NTAPS]);      // TapBank::one
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = 0ull;       // (+0, +0)
#pragma unroll
    for (int d = 0; d < R - 1; d++) {                // head
        const u64 xv = x[d];
#pragma unroll
        for (int r = 0; r <= d; r++) fir_tap<MODE>(acc[r], xv, taps2, d - r, one);
    }
#ifndef QPSK_STRIP_UNROLL
#define QPSK_STRIP_UNROLL 1
#endif
    constexpr int STRIP_UNROLL = QPSK_STRIP_UNROLL;
#pragma unroll STRIP_UNROLL
    for (int m = 0; m < TRIPS; m++) {                // steady, rolled
        const int d0 = R - 1 + m * R;
#pragma unroll
        for (int e = 0; e < R; e++) {
            const u64 xv = x[d0 + e];
#pragma unroll
            for (int r = 0; r < R; r++) fir_tap<MODE>(acc[r], xv, taps2, d0 + e - r, one);
        }
    }
#pragma unroll
    for (int d = R - 1 + TRIPS * R; d < R - 1 + TRIPS * R + REM; d++) {   // steady remainder
        const u64 xv = x[d];
#pragma unroll
        for (int r = 0; r < R; r++) fir_tap<MODE>(acc[r], xv, taps2, d - r, one);
    }
#pragma unroll
    for (int d = NTAPS; d < NTAPS - 1 + R; d++) {    // tail
        const u64 xv = x[d];
#pragma unroll
        for (int r = d - (NTAPS - 1); r < R; r++) fir_tap<MODE>(acc[r], xv, taps2, d - r, one);
    }
}

#ifdef QPSK_FRONT_PROF
// instrumented build (tools/front_prof.py): per CTA, global-timer stamps of the roles and cycle sums of the filter warps' phases
#define QPSK_FRONT_PROF_ROWS 4096
__device__ unsigned long long g_front_prof[QPSK_FRONT_PROF_ROWS][48];
// per-tile trace of the CTAs that ran on SM 0 and SM 77: [slot][tile][warp 0..9][start, end] in SM cycles, + a header row per slot
#define QPSK_FRONT_TRACE_SLOTS 128
__device__ long long g_front_trace[QPSK_FRONT_TRACE_SLOTS][256][10][2];
__device__ unsigned long long g_front_trace_hdr[QPSK_FRONT_TRACE_SLOTS][4];
__device__ int g_front_trace_count;
#endif

struct RxFrontArgs {
    const int16_t* pcm;        // [C][pcm_row] s16 PCM, channel-major; the F*N samples of this launch start each row
    size_t pcm_row;            // row stride in samples (>= F*N: a frame chunk of a longer call, or a staged slice)
    const int16_t* pcm_tail;   // [C][QPSK_CHUNK] the 128 samples before frame 0 (zeros at stream start)
    const float2*  phasor;     // [QPSK_CHUNK + F*N] from phasor_table_kernel
    float2* dec_ring;          // [nslots][nsym][Cpad] decimated symbols, channel-fastest
    int*    index_t;           // [F][Cpad] timing index per frame
    float2* fir_dbg;           // optional [C][F*N] matched-filter output (parity taps), may be null
    float2* timing_t;          // optional [F][Cpad] spectral-line timing statistic per frame (extension), may be null
    float*  scratch;           // [nsm * QPSK_SCRATCH_SLOTS + grid][512][2][32] the current frame's filter output per CTA: a CTA claims one of
                               // its SM's slots (the first nsm * SLOTS regions: rewritten every frame by whoever is resident, so they
                               // live in L2 and never travel to HBM) and falls back to the region of its block index if none is free
    int*    scratch_slots;     // [nsm * QPSK_SCRATCH_SLOTS] 0 = free (claimed with atomicCAS by the CTA's timing warps, released when they finish)
    int     scratch_nslots;    // nsm * QPSK_SCRATCH_SLOTS
    int C, Cpad, F, N;
    int chan_base, chan_count; // this launch covers channels [chan_base, chan_base + chan_count); pcm is indexed from chan_base
    int frames_per_block;      // frames handled by one CTA
    int slot_base, nslots;     // frame f goes to ring slot (slot_base + 1 + f) % nslots
    int ub_mode;
    int fuse_costas;           // 1: this CTA owns every frame of its channels, so its Costas warp runs the loop too
                               // 2: this CTA owns a frame BLOCK and its Costas warp still runs the loop: it takes the loop state
                               //    from the CTA of the previous block (relay word of the channel group in block_progress)
    int relay_group_base;      // fuse_costas == 2: relay word of this CTA's group = block_progress[relay_group_base + g]
    int* relay_watchdog;       // ... where a relay wait that gave up leaves relay_ticket (reported by qpsk_b200_rx_sync)
    int relay_ticket;
    unsigned long long* block_progress;  // optional [grid]: (ticket << 32 | frames of this CTA whose symbols are in the ring), for costas_follow_kernel
    unsigned long long progress_ticket;   // ticket << 32 of this launch
    int*    dephase_counters;  // [nsm] zeroed before the launch: arrival order of the first wave's CTAs on each SM
    int     dephase_cycles;    // the first-wave CTA that arrives second on its SM starts this many cycles late (see rx_front_kernel)
    CostasArgs costas;         // used when fuse_costas
};

template <int SPS>
struct RxFrontSmem {
    static constexpr int XS = 2 * QPSK_CHUNK + 1;   // odd row stride (in float2) => conflict-free 64-bit access
    static constexpr int OS = QPSK_CHUNK + 1;
    u64 x[QPSK_GROUP][XS];        // [0,128) previous tile (halo), [128,256) current tile
    u64 out[QPSK_GROUP][OS];      // raw matched-filter sums of ONE tile, handed from the filter warps to the timing warps
    float2 ph[2][QPSK_CHUNK];     // mixer phasors of the current / next tile
    short pcm[QPSK_GROUP][QPSK_CHUNK + 8];   // cp.async landing zone for the next tile's PCM (row stride 272 B: conflict-free 16 B reads)
    u64 hist[2][QPSK_GROUP];      // 7 x 8-bit amplitude-bin counters for I and for Q
    float2 tsum[2][QPSK_GROUP];   // per-component halves of the timing statistic (extension)
    int index[QPSK_GROUP];
    volatile int frames_decimated; // frames whose symbols are in the ring (producer: timing warps, consumer: Costas warp)
    int scr_slot;                  // the scratch slot this CTA claimed, -1 = private fallback region
};
// 110 KB: two CTAs per SM.  While one CTA sits at a barrier or in its fill phase the other one keeps the FP32
// pipe busy, and two Costas warps per SM run concurrently, so the loop's latency no longer paces the filter.

#ifndef QPSK_L2_PCM_HINT
#define QPSK_L2_PCM_HINT 1
#endif
#ifndef QPSK_L2_STORE_HINT
#define QPSK_L2_STORE_HINT 1
#endif
#ifndef QPSK_SCRATCH_SLOT_CLAIM
#define QPSK_SCRATCH_SLOT_CLAIM 1
#endif
// PCM is read exactly once: it enters L2 with the evict-first priority, so that streaming 4 GiB of it through does not push
// out the frame scratch and the symbol ring, which are re-read within a frame's time
__device__ __forceinline__ u64 l2_evict_first_policy() {
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// ... and the two buffers that ARE re-read within a frame's time, the frame scratch and the symbol ring, are stored with the
// evict-last priority
__device__ __forceinline__ u64 l2_evict_last_policy() {
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_keep_f4(float4* dst, const float4 v, u64 policy) {
    if (!QPSK_L2_STORE_HINT) { *dst = v; return; }
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy) : "memory");
}
__device__ __forceinline__ void st_keep_f2(float2* dst, const float2 v, u64 policy) {
    if (!QPSK_L2_STORE_HINT) { *dst = v; return; }
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(dst), "f"(v.x), "f"(v.y), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async16_stream(void* smem_dst, const void* gmem_src, u64 policy) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (!QPSK_L2_PCM_HINT) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory"); return; }
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

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

// one term of the timing statistic: p = y^2 of sample n (n % SPS == j) times e^{-2 pi i j / SPS}, added to (re, im).
// Every operation is a single rounded float op in this order (the oracle restates it: orc_timing_sum).
template <int SPS>
__device__ __forceinline__ void timing_accumulate(const int j, const float p, float& re, float& im) {
    if (SPS == 4) {
        if (j == 0) re = __fadd_rn(re, p);
        else if (j == 1) im = __fsub_rn(im, p);
        else if (j == 2) re = __fsub_rn(re, p);
        else im = __fadd_rn(im, p);
    } else {
        const float t = __fmul_rn(p, 0.70710678118654752f);
        switch (j) {
            case 0: re = __fadd_rn(re, p); break;
            case 1: re = __fadd_rn(re, t); im = __fsub_rn(im, t); break;
            case 2: im = __fsub_rn(im, p); break;
            case 3: re = __fsub_rn(re, t); im = __fsub_rn(im, t); break;
            case 4: re = __fsub_rn(re, p); break;
            case 5: re = __fsub_rn(re, t); im = __fadd_rn(im, t); break;
            case 6: im = __fadd_rn(im, p); break;
            default: re = __fadd_rn(re, t); im = __fadd_rn(im, t); break;
        }
    }
}

// Extension (QPSK_UB_TAU): the sampling phase as the sample nearest to the eye's maximum, round(tau) mod SPS with
// tau = -arg(S) SPS / (2 pi): the sector of S, found with comparisons only (one rounded multiply at 8 samples per symbol), so
// the oracle decides identically (orc_tau_index).
template <int SPS>
__device__ __forceinline__ int tau_index(const float re, const float im) {
    const float ax = fabsf(re), ay = fabsf(im);
    if (SPS == 4) {
        if (ax >= ay) return re >= 0.0f ? 0 : 2;
        return im < 0.0f ? 1 : 3;
    } else {
        const float T = 0.41421356237309503f;            // tan(pi / 8)
        if (ay <= __fmul_rn(T, ax)) return re >= 0.0f ? 0 : 4;
        if (ax <= __fmul_rn(T, ay)) return im < 0.0f ? 2 : 6;
        if (im < 0.0f) return re > 0.0f ? 1 : 3;
        return re > 0.0f ? 7 : 5;
    }
}

// named barriers (id 0 is __syncthreads)
__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

enum { BAR_ROWS = 1,      // FIR warps: sample rows of the tile are in shared memory
       BAR_FIR_DONE = 2,  // FIR warps: every strip has finished reading the rows
       BAR_FULL = 3,      // the raw sums of a tile are in sm.out (FIR arrive, timing sync)
       BAR_EMPTY = 4,     // the timing warps have consumed sm.out (timing arrive, FIR sync)
       BAR_AUX = 5 };     // the two timing warps among themselves
#define QPSK_FIR_THREADS 256
#define QPSK_AUX_THREADS 64
#define QPSK_FRONT_THREADS 352   // 8 FIR warps + 2 timing/decimation warps + 1 Costas warp
#ifndef QPSK_COSTAS_POLL_NS
#define QPSK_COSTAS_POLL_NS 2000
#endif
#define QPSK_SCRATCH_SLOTS 2     // resident CTAs per SM (__launch_bounds__ below)
// one scratch region: 48 chunks x 16 samples x 32 lanes of raw (I, Q) sums = 6 tiles (rx_front2_kernel); the round-1 kernel keeps one
// frame of gained samples (512 x 2 x 32 floats) in the first two thirds of it
#define QPSK_SCRATCH_REGION_FLOATS (48 * 16 * QPSK_GROUP * 2)

// Warp-specialised front end, two CTAs per SM.  Warps 0-7 mix and filter (the FP32-pipe-bound part) and hand
// each tile's raw sums to warps 8-9 through a single-tile shared-memory buffer (named barriers BAR_FULL /
// BAR_EMPTY); warps 8-9 apply the output gain, keep the filtered frame in the CTA's L2-resident scratch, run
// the amplitude histograms, and at the end of a frame compute the index and decimate; warp 10 runs the Costas
// loop of the previous frame when the CTA owns whole streams.  The second CTA on the SM fills the pipe slots
// this one leaves at its barriers and fill phases.
template <int NTAPS, int SPS, int MODE>
__global__ void __launch_bounds__(QPSK_FRONT_THREADS, 2) rx_front_kernel(const __grid_constant__ RxFrontArgs a, const __grid_constant__ TapBank<NTAPS> tb) {
    static_assert(NTAPS - 1 <= QPSK_CHUNK - 2, "halo must fit in one previous tile");
    constexpr int R = 16;
    constexpr int NSYM = 512 / SPS, TILE_SYMS = QPSK_CHUNK / SPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RxFrontSmem<SPS>& sm = *reinterpret_cast<RxFrontSmem<SPS>*>(smem_raw);

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ngroups = (a.chan_count + QPSK_GROUP - 1) / QPSK_GROUP;
    const int g = blockIdx.x % ngroups, fb = blockIdx.x / ngroups;
    const int f0 = fb * a.frames_per_block;
    const int f1 = min(a.F, f0 + a.frames_per_block);
    if (f0 >= f1) return;
    const int ch = a.chan_base + g * QPSK_GROUP + lane;
    const int ch_end = min(a.C, a.chan_base + a.chan_count);
    const bool live = ch < ch_end;
    const int chl = live ? ch : ch_end - 1;         // padded lanes recompute the last channel, stores are masked
    const int N = a.N;
    constexpr int tiles_per_frame = 512 / QPSK_CHUNK;
    const int nframes = f1 - f0;
    if (threadIdx.x == 0) sm.frames_decimated = 0;
#ifdef QPSK_FRONT_PROF
    unsigned long long g_entry;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_entry));
    unsigned long long* prof_row = g_front_prof[blockIdx.x % QPSK_FRONT_PROF_ROWS];
    __shared__ int trace_slot_s;
    if (threadIdx.x == 0) {
        unsigned smid_; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_)); prof_row[0] = g_entry; prof_row[1] = smid_;
        int ts = -1;
        if (smid_ == 0 || smid_ == 77) {
            ts = atomicAdd(&g_front_trace_count, 1) % QPSK_FRONT_TRACE_SLOTS;
            g_front_trace_hdr[ts][0] = g_entry; g_front_trace_hdr[ts][1] = smid_; g_front_trace_hdr[ts][2] = blockIdx.x; g_front_trace_hdr[ts][3] = clock64();
        }
        trace_slot_s = ts;
    }
    __syncthreads();
    const int trace_slot = trace_slot_s;
#define PROF_TRACE(tile, wi, se, val) do { if (trace_slot >= 0 && lane == 0 && (tile) < 256) g_front_trace[trace_slot][tile][wi][se] = (val); } while (0)
#define PROF_ROLE_END(slot) do { if (lane == 0) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); prof_row[slot] = g_; } } while (0)
#else
#define PROF_ROLE_END(slot) do { } while (0)
#endif
    // Two CTAs share an SM and take the same time per tile, so the pair launched together by the first wave would walk
    // through its fill phases (no FP32 work) in lock-step for the whole launch, and every later pair inherits the phase of
    // the pair it replaces.  The first-wave CTA that finds itself second on its SM starts half a tile late instead, so one
    // CTA's fill phase lies under the other one's filter.
    if (a.dephase_cycles > 0 && blockIdx.x < gridDim.x && (int)blockIdx.x < a.scratch_nslots && a.dephase_counters != nullptr) {
        if (threadIdx.x == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            const int order = atomicAdd(&a.dephase_counters[smid], 1);
            if (order & 1) {
                const long long t0 = clock64();
                while (clock64() - t0 < a.dephase_cycles) __nanosleep(500);
            }
        }
    }
    __syncthreads();

    if (w < 8) {
        // =================================== FIR warps ===================================
        const int16_t* pcm_row = a.pcm + (size_t)(chl - a.chan_base) * a.pcm_row;
        const int strip = w * R;                     // this thread's 16 samples inside a tile
        u64* xrow = &sm.x[lane][0];
        u64* xcur = xrow + QPSK_CHUNK + strip;

        // prologue: the tile before frame f0 becomes the halo
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
        // PCM and phasors of the next tile travel HBM -> shared memory by cp.async while the filter runs
        short* stage = &sm.pcm[lane][strip];
        const size_t s0 = (size_t)f0 * N;
        const int ntiles = nframes * tiles_per_frame;
        const u64 pcm_policy = l2_evict_first_policy();
        {
            const int16_t* src = pcm_row + s0 + strip;
            cp_async16_stream(stage, src, pcm_policy);
            cp_async16_stream(stage + 8, src + 8, pcm_policy);
            if (threadIdx.x < QPSK_CHUNK) cp_async8(&sm.ph[0][threadIdx.x], &a.phasor[QPSK_CHUNK + s0 + threadIdx.x]);
        }
        cp_async_wait_all();
        bar_sync(BAR_FIR_DONE, QPSK_FIR_THREADS);

#ifdef QPSK_FRONT_PROF
        long long pc_fill = 0, pc_strip = 0, pc_empty = 0, pc_post = 0, pt, pt2;
        unsigned long long g64 = 0, g65 = 0;
        const long long pstart = clock64();
#endif
        for (int k = 0; k < ntiles; k++) {
            const size_t tbase = s0 + (size_t)k * QPSK_CHUNK;   // first sample of this tile in the launch
#ifdef QPSK_FRONT_PROF
            pt = clock64();
#endif
            // shift own strip: current -> halo, then mix the staged PCM in as the new current tile
#pragma unroll
            for (int e = 0; e < R; e++) xrow[strip + e] = xcur[e];
            {
                const uint4 n0 = *reinterpret_cast<const uint4*>(stage);
                const uint4 n1 = *reinterpret_cast<const uint4*>(stage + 8);
                mix_store(xcur, n0, n1, &sm.ph[k & 1][strip]);
            }
            bar_sync(BAR_ROWS, QPSK_FIR_THREADS);

            if (k + 1 < ntiles) {
                const int16_t* src = pcm_row + tbase + QPSK_CHUNK + strip;
                cp_async16_stream(stage, src, pcm_policy);
                cp_async16_stream(stage + 8, src + 8, pcm_policy);
                if (threadIdx.x < QPSK_CHUNK) cp_async8(&sm.ph[(k + 1) & 1][threadIdx.x], &a.phasor[QPSK_CHUNK + tbase + QPSK_CHUNK + threadIdx.x]);
            }

            // matched filter: rrc_fir.c:22-28
            u64 acc[R];
#ifdef QPSK_FRONT_PROF
            pt2 = clock64(); pc_fill += pt2 - pt; pt = pt2;
            PROF_TRACE(k, w, 0, pt2);
            if (k == 64) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g64));
            if (k == 65) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g65));
#endif
            fir_strip<NTAPS, R, MODE>(xcur - (NTAPS - 1), tb.t, acc);
#ifdef QPSK_FRONT_PROF
            pt2 = clock64(); pc_strip += pt2 - pt; pt = pt2;
            PROF_TRACE(k, w, 1, pt2);
#endif
            // sm.out is a single tile: wait until the timing warps have taken the previous one
            if (k > 0) bar_sync(BAR_EMPTY, QPSK_FIR_THREADS + QPSK_AUX_THREADS);
#ifdef QPSK_FRONT_PROF
            pt2 = clock64(); pc_empty += pt2 - pt; pt = pt2;
#endif
            // raw sums go to shared memory; the output gain (a double multiply behind two conversions on the
            // narrow XU pipe) is applied by the timing warps, off the filter's critical path
            u64* orow = &sm.out[lane][strip];
#pragma unroll
            for (int r = 0; r < R; r++) orow[r] = acc[r];
            bar_arrive(BAR_FULL, QPSK_FIR_THREADS + QPSK_AUX_THREADS);
            cp_async_wait_all();                       // next tile's PCM and phasors have landed
            bar_sync(BAR_FIR_DONE, QPSK_FIR_THREADS);
#ifdef QPSK_FRONT_PROF
            pt2 = clock64(); pc_post += pt2 - pt;
#endif
        }
#ifdef QPSK_FRONT_PROF
        if (lane == 0) {
            prof_row[8 + w] = pc_strip; prof_row[16 + w] = pc_post; prof_row[24 + w] = pc_fill; prof_row[32 + w] = pc_empty;
            if (w == 0) { prof_row[6] = clock64() - pstart; prof_row[7] = g64; prof_row[40] = g65; }
        }
#endif
        if (w == 0) PROF_ROLE_END(2);
        if (w == 7) PROF_ROLE_END(3);
    } else if (w < 10) {
        // ============================ timing + decimation warps ============================
        // warp 8 = I, warp 9 = Q, lane = channel: amplitude histograms of qpsk.c:131-167, one tile behind the filter
        const int comp = w - 8;
        const int nsym = N / SPS;
        // this CTA's frame scratch: [symbol][component][lane][SPS] floats, so that a symbol's SPS samples of one lane are one
        // (or two) 128-bit stores and a warp writes whole 512-byte runs
        if (threadIdx.x == 8 * 32) {
            int slot = -1;
            if (QPSK_SCRATCH_SLOT_CLAIM && a.scratch_slots != nullptr) {
                unsigned smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                if ((int)smid * QPSK_SCRATCH_SLOTS + QPSK_SCRATCH_SLOTS <= a.scratch_nslots) {
                    for (int tries = 0; tries < 4 && slot < 0; tries++)
                        for (int i = 0; i < QPSK_SCRATCH_SLOTS; i++)
                            if (atomicCAS(&a.scratch_slots[smid * QPSK_SCRATCH_SLOTS + i], 0, 1) == 0) { slot = (int)smid * QPSK_SCRATCH_SLOTS + i; break; }
                }
            }
            sm.scr_slot = slot;
        }
        bar_sync(BAR_AUX, QPSK_AUX_THREADS);
        const int scr_slot = sm.scr_slot;
        const u64 keep_policy = l2_evict_last_policy();
        float* scr = a.scratch + (size_t)(scr_slot >= 0 ? scr_slot : a.scratch_nslots + (int)blockIdx.x) * (size_t)QPSK_SCRATCH_REGION_FLOATS;
        auto scr_at = [&](int n, int c) -> const float* {         // sample n of the frame, component c, this lane
            return scr + ((size_t)((n / SPS) * 2 + c) * QPSK_GROUP + lane) * SPS + (n % SPS);
        };
        const int ntiles = nframes * tiles_per_frame;
        for (int fr = 0; fr < nframes; fr++) {
            const int f = f0 + fr;
            float av = 0.0f, mx = 0.0f;
            u64 hist = 0ull;
            // extension (ESTIMATE_TIMING): this component's half of S = sum_n y_n^2 e^{-2 pi i n / SPS}, the symbol-rate line
            // of the squared matched-filter output (Oerder & Meyr); its argument is the sampling phase of the eye
            const bool est = a.timing_t != nullptr || a.ub_mode == QPSK_UB_TAU;
            float sre = 0.0f, sim = 0.0f;
            for (int t = 0; t < tiles_per_frame; t++) {
                bar_sync(BAR_FULL, QPSK_FIR_THREADS + QPSK_AUX_THREADS);
#ifdef QPSK_FRONT_PROF
                PROF_TRACE(fr * tiles_per_frame + t, w, 0, clock64());
#endif
                const float* ow = reinterpret_cast<const float*>(&sm.out[lane][0]) + comp;
#pragma unroll 2
                for (int s = 0; s < TILE_SYMS; s++) {
                    float ys[SPS];
#pragma unroll
                    for (int j = 0; j < SPS; j++) {
                        const int ti = s * SPS + j;
                        const float y = gain_exact(ow[2 * ti]);                  // rrc_fir.c:28, this warp's component
                        ys[j] = y;
                        av = __fadd_rn(av, fabsf(y));
                        if (est) timing_accumulate<SPS>(j, __fmul_rn(y, y), sre, sim);
                    }
                    {                                                          // kept for the decimation
                        float4* dst = reinterpret_cast<float4*>(scr + ((size_t)((t * TILE_SYMS + s) * 2 + comp) * QPSK_GROUP + lane) * SPS);
#pragma unroll
                        for (int q = 0; q < SPS / 4; q++) st_keep_f4(dst + q, make_float4(ys[4 * q], ys[4 * q + 1], ys[4 * q + 2], ys[4 * q + 3]), keep_policy);
                    }
                    av = __fmul_rn(av, 1.0f / SPS);              // av /= CYCLES, exact for a power of two
                    mx = fmaxf(mx, av);                            // qpsk.c:140-146 (strict > or >= give the same maximum)
                    const float hv = __fmul_rn(mx, 0.125f);      // max / 8.0f
                    // first k in 1..7 with av <= hv*k, 0 if none (qpsk.c:154-166).  The rounded products hv*k are monotone
                    // in k (hv >= 0), so three compares of a bisection find it; only the three edges looked at are formed.
                    const bool p4 = av <= __fmul_rn(hv, 4.0f);
                    const bool p26 = av <= __fmul_rn(hv, p4 ? 2.0f : 6.0f);
                    const int lo = p4 ? (p26 ? 1 : 3) : (p26 ? 5 : 7);          // the odd edge that decides between lo and lo + 1
                    const bool p1 = av <= __fmul_rn(hv, (float)lo);
                    const int bin = (lo + (p1 ? 0 : 1)) & 7;                    // 7 + 1 -> 0: no edge reached
                    hist += 1ull << (8 * bin);                     // byte 0 collects "no bin"; counts <= 128 fit a byte
                }
#ifdef QPSK_FRONT_PROF
                PROF_TRACE(fr * tiles_per_frame + t, w, 1, clock64());
#endif
                if (fr * tiles_per_frame + t + 1 < ntiles) bar_arrive(BAR_EMPTY, QPSK_FIR_THREADS + QPSK_AUX_THREADS);
            }
            sm.hist[comp][lane] = hist;
            if (est) sm.tsum[comp][lane] = make_float2(sre, sim);
            __threadfence();                                       // both components of the frame are in the scratch
            bar_sync(BAR_AUX, QPSK_AUX_THREADS);
            int index = 0;
            {                                                      // qpsk.c:173-180 first strict maximum
                const u64 hi = sm.hist[0][lane], hq = sm.hist[1][lane];
                int hmax = 0;
#pragma unroll
                for (int kk = 1; kk < 8; kk++) {
                    const int h = (int)((hi >> (8 * kk)) & 0xff) + (int)((hq >> (8 * kk)) & 0xff);
                    if (h > hmax) { hmax = h; index = kk; }
                }
            }
            if (est) {
                const float2 ti = sm.tsum[0][lane], tq = sm.tsum[1][lane];
                const float2 S = make_float2(__fadd_rn(ti.x, tq.x), __fadd_rn(ti.y, tq.y));
                if (a.timing_t != nullptr && comp == 0 && live) a.timing_t[(size_t)f * a.Cpad + ch] = S;
                if (a.ub_mode == QPSK_UB_TAU) index = tau_index<SPS>(S.x, S.y);       // both warps of the pair decide alike
            }
            if (comp == 0 && live) a.index_t[(size_t)f * a.Cpad + ch] = index;
            if (a.fir_dbg != nullptr && live) {                    // parity tap: the whole filtered frame
                float2* dst = a.fir_dbg + (size_t)ch * ((size_t)a.F * N) + (size_t)f * N;
                for (int i = comp; i < N; i += 2) dst[i] = make_float2(__ldcg(scr_at(i, 0)), __ldcg(scr_at(i, 1)));
            }
            // decimate, qpsk.c:186-191: symbol i = sample i*SPS + index, stored channel-fastest
            {
                const int slot = (a.slot_base + 1 + f) % a.nslots;
                float2* dst = a.dec_ring + (size_t)slot * nsym * a.Cpad + ch;
                const int first = a.ub_mode == QPSK_UB_PHASE ? index % SPS : index;   // extension: a sampling phase, never a slip
                // The scratch reads come from L2 (~650 cycles each) and the filter warps wait on this warp for their next
                // hand-off, so they are issued in batches of 16 symbols (32 loads in flight) before anything is stored.
                constexpr int BATCH = 16;
                static_assert((NSYM / 2) % BATCH == 0, "decimation batches");
                for (int i0 = comp; i0 < NSYM; i0 += 2 * BATCH) {
                    float2 v[BATCH];
#pragma unroll
                    for (int b = 0; b < BATCH; b++) {
                        int j = (i0 + 2 * b) * SPS + first;
                        v[b] = make_float2(0.0f, 0.0f);            // aliasing read of decimated_frame[j-N]: patched by the Costas stage
                        if (j >= N && a.ub_mode == QPSK_UB_CLAMP) j = N - 1;
                        if (j < N) v[b] = make_float2(__ldcg(scr_at(j, 0)), __ldcg(scr_at(j, 1)));
                    }
#pragma unroll
                    for (int b = 0; b < BATCH; b++)
                        if (live) st_keep_f2(dst + (size_t)(i0 + 2 * b) * a.Cpad, v[b], keep_policy);
                }
            }
            if (a.fuse_costas) {
                __threadfence();                                   // ring + index writes before the flag
                bar_sync(BAR_AUX, QPSK_AUX_THREADS);
                if (threadIdx.x == 8 * 32) sm.frames_decimated = fr + 1;
            } else {
                if (a.block_progress != nullptr) __threadfence();  // ring + index writes before the progress word
                bar_sync(BAR_AUX, QPSK_AUX_THREADS);               // sm.hist is rewritten next frame
                if (a.block_progress != nullptr && threadIdx.x == 8 * 32)
                    *reinterpret_cast<volatile unsigned long long*>(a.block_progress + blockIdx.x) = a.progress_ticket | (unsigned)(fr + 1);
            }
        }
        // both timing warps are done with the scratch: the slot goes back to the SM
        bar_sync(BAR_AUX, QPSK_AUX_THREADS);
        if (threadIdx.x == 8 * 32 && scr_slot >= 0) {
            __threadfence();
            atomicExch(&a.scratch_slots[scr_slot], 0);
        }
        if (w == 8) PROF_ROLE_END(4);
    } else {
        // ================================== Costas warp ==================================
        if (!a.fuse_costas || !live) return;
        const CostasParams p = costas_params(a.costas);
        float2 st;
        if (a.fuse_costas == 2 && fb > 0) {
            const unsigned amask = __activemask();                     // the live lanes; lane 0 is always one of them
            volatile unsigned long long* relay = a.block_progress + a.relay_group_base + g;
            // The loop state comes from the CTA that runs the previous frame block of these channels.  That CTA has a lower block
            // index (blocks are numbered frame block by frame block), so it was dispatched earlier -- a whole wave earlier when a
            // frame block is a wave of CTAs, and then this wait is over before it begins; with fewer channel groups the blocks of
            // a group sit on the machine together and the loop warps take turns while the filter warps of all of them run.  A CTA
            // never waits for a later one, so the chain always ends at block 0.  Same four-second watchdog as the chasing loop.
            int gave_up = 0;
            if (lane == 0) {
                const unsigned long long want = a.progress_ticket | (unsigned)f0;
                unsigned long long t0 = 0;
                for (int spins = 0; *relay < want; spins++) {
                    __nanosleep(QPSK_COSTAS_POLL_NS);
                    if ((spins & 1023) == 1023) {
                        unsigned long long t;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                        if (t0 == 0) t0 = t;
                        else if (t - t0 > 4000000000ull) { gave_up = 1; *a.relay_watchdog = a.relay_ticket; break; }
                    }
                }
                __threadfence();
            }
            gave_up = __shfl_sync(amask, gave_up, 0);
            if (gave_up) {                                             // let the later blocks of the group through: the call is lost anyway
                if (lane == 0) *relay = a.progress_ticket | (unsigned)f1;
                return;
            }
            st = __ldcg(&a.costas.loop_state[ch]);                     // past L1: an earlier CTA of this SM may have read the line
        } else {
            st = a.costas.loop_state[ch];
        }
        float phase = st.x, freq = st.y;
        for (int fr = 0; fr < nframes; fr++) {
            // call f consumes the frame decimated one call earlier (already in the ring) and patches the
            // frame produced by this call, so it must wait until that one has been decimated
            // a frame takes ~70 us to arrive; every poll is four issue slots taken from the filter warps (profiles/r02_notes.md:
            // the kernel is bound by the schedulers' issue ports), so the wait sleeps QPSK_COSTAS_POLL_NS between looks
            while (sm.frames_decimated < fr + 1) __nanosleep(QPSK_COSTAS_POLL_NS);
            __threadfence();
            costas_run_frame<true, 1>(a.costas, p, f0 + fr, ch, phase, freq);
            __syncwarp(__activemask());                            // every lane has read the slot
            costas_discard_slot(a.costas, f0 + fr, ch - lane, lane);
        }
        a.costas.loop_state[ch] = make_float2(phase, freq);
        if (a.fuse_costas == 2) {                                      // hand the loop over to the next frame block of these channels
            __threadfence();                                           // (nothing of this lives in a register across the loop: the kernel sits at its 80)
            __syncwarp(__activemask());
            if (lane == 0) {
                const int ng = (a.chan_count + QPSK_GROUP - 1) / QPSK_GROUP;
                const int gg = blockIdx.x % ng, ff1 = min(a.F, (blockIdx.x / ng + 1) * a.frames_per_block);
                *reinterpret_cast<volatile unsigned long long*>(a.block_progress + a.relay_group_base + gg) = a.progress_ticket | (unsigned)ff1;
            }
        }
        PROF_ROLE_END(5);
    }
}
