// rx_front2.cuh -- K1 of the batched receiver, second generation: the same arithmetic as rx_front_kernel (mixer + RRC
// matched filter + timing histogram + decimation, qpsk.c:114-191 and rrc_fir.c:17-30, bit-exact in QPSK_MODE_EXACT)
// with the filter warps taken off every CTA-wide barrier.
//
// Why: the per-tile trace of rx_front_kernel (profiles/r02_notes.md, "front-end timeline") shows that the SM's warp
// schedulers serve the OLDER of the two co-resident CTAs first.  That CTA runs almost as if it were alone, and since all
// eight of its filter warps leave the FP32 pipe together once per 128-sample tile (mix the next tile, three barriers,
// hand the sums over), the pipe idles for ~3 k of every ~22 k cycles; the younger CTA only fills half of those gaps,
// because its own barrier-coupled warps cannot make progress in 3 k-cycle windows.  FMA pipe 83 % active.
//
// Here a CTA has no tile phases at all:
//   * the mixed samples live in a ring of 24 chunks of 16 samples (three tiles) per channel, written by ONE producer
//     warp that runs a strip-time ahead of the filter (PCM straight from HBM into registers, phasors from L1/L2);
//   * the filter warps take work items (one 16-sample strip = one chunk of outputs) from a shared counter, in time
//     order, so a warp that shares its scheduler with the loop warp simply takes fewer strips -- no static split to
//     balance -- and go from strip to strip without meeting anybody: the only things they ever wait for are
//     monotonic counters that are normally already past (chunks mixed, frames decimated);
//   * raw sums go to a per-CTA ring of 48 chunks (six tiles) in L2 (the former frame scratch), flagged per item; the
//     timing warps follow item by item, apply the output gain on the fly, and at the end of a frame decimate by
//     re-reading the ring (the gain is one rounded double multiply, so re-applying it gives the same bits);
//   * the loop warp is unchanged.
// The younger CTA on the SM now only has to cover the prologue and the tail of the older one.
#pragma once

#include "rx_front.cuh"

#ifndef QPSK_F2_FILTER_WARPS
#define QPSK_F2_FILTER_WARPS 4
#endif
#define QPSK_F2_THREADS (128 + 32 * QPSK_F2_FILTER_WARPS)   // 1 producer + 2 timing/decimation + 1 Costas + the filter warps (one per scheduler)
// poll intervals of the waits (ns)
#ifndef F2_NS_MIXED
#define F2_NS_MIXED 100
#endif
#ifndef F2_NS_RAW
#define F2_NS_RAW 400
#endif
#ifndef F2_NS_PRODUCER
#define F2_NS_PRODUCER 1000
#endif
#ifndef F2_NS_ITEM
#define F2_NS_ITEM 1000
#endif
#ifndef F2_NS_COSTAS
#define F2_NS_COSTAS 4000
#endif
#define QPSK_F2_RING_CHUNKS 24       // mixed-sample ring: three tiles
#define QPSK_F2_RAW_CHUNKS 48        // raw-sum ring in L2: six tiles = QPSK_SCRATCH_REGION_FLOATS
#define QPSK_F2_FLAGS 64             // > QPSK_F2_RAW_CHUNKS: an item's flag slot is not reused while the item can still be awaited

struct RxFront2Smem {
    static constexpr int XS = QPSK_F2_RING_CHUNKS * 16 + 3;   // [384], [385] mirror [0], [1]; odd row stride => conflict-free 64-bit access
    u64 x[QPSK_GROUP][XS];
    u64 hist[2][QPSK_GROUP];
    float2 tsum[2][QPSK_GROUP];
    int next_item;                    // work counter of the filter warps (chunk index of the next strip)
    volatile int mixed;               // chunks 0 .. mixed-1 are in the ring (producer)
    volatile int cur[QPSK_F2_FILTER_WARPS];   // the chunk each filter warp is working on: everything below the minimum is finished
    volatile int done[QPSK_F2_FLAGS]; // done[it % 64] == it + 1: the raw sums of item it are in the L2 ring
    volatile int frames_decimated;    // frames whose symbols are in the symbol ring, i.e. whose raw chunks are free again
    int scr_slot;
#ifdef F2_PAD_SMEM
    char pad[F2_PAD_SMEM];
#endif
};

__device__ __forceinline__ void st_keep_u64(u64* dst, const u64 v, u64 policy) {
    if (!QPSK_L2_STORE_HINT) { *dst = v; return; }
    asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(dst), "l"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ uint4 ldg_stream16(const void* p, u64 policy) {
    uint4 v;
    if (!QPSK_L2_PCM_HINT) { v = __ldg(reinterpret_cast<const uint4*>(p)); return v; }
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(policy));
    return v;
}
// waits poll: every poll is a handful of issue slots taken from warps that have work, so they sleep between polls
__device__ __forceinline__ void spin_until_ge(const volatile int* p, int v, unsigned ns) {
    while (*p < v) __nanosleep(ns);
}

// The taps of rx_front2_kernel: 127 taps + a zero 128th (so that the tap walk is eight equal trips) + (1, 1)
struct TapBank2 {
    float2 t[128];
    float2 one;
};

// One strip out of the sample ring, as a walk over the TAPS: acc[r] += tap[k] * x[r + k] for k = 0 .. 126, oldest tap first
// -- the order of rrc_fir.c:22-26 -- with the 16 samples x[k .. k + 15] that one tap meets held in registers and one new
// sample fetched per tap.  Unlike the walk over the samples (fir_strip) this has no triangular head and tail: the whole
// strip is ONE rolled body of 16 taps x 16 outputs (8 KB of code) executed eight times.  That matters here because the
// filter warps of this kernel are not in lock-step: warps at eight different places of a 16 KB head / loop / tail
// sequence, plus the auxiliary warps, missed the instruction cache so often that "no instruction" was the largest stall
// reason of the first build (profiles/r02_notes.md).  The 128th tap is zero: x * 0 = +-0 and acc + (+-0) = acc, because a
// sum that starts at +0 is never -0.  The two ring samples behind the window that the last trip also fetches are stale
// but finite (the ring is zeroed at start), so the zero tap annihilates them.
// Ring coordinates: item q's window starts at chunk q - 8, offset 2; trip t fetches x[16 + 16 t + j] = chunk q - 7 + t,
// offset 2 + j, j = 0 .. 15 -- offsets 16 and 17 are the next chunk's first two samples, which the ring mirrors at
// [384], [385] for its last chunk, so a trip's fetches are contiguous and the wrap is a pointer reset between trips.
template <int MODE>
__device__ __forceinline__ void fir_strip_taps(const u64* __restrict__ xrow, const int q, const float2* __restrict__ taps2, u64 (&acc)[16]) {
    constexpr int R = 16, TRIPS = 8;
    const u64 one = *reinterpret_cast<const u64*>(&taps2[128]);
    const u64* const wrap = xrow + QPSK_F2_RING_CHUNKS * 16 + 2;
    const u64* xp = xrow + ((q - 8) % QPSK_F2_RING_CHUNKS) * 16 + 2;
    u64 X[R];
#pragma unroll
    for (int r = 0; r < R; r++) { acc[r] = 0ull; X[r] = xp[r]; }
#pragma unroll 1
    for (int t = 0; t < TRIPS; t++) {
        xp += 16;
        if (xp == wrap) xp = xrow + 2;
        const int k0 = t * R;
#pragma unroll
        for (int j = 0; j < R; j++) {
#pragma unroll
            for (int r = 0; r < R; r++) fir_tap<MODE>(acc[r], X[(r + j) & (R - 1)], taps2, k0 + j, one);
            X[j] = xp[j];                            // x[k0 + 16 + j]: first needed by output 15 of the next tap
        }
    }
}

template <int NTAPS, int SPS, int MODE>
__global__ void __launch_bounds__(QPSK_F2_THREADS, 2) rx_front2_kernel(const __grid_constant__ RxFrontArgs a, const __grid_constant__ TapBank2 tb) {
    static_assert(NTAPS == 127, "the ring walk assumes 126 samples of history");
    constexpr int NSYM = 512 / SPS, CHUNK_SYMS = 16 / SPS, CPF = 512 / 16;     // chunks per frame
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RxFront2Smem& sm = *reinterpret_cast<RxFront2Smem*>(smem_raw);

    // warp roles: warps 0-3 = producer, timing I, timing Q, Costas (one per scheduler), warps 4-7 = filter (one per scheduler).
    // The schedulers share issue slots about evenly among the warps that are ready, and a filter warp is always ready: with
    // one filter warp per scheduler and CTA (two with the co-resident CTA -- enough to keep the FP32 pipe full, tools/
    // packed_peak_bench.cu) the latency-bound auxiliary warps get every third slot instead of every fifth.
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int w = wid >= 4 ? wid - 4 : wid + 8;     // role index: 0-3 filter, 8 producer, 9-10 timing, 11 Costas
    const int ngroups = (a.chan_count + QPSK_GROUP - 1) / QPSK_GROUP;
    const int g = blockIdx.x % ngroups, fb = blockIdx.x / ngroups;
    const int f0 = fb * a.frames_per_block;
    const int f1 = min(a.F, f0 + a.frames_per_block);
    if (f0 >= f1) return;
    const int ch = a.chan_base + g * QPSK_GROUP + lane;
    const int ch_end = min(a.C, a.chan_base + a.chan_count);
    const bool live = ch < ch_end;
    const int chl = live ? ch : ch_end - 1;         // padded lanes recompute the last channel, stores are masked
    const int N = a.N;                              // 512
    const int nframes = f1 - f0;
    const int nitems = nframes * CPF;               // strips of this CTA; item it <-> chunk q = it + 8 (chunks 0..7 are the history)
    const int Q = nitems + 8;

    for (int i = threadIdx.x; i < QPSK_GROUP * RxFront2Smem::XS; i += QPSK_F2_THREADS) (&sm.x[0][0])[i] = 0ull;   // finite everywhere, see fir_strip_taps
    if (threadIdx.x < QPSK_F2_FLAGS) sm.done[threadIdx.x] = 0;
    if (threadIdx.x < QPSK_F2_FILTER_WARPS) sm.cur[threadIdx.x] = 8;
    if (threadIdx.x == 0) {
        sm.next_item = 8; sm.mixed = 0; sm.frames_decimated = 0;
        // this CTA's region of the raw ring: one of its SM's slots (rewritten by whoever is resident, so it lives in L2), or the
        // private region of its block index if the slots are taken
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
    __syncthreads();
    const int scr_slot = sm.scr_slot;
    u64* const raw = reinterpret_cast<u64*>(a.scratch + (size_t)(scr_slot >= 0 ? scr_slot : a.scratch_nslots + (int)blockIdx.x) * (size_t)QPSK_SCRATCH_REGION_FLOATS);
    // raw[((it % 48) * 16 + s) * 32 + lane] = (I, Q) sums of sample s of item it

    if (w < QPSK_F2_FILTER_WARPS) {
        // =================================== filter warps ===================================
        const u64* xrow = &sm.x[lane][0];
        const u64 keep_policy = l2_evict_last_policy();
#ifdef QPSK_FRONT_PROF
        __shared__ int f2_trace_slot;
        if (w == 0 && lane == 0) {
            unsigned smid_; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));
            int ts = -1;
            if (smid_ == 0 || smid_ == 77) { ts = atomicAdd(&g_front_trace_count, 1) % 64; unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_));
                g_front_trace_hdr[ts][0] = g_; g_front_trace_hdr[ts][1] = smid_; g_front_trace_hdr[ts][2] = blockIdx.x; g_front_trace_hdr[ts][3] = clock64(); }
            f2_trace_slot = ts;
        }
        bar_sync(6, 32 * QPSK_F2_FILTER_WARPS);
        const int trace_slot = f2_trace_slot;
        long long* f2_trace = reinterpret_cast<long long*>(g_front_trace) + (size_t)(trace_slot < 0 ? 0 : trace_slot) * 2048 * 4;
        long long pc_mixed = 0, pc_strip = 0, pc_rawwait = 0, pc_store = 0, pc_items = 0, pt, pt2;
        unsigned long long* prof_row = g_front_prof[blockIdx.x % QPSK_FRONT_PROF_ROWS];
        const long long pstart = clock64();
        if (w == 0 && lane == 0) { unsigned smid_; unsigned long long g_; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_)); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); prof_row[0] = g_; prof_row[1] = smid_; }
#define F2PROF(var) do { pt2 = clock64(); var += pt2 - pt; pt = pt2; } while (0)
#else
#define F2PROF(var) do { } while (0)
#endif
        int pending = -1;                                         // item whose sums are stored but not yet flagged
        for (;;) {
            int q = 0;
            if (lane == 0) q = atomicAdd(&sm.next_item, 1);
            q = __shfl_sync(0xffffffffu, q, 0);
            if (q >= Q) break;
            if (lane == 0) sm.cur[w] = q;
#ifdef QPSK_FRONT_PROF
            pt = clock64(); pc_items++;
            const long long tr_start = pt;
#endif
            spin_until_ge(&sm.mixed, q + 1, F2_NS_MIXED);                      // chunks q - 8 .. q are in the ring (normally long since)
            asm volatile("" ::: "memory");                         // flag and samples are both shared memory: loads of one warp stay in order
            u64 acc[16];
            F2PROF(pc_mixed);
            fir_strip_taps<MODE>(xrow, q, tb.t, acc);      // rrc_fir.c:22-28
            F2PROF(pc_strip);
            const int it = q - 8;
            // the ring slot of this item was last used by item it - 48: its frame must have been decimated
            if (it >= QPSK_F2_RAW_CHUNKS) spin_until_ge(&sm.frames_decimated, (it - QPSK_F2_RAW_CHUNKS) / CPF + 1, F2_NS_RAW);
            F2PROF(pc_rawwait);
            // the previous item of this warp is flagged now, a strip after its stores were issued: the release fence (CTA scope --
            // the readers are warps of this CTA) finds them completed long ago and costs nothing
            if (pending >= 0) {
                __threadfence_block();
                __syncwarp();
                if (lane == 0) sm.done[pending & (QPSK_F2_FLAGS - 1)] = pending + 1;
            }
            u64* dst = raw + (size_t)((it % QPSK_F2_RAW_CHUNKS) * 16) * QPSK_GROUP + lane;
#pragma unroll
            for (int r = 0; r < 16; r++) st_keep_u64(dst + r * QPSK_GROUP, acc[r], keep_policy);
            pending = it;
            F2PROF(pc_store);
#ifdef QPSK_FRONT_PROF
            if (trace_slot >= 0 && lane == 0 && it < 2048) { long long* r = f2_trace + (size_t)it * 4; r[0] = tr_start; r[1] = pt - tr_start; r[2] = w; r[3] = 0; }
#endif
        }
        if (pending >= 0) {
            __threadfence_block();
            __syncwarp();
            if (lane == 0) sm.done[pending & (QPSK_F2_FLAGS - 1)] = pending + 1;
        }
        if (lane == 0) sm.cur[w] = 0x7fffffff;
#ifdef QPSK_FRONT_PROF
        if (lane == 0) {
            prof_row[8 + w] = pc_strip; prof_row[16 + w] = pc_store; prof_row[24 + w] = pc_mixed; prof_row[32 + w] = pc_rawwait; prof_row[40 + w] = pc_items;
            if (w == 0) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); prof_row[2] = g_; prof_row[6] = clock64() - pstart; }
        }
#endif
    } else if (w == 8) {
        // =================================== producer warp ===================================
        // mixes chunk after chunk into the ring: qpsk.c:115-117.  lane = channel; PCM comes straight from HBM (32 bytes per
        // lane and chunk = one sector), a chunk ahead; the 16 phasors of a chunk are the same for every lane.
        const int16_t* pcm_row = a.pcm + (size_t)(chl - a.chan_base) * a.pcm_row;
        const u64 pcm_policy = l2_evict_first_policy();
        u64* xrow = &sm.x[lane][0];
        auto chunk_src = [&](int q) -> const int16_t* {
            if (q >= 8) return pcm_row + (size_t)f0 * N + (size_t)(q - 8) * 16;
            return (f0 == 0) ? a.pcm_tail + (size_t)chl * QPSK_CHUNK + q * 16 : pcm_row + (size_t)f0 * N - QPSK_CHUNK + q * 16;
        };
        uint4 n0 = ldg_stream16(chunk_src(0), pcm_policy), n1 = ldg_stream16(chunk_src(0) + 8, pcm_policy);
        const float4* ph4 = reinterpret_cast<const float4*>(a.phasor + (size_t)f0 * N);     // table index of chunk q, sample e: f0 N + 16 q + e
#ifdef QPSK_FRONT_PROF
        long long pw_wait = 0; const long long pw_start = clock64();
#endif
        for (int q = 0; q < Q; q++) {
            const uint4 p0 = n0, p1 = n1;
            if (q + 1 < Q) {
                const int16_t* src = chunk_src(q + 1);
                n0 = ldg_stream16(src, pcm_policy);
                n1 = ldg_stream16(src + 8, pcm_policy);
            }
            float2 phr[16];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const float4 v = __ldg(ph4 + (size_t)q * 8 + e);
                phr[2 * e] = make_float2(v.x, v.y);
                phr[2 * e + 1] = make_float2(v.z, v.w);
            }
            // chunk q overwrites chunk q - 24, whose last reader is item q - 16
#ifdef QPSK_FRONT_PROF
            const long long pw0 = clock64();
#endif
            if (q >= QPSK_F2_RING_CHUNKS) {
                for (;;) {
                    int l = lane < QPSK_F2_FILTER_WARPS ? sm.cur[lane] : 0x7fffffff;
                    l = __reduce_min_sync(0xffffffffu, l);
                    if (l > q - 16) break;
                    __nanosleep(F2_NS_PRODUCER);
                }
            }
#ifdef QPSK_FRONT_PROF
            pw_wait += clock64() - pw0;
#endif
            const int c = q % QPSK_F2_RING_CHUNKS;
#ifndef F2_SKIP_MIX
            mix_store(xrow + c * 16, p0, p1, phr);
#endif
            if (c == 0) { xrow[QPSK_F2_RING_CHUNKS * 16] = xrow[0]; xrow[QPSK_F2_RING_CHUNKS * 16 + 1] = xrow[1]; }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) sm.mixed = q + 1;
        }
#ifdef QPSK_FRONT_PROF
        if (lane == 0) { unsigned long long* prof_row = g_front_prof[blockIdx.x % QPSK_FRONT_PROF_ROWS]; prof_row[3] = pw_wait; prof_row[7] = clock64() - pw_start; }
#endif
    } else if (w < 11) {
        // ============================ timing + decimation warps ============================
        // warp 9 = I, warp 10 = Q, lane = channel: amplitude histograms of qpsk.c:131-167, an item behind the filter
        const int comp = w - 9;
        const int nsym = N / SPS;
        const u64 keep_policy = l2_evict_last_policy();
        const float* rawf = reinterpret_cast<const float*>(raw) + comp;
        auto wait_item = [&](int it) { while (sm.done[it & (QPSK_F2_FLAGS - 1)] != it + 1) __nanosleep(F2_NS_ITEM); __threadfence_block(); };
        auto load_item = [&](int it, float (&v)[16]) {
            const float* p = rawf + ((size_t)((it % QPSK_F2_RAW_CHUNKS) * 16) * QPSK_GROUP + lane) * 2;
#pragma unroll
            for (int s = 0; s < 16; s++) v[s] = __ldcg(p + (size_t)s * QPSK_GROUP * 2);
        };
        auto raw_at = [&](int fr, int n) -> const float2* {       // raw sums of sample n of frame fr, this lane
            const int it = fr * CPF + (n >> 4);
            return reinterpret_cast<const float2*>(raw + (size_t)((it % QPSK_F2_RAW_CHUNKS) * 16 + (n & 15)) * QPSK_GROUP + lane);
        };
#ifdef QPSK_FRONT_PROF
        long long tw_wait = 0, tw_dec = 0; int tw_pre = 0; const long long tw_start = clock64();
#define PROF_TW_WAIT(stmt) do { const long long tw0_ = clock64(); stmt; tw_wait += clock64() - tw0_; } while (0)
#else
#define PROF_TW_WAIT(stmt) do { stmt; } while (0)
#endif
        float va[16], vb[16];
        for (int fr = 0; fr < nframes; fr++) {
            const int f = f0 + fr;
            float av = 0.0f, mx = 0.0f;
            u64 hist = 0ull;
            const bool est = a.timing_t != nullptr || a.ub_mode == QPSK_UB_TAU;
            float sre = 0.0f, sim = 0.0f;
            auto work = [&](const float (&v)[16]) {               // one item: 16 samples = CHUNK_SYMS symbols
#pragma unroll
                for (int s = 0; s < CHUNK_SYMS; s++) {
#pragma unroll
                    for (int j = 0; j < SPS; j++) {
                        const float y = gain_exact(v[s * SPS + j]);                  // rrc_fir.c:28, this warp's component
                        av = __fadd_rn(av, fabsf(y));
                        if (est) timing_accumulate<SPS>(j, __fmul_rn(y, y), sre, sim);
                    }
                    av = __fmul_rn(av, 1.0f / SPS);              // av /= CYCLES, exact for a power of two
                    mx = fmaxf(mx, av);                            // qpsk.c:140-146 (strict > or >= give the same maximum)
                    const float hv = __fmul_rn(mx, 0.125f);      // max / 8.0f
                    // first k in 1..7 with av <= hv*k, 0 if none (qpsk.c:154-166): exact bisection over the monotone edges
                    const bool p4 = av <= __fmul_rn(hv, 4.0f);
                    const bool p26 = av <= __fmul_rn(hv, p4 ? 2.0f : 6.0f);
                    const int lo = p4 ? (p26 ? 1 : 3) : (p26 ? 5 : 7);
                    const bool p1 = av <= __fmul_rn(hv, (float)lo);
                    const int bin = (lo + (p1 ? 0 : 1)) & 7;                    // 7 + 1 -> 0: no edge reached
                    hist += 1ull << (8 * bin);                     // byte 0 collects "no bin"; counts <= 128 fit a byte
                }
            };
            // two items per trip, the sums of the next one requested before the current one is worked on (L2 latency)
            PROF_TW_WAIT(wait_item(fr * CPF); load_item(fr * CPF, va));
            for (int cc = 0; cc < CPF; cc += 2) {
                const int it = fr * CPF + cc;
                PROF_TW_WAIT(wait_item(it + 1); load_item(it + 1, vb));
#ifndef F2_SKIP_TIMING
                work(va);
#endif
                if (cc + 2 < CPF) { PROF_TW_WAIT(wait_item(it + 2); load_item(it + 2, va)); }
#ifndef F2_SKIP_TIMING
                work(vb);
#endif
            }
#ifdef QPSK_FRONT_PROF
            const long long td0 = clock64();
#endif
            sm.hist[comp][lane] = hist;
            if (est) sm.tsum[comp][lane] = make_float2(sre, sim);
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
                for (int i = comp; i < N; i += 2) {
                    const float2 r = __ldcg(raw_at(fr, i));
                    dst[i] = make_float2(gain_exact(r.x), gain_exact(r.y));
                }
            }
            // decimate, qpsk.c:186-191: symbol i = sample i*SPS + index, stored channel-fastest.  The gain is applied again to
            // the raw sums of the chosen samples (one rounded multiply: the same bits the histogram saw).
            {
                const int slot = (a.slot_base + 1 + f) % a.nslots;
                float2* dst = a.dec_ring + (size_t)slot * nsym * a.Cpad + ch;
                const int first = a.ub_mode == QPSK_UB_PHASE ? index % SPS : index;   // extension: a sampling phase, never a slip
                constexpr int BATCH = 16;
                static_assert((NSYM / 2) % BATCH == 0, "decimation batches");
                for (int i0 = comp; i0 < NSYM; i0 += 2 * BATCH) {
                    float2 r[BATCH];
                    bool in[BATCH];
#pragma unroll
                    for (int b = 0; b < BATCH; b++) {
                        int j = (i0 + 2 * b) * SPS + first;
                        if (j >= N && a.ub_mode == QPSK_UB_CLAMP) j = N - 1;
                        in[b] = j < N;                             // else: aliasing read of decimated_frame[j-N], patched by the Costas stage
                        r[b] = make_float2(0.0f, 0.0f);
                        if (in[b]) r[b] = __ldcg(raw_at(fr, j));
                    }
#pragma unroll
                    for (int b = 0; b < BATCH; b++) {
                        const float2 y = in[b] ? make_float2(gain_exact(r[b].x), gain_exact(r[b].y)) : make_float2(0.0f, 0.0f);
                        if (live) st_keep_f2(dst + (size_t)(i0 + 2 * b) * a.Cpad, y, keep_policy);
                    }
                }
            }
            __threadfence();                                       // ring + index writes before the flag
            bar_sync(BAR_AUX, QPSK_AUX_THREADS);                   // both warps are done with the frame's raw chunks
            if (w == 9 && lane == 0) {
                sm.frames_decimated = fr + 1;
                if (a.block_progress != nullptr) *reinterpret_cast<volatile unsigned long long*>(a.block_progress + blockIdx.x) = a.progress_ticket | (unsigned)(fr + 1);
            }
#ifdef QPSK_FRONT_PROF
            tw_dec += clock64() - td0;
#endif
        }
#ifdef QPSK_FRONT_PROF
        if (lane == 0 && w == 9) { unsigned long long* prof_row = g_front_prof[2048 + blockIdx.x % 2048]; unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_));
            prof_row[4] = g_; prof_row[0] = tw_wait; prof_row[1] = tw_dec; prof_row[2] = tw_pre; prof_row[3] = clock64() - tw_start; }
#endif
        // both timing warps are done with the raw ring (and so are the filter warps): the slot goes back to the SM
        if (w == 9 && lane == 0 && scr_slot >= 0) {
            __threadfence();
            atomicExch(&a.scratch_slots[scr_slot], 0);
        }
    } else {
        // ================================== Costas warp ==================================
        if (!a.fuse_costas || !live) return;
        const CostasParams p = costas_params(a.costas);
        const float2 st = a.costas.loop_state[ch];
        float phase = st.x, freq = st.y;
        for (int fr = 0; fr < nframes; fr++) {
            // call f consumes the frame decimated one call earlier (already in the ring) and patches the
            // frame produced by this call, so it must wait until that one has been decimated
            while (sm.frames_decimated < fr + 1) __nanosleep(F2_NS_COSTAS);
            __threadfence();
#ifndef F2_SKIP_COSTAS
            costas_run_frame<true, 1>(a.costas, p, f0 + fr, ch, phase, freq);
#endif
            __syncwarp(__activemask());                            // every lane has read the slot
            costas_discard_slot(a.costas, f0 + fr, ch - lane, lane);
        }
        a.costas.loop_state[ch] = make_float2(phase, freq);
    }
}
