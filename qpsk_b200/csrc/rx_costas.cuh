// rx_costas.cuh -- K3: per-channel Costas carrier recovery, slicer and Gray demapping.
//
// Replaces the symbol loop of rx_frame (reference qpsk.c:196-217) together with
// costas_loop.c:44-74 (phase_detector, advance_loop, phase_wrap, frequency_limit) and
// qpsk_demod (qpsk.c:74-79).  The loop is a strict recurrence over the symbols of one stream, so
// a stream maps to one thread with (phase, freq) in registers; channels are the parallel axis
// and every global access is channel-fastest (coalesced).
#pragma once

#include "common.cuh"

struct CostasArgs {
    float2* dec_ring;        // [nslots][nsym][Cpad]; call f of this launch consumes slot (slot_base + f) % nslots
    const int* index_t;      // [F][Cpad] timing index of the frame stored in slot (slot_base + 1 + f)
    float2* loop_state;      // [Cpad] (d_phase, d_freq), carried across launches
    unsigned* dibits_t;      // [F][nsym/16][Cpad] 16 dibits per word, symbol i at bits 2*(i%16)
    float2* costas_dbg;      // optional [F][nsym][Cpad] derotated symbols (costas_frame), may be null
    float2* track_t;         // [F][Cpad] (phase, freq) after each frame
    int C, Cpad, F, nsym, sps, N;
    int c0, c1;              // channels [c0, c1) are processed by the standalone kernel launch
    int slot_base, nslots, ub_mode;
    float2* est_bursts;      // optional [C][est_n]: the loop leaves the 4th power of the call's first est_n symbols here (the estimator's
    int est_n, est_f_off;    // input) as it consumes them; est_f_off = position of this launch's frame 0 in the call
    int discard_from;        // QPSK_B200_TRANSIENT_SYMBOLS: ring slots consumed by frame 0 and by frames >= discard_from are dropped from L2
                             // without write-back once the loop has read them (-1: every slot is kept, OUT_DEC stays readable)
    float alpha, beta, max_freq, min_freq;
    float2 rot45;            // cmplx(ROTATE45) from the host libm, qpsk.c:75
};

struct CostasParams {
    float alpha, beta, max_freq, min_freq;
    float2 rot45;
};

// one symbol of the loop: derotate, detect, update, wrap, clamp, slice.  Returns bits[0] | bits[1] << 1.
__device__ __forceinline__ unsigned costas_symbol(const float2 d, float& phase, float& freq, const CostasParams& p, float2& y) {
    float s, co;
    sincosf_glibc(phase, s, co);                               // cmplxconj(get_phase()), qpsk.h:36
    y = cmul_exact_packed(d, make_float2(co, -s));             // qpsk.c:197
    // costas_loop.c:44-47: sign(I)*Q - sign(Q)*I with 0 -> -1
    const float e = __fsub_rn((y.x > 0.0f ? y.y : -y.y), (y.y > 0.0f ? y.x : -y.x));
    freq = __fadd_rn(freq, __fmul_rn(p.beta, e));              // costas_loop.c:57
    phase = __fadd_rn(__fadd_rn(phase, freq), __fmul_rn(p.alpha, e));   // :58
    // costas_loop.c:61-67 compares and subtracts in double against TAU = 6.283185307179586.  The nearest floats
    // around TAU are 0x40C90FDA = 6.2831850051879883 (< TAU) and 0x40C90FDB = 6.2831854820251465 (> TAU), so
    // (double)phase > TAU  <=>  phase > 6.283185f as floats: the common case needs no conversion and no branch
    // on a double compare; the correction itself (rare, once per ~2*pi/freq symbols) stays in double.
    if (fabsf(phase) > 6.2831850051879883f) {
        while ((double)phase > 6.283185307179586) phase = __double2float_rn(__dsub_rn((double)phase, 6.283185307179586));
        while ((double)phase < -6.283185307179586) phase = __double2float_rn(__dadd_rn((double)phase, 6.283185307179586));
    }
    if (freq > p.max_freq) freq = p.max_freq;                  // :69-74
    else if (freq < p.min_freq) freq = p.min_freq;
    const float2 r = cmul_exact_packed(y, p.rot45);            // qpsk.c:74-79
    return (r.x < 0.0f ? 1u : 0u) | (r.y < 0.0f ? 2u : 0u);
}

// z^4 as the frequency estimator wants it (QPSK on the axes raised to the 4th power is a tone at 4 x offset); one definition
// for the loop's own emission and for symbol_power4_kernel, so both give the same bits
__device__ __forceinline__ float2 power4_exact(const float2 z) {
    const float2 z2 = make_float2(__fsub_rn(__fmul_rn(z.x, z.x), __fmul_rn(z.y, z.y)), __fmul_rn(__fmul_rn(2.0f, z.x), z.y));
    return make_float2(__fsub_rn(__fmul_rn(z2.x, z2.x), __fmul_rn(z2.y, z2.y)), __fmul_rn(__fmul_rn(2.0f, z2.x), z2.y));
}

// The ring slot frame f has just consumed holds frame m = (call frame of f) - 1 of this call; if it belongs to the estimator's
// burst, its 4th powers go to est_bursts[c][m * nsym ...] now, while the slot is hot in L2 -- instead of a pass over the ring
// after the kernel (0.31 ms and 0.5 GB of reads per 65,536-channel call).  Off the loop's path on purpose (a call, no inlining):
// it runs for 8 of 64 frames and must not cost the frame loop a register.
__device__ __noinline__ void costas_emit_power4(const float2* __restrict__ cur, float2* __restrict__ dst, int nsym, int Cpad) {
    for (int k = 0; k < nsym; k += 4) {
        float2 w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) w[i] = power4_exact(__ldcg(cur + (size_t)(k + i) * Cpad));
        reinterpret_cast<float4*>(dst + k)[0] = make_float4(w[0].x, w[0].y, w[1].x, w[1].y);
        reinterpret_cast<float4*>(dst + k)[1] = make_float4(w[2].x, w[2].y, w[3].x, w[3].y);
    }
}

// One rx_frame call's worth of loop iterations for channel c (qpsk.c:196-212): consumes ring slot
// (slot_base + f), after patching the last symbol of the frame this call produced (slot + 1).
// CG loads bypass L1
// for the fused kernel, where other warps of the same CTA have just written the ring.
template <bool CG, int AHEAD>
__device__ __forceinline__ void costas_run_frame(const CostasArgs& a, const CostasParams& p, int f, int c, float& phase, float& freq) {
    const size_t slot_elems = (size_t)a.nsym * a.Cpad;
    const float2* cur = a.dec_ring + (size_t)((a.slot_base + f) % a.nslots) * slot_elems + c;
    auto ld = [](const float2* q) -> float2 { return CG ? __ldcg(q) : *q; };
    // Out-of-frame read of qpsk.c:190 (sps 4, index >= 4): in the Makefile build input_frame[512+k] is
    // decimated_frame[k], which at that point already holds symbol k of the frame consumed now.
    if (a.ub_mode == 0) {
        const int idx = CG ? __ldcg(&a.index_t[(size_t)f * a.Cpad + c]) : a.index_t[(size_t)f * a.Cpad + c];
        const int j = (a.nsym - 1) * a.sps + idx - a.N;
        if (j >= 0) {
            float2* nxt = a.dec_ring + (size_t)((a.slot_base + f + 1) % a.nslots) * slot_elems + c;
            nxt[(size_t)(a.nsym - 1) * a.Cpad] = ld(cur + (size_t)j * a.Cpad);
        }
    }
    // Symbols are fetched AHEAD groups of four ahead of the recurrence.  The fused kernel uses one group (the
    // loop stays ~9 KB, so the Costas warp does not evict the filter loop from the instruction cache and the
    // ring is L2-warm); the stand-alone kernel uses four (one output word): with few channels there are few
    // warps per SM to hide the ~1 us DRAM latency of a cold ring.
    const int groups = a.nsym / 4;
    float2 d[AHEAD][4];
#pragma unroll
    for (int q = 0; q < AHEAD; q++)
#pragma unroll
        for (int k = 0; k < 4; k++) d[q][k] = ld(cur + (size_t)(q * 4 + k) * a.Cpad);
    unsigned bits = 0u;
#pragma unroll 1
    for (int g0 = 0; g0 < groups; g0 += AHEAD) {
#pragma unroll
        for (int q = 0; q < AHEAD; q++) {
            const int gq = g0 + q;
            float2 cur4[4];
#pragma unroll
            for (int k = 0; k < 4; k++) cur4[k] = d[q][k];
            if (gq + AHEAD < groups) {                         // refill this slot for AHEAD groups later
#pragma unroll
                for (int k = 0; k < 4; k++) d[q][k] = ld(cur + (size_t)((gq + AHEAD) * 4 + k) * a.Cpad);
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float2 y;
                bits |= costas_symbol(cur4[k], phase, freq, p, y) << (2 * ((gq & 3) * 4 + k));
                if (a.costas_dbg != nullptr) a.costas_dbg[((size_t)f * a.nsym + gq * 4 + k) * a.Cpad + c] = y;
            }
            if ((gq & 3) == 3) {                               // 16 dibits per output word
                a.dibits_t[((size_t)f * (a.nsym / 16) + (gq >> 2)) * a.Cpad + c] = bits;
                bits = 0u;
            }
        }
    }
    a.track_t[(size_t)f * a.Cpad + c] = make_float2(phase, freq);
    if (a.est_bursts != nullptr) {
        const int m = f + a.est_f_off - 1;
        if (m >= 0 && m * a.nsym < a.est_n) costas_emit_power4(cur, a.est_bursts + (size_t)c * a.est_n + (size_t)m * a.nsym, a.nsym, a.Cpad);
    }
}

// The ring slot frame f has just consumed is dead (nobody reads it again: the next reader of that slot is a later call's frame
// after it has been rewritten): drop its lines from L2 instead of letting them be written back to HBM.  One warp = the 32
// channels of a group = two 128-byte lines per symbol.
__device__ __forceinline__ void costas_discard_slot(const CostasArgs& a, int f, int group_first_channel, int lane) {
    if (a.discard_from < 0 || !(f == 0 || f >= a.discard_from)) return;
    const char* base = reinterpret_cast<const char*>(a.dec_ring + (size_t)((a.slot_base + f) % a.nslots) * ((size_t)a.nsym * a.Cpad) + group_first_channel);
    for (int i = lane; i < 2 * a.nsym; i += 32) {
        const char* q = base + (size_t)(i >> 1) * a.Cpad * sizeof(float2) + (size_t)(i & 1) * 128;
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(q) : "memory");
    }
}

__device__ __forceinline__ CostasParams costas_params(const CostasArgs& a) {
    CostasParams p;
    p.alpha = a.alpha; p.beta = a.beta; p.max_freq = a.max_freq; p.min_freq = a.min_freq; p.rot45 = a.rot45;
    return p;
}

__global__ void __launch_bounds__(128) costas_kernel(const CostasArgs a) {
    const int c = a.c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.c1) return;
    const CostasParams p = costas_params(a);
    const float2 st = a.loop_state[c];
    float phase = st.x, freq = st.y;
    for (int f = 0; f < a.F; f++) costas_run_frame<false, 4>(a, p, f, c, phase, freq);
    a.loop_state[c] = make_float2(phase, freq);
}

// The loop of a call whose front end runs as frame chunks, as ONE kernel that chases the chunks: chunk k's frames are
// processed as soon as flags[k] carries this call's ticket (chunk_signal_kernel, enqueued behind chunk k's front end on the
// front end's stream).  Saves the launch and the cold start of one loop kernel per chunk -- for 1,024 streams the loop is a
// pure dependency chain and every chunk boundary was ~35 us of it.  The host launches this kernel behind chunk 0's front end;
// should a later launch of the call fail, it writes the ticket into every flag itself (rx_run_call's bail), so the kernel
// cannot outlive the call.  Symbols and indices are read past L1 (they are written while this kernel runs).
__global__ void chunk_signal_kernel(int* flag, int ticket) {
    __threadfence();
    *reinterpret_cast<volatile int*>(flag) = ticket;
}
__global__ void __launch_bounds__(128) costas_chase_kernel(const CostasArgs a, int* __restrict__ flags, int ticket, int frames_per_chunk, int watchdog_slot) {
    const int c = a.c0 + blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = c < a.c1;
    const int cl = live ? c : a.c1 - 1;
    const CostasParams p = costas_params(a);
    const float2 st = a.loop_state[cl];
    float phase = st.x, freq = st.y;
    __shared__ int give_up;
    if (threadIdx.x == 0) give_up = 0;
    for (int f0 = 0, k = 0; f0 < a.F; f0 += frames_per_chunk, k++) {
        if (threadIdx.x == 0) {
            const volatile int* fl = flags + k;
            unsigned long long t0 = 0;
            // watchdog: a chunk that has not come after four seconds never will (a bug in the host's ordering); the kernel then
            // says so in flags[watchdog_slot] and leaves instead of hanging the device -- the host reports it at its next sync
            for (int spins = 0; *fl != ticket; spins++) {
                __nanosleep(500);
                if ((spins & 1023) == 1023) {
                    unsigned long long t;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                    if (t0 == 0) t0 = t;
                    else if (t - t0 > 4000000000ull) { give_up = 1; flags[watchdog_slot] = ticket; break; }
                }
            }
            __threadfence();
        }
        __syncthreads();
        if (give_up) return;
        const int f1 = min(a.F, f0 + frames_per_chunk);
        if (live)
            for (int f = f0; f < f1; f++) costas_run_frame<true, 4>(a, p, f, c, phase, freq);
    }
    if (live) a.loop_state[c] = make_float2(phase, freq);
}

// The loop beside a front end whose CTAs own frame BLOCKS (channel counts that are an awkward number of waves of whole-stream
// CTAs: 16,384 channels are 1.73 waves, as frame blocks they are 6.92): a persistent grid of one-warp CTAs, warp w following
// channel groups w, w + gridDim, ... frame by frame through the progress words the front-end CTAs publish
// (progress[fb * ngroups + g] = ticket << 32 | frames done).  Block order = the front end's launch order (all groups of frame
// block 0, then block 1, ...), and with gridDim = the number of resident front-end CTAs a warp's groups sit in different
// waves, so it is never more than a frame behind.  Loop state travels through loop_state between blocks.
__global__ void __launch_bounds__(32, 32) costas_follow_kernel(const CostasArgs a, const unsigned long long* __restrict__ progress, unsigned long long ticket_hi,
                                                           int ngroups, int frames_per_block, int* __restrict__ watchdog, int ticket) {
    const int lane = threadIdx.x;
    const CostasParams p = costas_params(a);
    const int nfb = (a.F + frames_per_block - 1) / frames_per_block;
    for (int fb = 0; fb < nfb; fb++) {
        const int f0 = fb * frames_per_block, f1 = min(a.F, f0 + frames_per_block);
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
            const int c = a.c0 + g * 32 + lane;
            const bool live = c < a.c1;
            float phase = 0.0f, freq = 0.0f;
            if (live) { const float2 st = a.loop_state[c]; phase = st.x; freq = st.y; }
            const volatile unsigned long long* pw = progress + (size_t)fb * ngroups + g;
            for (int f = f0; f < f1; f++) {
                int gave_up = 0;
                if (lane == 0) {
                    const unsigned long long want = ticket_hi | (unsigned)(f - f0 + 1);
                    unsigned long long t0 = 0;
                    for (int spins = 0; *pw < want; spins++) {
                        __nanosleep(1000);
                        if ((spins & 1023) == 1023) {                      // watchdog, see costas_chase_kernel
                            unsigned long long t;
                            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                            if (t0 == 0) t0 = t;
                            else if (t - t0 > 4000000000ull) { gave_up = 1; *watchdog = ticket; break; }
                        }
                    }
                    __threadfence();
                }
                gave_up = __shfl_sync(0xffffffffu, gave_up, 0);
                if (gave_up) return;
                if (live) costas_run_frame<true, 1>(a, p, f, c, phase, freq);
                __syncwarp();                                              // every lane has read the slot
                costas_discard_slot(a, f, c - lane, lane);
            }
            if (live) a.loop_state[c] = make_float2(phase, freq);
        }
    }
}

// [rows][Cpad] channel-fastest -> [C][rows] channel-major (download layout), element = T
template <typename T>
__global__ void transpose_to_channel_major(const T* __restrict__ src, T* __restrict__ dst, int rows, int C, int Cpad) {
    __shared__ T tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        if (r < rows && c < C) tile[j][threadIdx.x] = src[(size_t)r * Cpad + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < C) dst[(size_t)c * rows + r] = tile[threadIdx.x][j];
    }
}
