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
    int slot_base, nslots, ub_mode;
    float alpha, beta, max_freq, min_freq;
    float2 rot45;            // cmplx(ROTATE45) from the host libm, qpsk.c:75
};

__global__ void __launch_bounds__(128) costas_kernel(const CostasArgs a) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.C) return;
    float2 st = a.loop_state[c];
    float phase = st.x, freq = st.y;
    const size_t slot_elems = (size_t)a.nsym * a.Cpad;
    const int words = a.nsym / 16;

    for (int f = 0; f < a.F; f++) {
        const float2* cur = a.dec_ring + (size_t)((a.slot_base + f) % a.nslots) * slot_elems + c;
        // Out-of-frame read of qpsk.c:190 (sps 4, index >= 4): in the Makefile build
        // input_frame[512+k] is decimated_frame[k], which at that point already holds symbol k of
        // the frame consumed now.  Patch the last symbol of the frame produced by this call.
        if (a.ub_mode == 0) {
            const int idx = a.index_t[(size_t)f * a.Cpad + c];
            const int j = (a.nsym - 1) * a.sps + idx - a.N;
            if (j >= 0) {
                float2* nxt = a.dec_ring + (size_t)((a.slot_base + f + 1) % a.nslots) * slot_elems + c;
                nxt[(size_t)(a.nsym - 1) * a.Cpad] = cur[(size_t)j * a.Cpad];
            }
        }
        for (int wd = 0; wd < words; wd++) {
            unsigned bits = 0u;
#pragma unroll 4
            for (int k = 0; k < 16; k++) {
                const int i = wd * 16 + k;
                const float2 d = cur[(size_t)i * a.Cpad];
                float s, co;
                sincosf_glibc(phase, s, co);                       // cmplxconj(get_phase()), qpsk.h:36
                const float2 y = cmul_exact(d, make_float2(co, -s)); // qpsk.c:197
                if (a.costas_dbg != nullptr)
                    a.costas_dbg[((size_t)f * a.nsym + i) * a.Cpad + c] = y;
                // costas_loop.c:44-47: sign(I)*Q - sign(Q)*I with 0 -> -1
                const float e = __fsub_rn((y.x > 0.0f ? y.y : -y.y), (y.y > 0.0f ? y.x : -y.x));
                freq = __fadd_rn(freq, __fmul_rn(a.beta, e));       // costas_loop.c:57
                phase = __fadd_rn(__fadd_rn(phase, freq), __fmul_rn(a.alpha, e));   // :58
                // costas_loop.c:61-67: compares and subtracts in double against TAU
                while ((double)phase > 6.283185307179586) phase = __double2float_rn(__dsub_rn((double)phase, 6.283185307179586));
                while ((double)phase < -6.283185307179586) phase = __double2float_rn(__dadd_rn((double)phase, 6.283185307179586));
                if (freq > a.max_freq) freq = a.max_freq;           // :69-74
                else if (freq < a.min_freq) freq = a.min_freq;
                // qpsk.c:74-79
                const float2 r = cmul_exact(y, a.rot45);
                bits |= ((r.x < 0.0f ? 1u : 0u) | (r.y < 0.0f ? 2u : 0u)) << (2 * k);
            }
            a.dibits_t[((size_t)f * words + wd) * a.Cpad + c] = bits;
        }
        a.track_t[(size_t)f * a.Cpad + c] = make_float2(phase, freq);
    }
    a.loop_state[c] = make_float2(phase, freq);
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
