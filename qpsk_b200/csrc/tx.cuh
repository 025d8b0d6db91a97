// tx.cuh -- K5: batched transmit path, qpsk_packet_mod + tx_frame of the reference
// (qpsk.c:58-63 constellation, :225-264 tx_frame, :269-285) for many channels.
//
// Bit-exact: the zero-stuffed RRC filter is evaluated in polyphase form (only the taps that meet
// a symbol); the terms skipped are +-0 products, and adding +-0 never changes a sum that started
// at +0 (it can never be -0), so the result equals rrc_fir.c:22-26 over the stuffed signal.  The
// up-mixing phasor is a per-channel recurrence (each channel has its own carrier), run by one
// warp with lane = channel while the other warps filter.
#pragma once

#include "common.cuh"
#include "rx_front.cuh"   // TapBank

struct TxArgs {
    const uint8_t* symbols;   // [C][nsym] constellation index per symbol: (tx_bits[2k] << 1) | tx_bits[2k+1]
    const float2* symbols_cf; // or, when non-null, arbitrary complex symbols [C][nsym] (tx_frame's own argument)
    int16_t* pcm;             // [C][nsym*SPS]
    float2* phase_state;      // [Cpad] fbb_tx_phase
    const float2* rect;       // [Cpad] fbb_tx_rect = cmplx(TAU * carrier / FS)
    float2* sym_hist;         // [Cpad][128/SPS] the symbols of the previous 128 samples (zeros at stream start)
    int C, Cpad, nsym;
    int packet_samples;       // fbb_tx_phase is renormalised after every packet (qpsk.c:253)
    int sample_pos;           // samples already sent in the current packet
};

template <int SPS>
struct TxSmem {
    static constexpr int TS = QPSK_CHUNK / SPS;          // symbols per 128-sample tile
    u64 sym[QPSK_GROUP][2 * TS + 1];                     // [0,TS) previous tile, [TS,2TS) current tile
    float2 ph[QPSK_GROUP][QPSK_CHUNK + 1];               // up-mix phasor per sample of the tile
    short out[QPSK_GROUP][QPSK_CHUNK + 8];               // staged PCM tile
};

template <int NTAPS, int SPS>
__global__ void __launch_bounds__(256, 3) tx_kernel(const __grid_constant__ TxArgs a, const __grid_constant__ TapBank<NTAPS> tb) {
    constexpr int R = 16, TS = QPSK_CHUNK / SPS, SPT = TS / 8;   // SPT symbols loaded per thread per tile
    static_assert(TS % 8 == 0 && (NTAPS - 1) <= QPSK_CHUNK, "tile geometry");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TxSmem<SPS>& sm = *reinterpret_cast<TxSmem<SPS>*>(smem_raw);

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ch = blockIdx.x * QPSK_GROUP + lane;
    const bool live = ch < a.C;
    const int chl = live ? ch : a.C - 1;
    const int strip = w * R;
    u64* srow = &sm.sym[lane][0];
    const uint8_t* symrow = a.symbols + (size_t)chl * a.nsym;

    // qpsk.c:58-63, Gray coded: index 0 -> +1, 1 -> +j, 2 -> -j, 3 -> -1
    auto point = [](unsigned idx) -> u64 {
        const float re = (idx == 0) ? 1.0f : (idx == 3 ? -1.0f : 0.0f);
        const float im = (idx == 1) ? 1.0f : (idx == 2 ? -1.0f : 0.0f);
        return pack2(re, im);
    };

    // history -> "current" slots; the loop's shift moves it into place
    for (int i = w; i < TS; i += 8) srow[TS + i] = reinterpret_cast<const u64*>(a.sym_hist)[(size_t)chl * TS + i];
    float2 phase = make_float2(1.f, 0.f), rect = make_float2(1.f, 0.f);
    if (w == 0) { phase = a.phase_state[chl]; rect = a.rect[chl]; }
    int pos = a.sample_pos;
    __syncthreads();

    // any length (qpsk.c:225-264 takes any): the last tile may be partial.  Its missing symbols are zeros that only meet
    // outputs which are never stored, the phasor stops at the call's last sample, and the history keeps the last TS real symbols.
    const int ntiles = (a.nsym + TS - 1) / TS;
    int valid = TS;                                      // real symbols of the current tile
    for (int k = 0; k < ntiles; k++) {
        valid = (a.nsym - k * TS < TS) ? a.nsym - k * TS : TS;
        // shift own symbol slots and load this tile's symbols
#pragma unroll
        for (int e = 0; e < SPT; e++) {
            const int i = w * SPT + e;
            srow[i] = srow[TS + i];
            if (i >= valid) srow[TS + i] = 0ull;
            else if (a.symbols_cf != nullptr) srow[TS + i] = reinterpret_cast<const u64*>(a.symbols_cf)[(size_t)chl * a.nsym + (size_t)k * TS + i];
            else srow[TS + i] = point(symrow[(size_t)k * TS + i] & 3u);
        }
        __syncthreads();

        // warp 0: the up-mix phasor of every sample of the tile, qpsk.c:248-253 (lane = channel)
        if (w == 0) {
            const int nsamp = valid * SPS;
            for (int t = 0; t < nsamp; t++) {
                phase = cmul_exact(phase, rect);
                sm.ph[lane][t] = phase;
                if (++pos == a.packet_samples) {     // end of a tx_frame call: normalise (qpsk.c:253)
                    const double dr = (double)phase.x, di = (double)phase.y;
                    const float mag = __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(di, di))));
                    phase.x = __fdiv_rn(phase.x, mag);
                    phase.y = __fdiv_rn(phase.y, mag);
                    pos = 0;
                }
            }
        }

        // pulse shaping: outputs n = 128k + strip + r.  Symbol slot q (sample 4*(q - TS) relative to the tile
        // start) meets output r through tap i = SPS*(q - TS) - (strip + r) + NTAPS-1.
        u64 acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = 0ull;
        // slots that can reach this strip: SPS*(q-TS) in [strip - (NTAPS-1), strip + R - 1]
#pragma unroll
        for (int dq = -((NTAPS - 1) / SPS) - 1; dq <= (R - 1) / SPS; dq++) {
            // q = TS + strip/SPS + dq; strip is a multiple of 16, hence of SPS
            const u64 sv = srow[TS + strip / SPS + dq];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = SPS * dq - r + (NTAPS - 1);
                if (i >= 0 && i < NTAPS) {
                    const u64 cc = *reinterpret_cast<const u64*>(&tb.t[i]);
                    acc[r] = add2(acc[r], mul2_exact(sv, cc));
                }
            }
        }
        __syncthreads();    // phasors ready

#pragma unroll
        for (int r = 0; r < R; r++) {
            float yr, yi;
            unpack2(acc[r], yr, yi);
            yr = gain_exact(yr);                                       // rrc_fir.c:28
            yi = gain_exact(yi);
            const float2 p = sm.ph[lane][strip + r];                   // (stale beyond a partial tile's end: those outputs are not stored)
            const float re = __fsub_rn(__fmul_rn(yr, p.x), __fmul_rn(yi, p.y));   // crealf(signal[i] * fbb_tx_phase), qpsk.c:250
            sm.out[lane][strip + r] = (short)__float2int_rz(__fmul_rn(re, 16384.0f));   // qpsk.c:260 truncation
        }
        __syncthreads();

        // coalesced store: each warp writes 4 channel rows, 8 bytes per lane
        for (int rr = w; rr < QPSK_GROUP; rr += 8) {
            const int c = blockIdx.x * QPSK_GROUP + rr;
            if (c < a.C && lane * 4 < valid * SPS) {                  // valid * SPS is a multiple of 4: whole 8-byte groups
                const uint2 v = *reinterpret_cast<const uint2*>(&sm.out[rr][lane * 4]);
                *reinterpret_cast<uint2*>(a.pcm + (size_t)c * a.nsym * SPS + (size_t)k * QPSK_CHUNK + lane * 4) = v;
            }
        }
        // the next iteration's shift touches only symbol slots; sm.out / sm.ph are rewritten after its first barrier
    }
    __syncthreads();
    if (live) {
        // the last TS real symbols: slots [valid, valid + TS) of (previous tile | current tile)
        for (int i = w; i < TS; i += 8) reinterpret_cast<u64*>(a.sym_hist)[(size_t)ch * TS + i] = srow[valid + i];
        if (w == 0) a.phase_state[ch] = phase;
    }
}
