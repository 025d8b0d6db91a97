/*
 * ref_harness.c -- wraps the UNMODIFIED reference sources (found through -I$(REF_DIR), never
 * copied into this repository) so tests can call their file-static functions and read their
 * globals.  TEST INFRASTRUCTURE ONLY.  Output goes to oracle/_ref/ (git-ignored).
 *
 * One translation unit: the reference headers use #pragma once, so including them first and
 * re-defining RS / NTAPS afterwards re-parameterises rrc_fir.c / qpsk.c without editing them
 * (SURVEY.md Appendix C).  `main` is renamed so qpsk.c:289 does not become the entry point.
 *
 * The parity flavour MUST be built like the reference's Makefile:7 (-std=c11, no -O): at
 * 2400 baud rx_frame reads input_frame[512..515] (qpsk.c:190), which in that build is
 * decimated_frame[0..3]; ref_layout_ok() reports whether this build has that layout.
 */
#define _POSIX_C_SOURCE 199309L
#include <stdint.h>
#include <string.h>
#include <time.h>

#include "qpsk.h"
#include "rrc_fir.h"
#include "costas_loop.h"

#ifdef REF_RS
#undef RS
#define RS REF_RS
#endif
#ifdef REF_NTAPS
#undef NTAPS
#define NTAPS REF_NTAPS
#endif

#include "rrc_fir.c"
#include "costas_loop.c"
#define main qpsk_ref_main
#include "qpsk.c"
#undef main

int ref_sps(void) { return CYCLES; }
int ref_ntaps(void) { return NTAPS; }
int ref_frame_size(void) { return FRAME_SIZE; }

/* 1 when input_frame[] is immediately followed by decimated_frame[] (the Makefile layout) */
int ref_layout_ok(void) {
    return (uintptr_t)&input_frame[0] + sizeof input_frame == (uintptr_t)&decimated_frame[0];
}

void ref_get_taps(float *out) { memcpy(out, coeffs, sizeof coeffs); }

void ref_rx_reset(void) {
    memset(rx_filter, 0, sizeof rx_filter);
    memset(input_frame, 0, sizeof input_frame);
    memset(decimated_frame, 0, sizeof decimated_frame);
    memset(costas_frame, 0, sizeof costas_frame);
    create_control_loop((TAU / 100.0f), -1.0f, 1.0f);      /* qpsk.c:302 */
    fbb_rx_phase = cmplx(0.0f);                             /* qpsk.c:341 */
    fbb_rx_rect = cmplxconj(TAU * CENTER / FS);             /* qpsk.c:342 */
}

void ref_init(void) {
    rrc_make(FS, RS, .35f);                                 /* qpsk.c:308 */
    ref_rx_reset();
}

/* run rx_frame once and copy out every observable; any pointer may be NULL */
void ref_rx_frame(const int16_t *pcm, float *fir, float *dec_new, float *costas, uint8_t *dibit, float *phase_freq) {
    int16_t buf[FRAME_SIZE];
    memcpy(buf, pcm, sizeof buf);
    rx_frame(buf);
    const int nsym = FRAME_SIZE / CYCLES;
    if (fir) memcpy(fir, input_frame, sizeof input_frame);
    if (dec_new) memcpy(dec_new, &decimated_frame[nsym], (size_t)nsym * sizeof(complex float));
    if (costas) memcpy(costas, costas_frame, (size_t)nsym * sizeof(complex float));
    if (dibit) {
        for (int i = 0; i < nsym; i++) {
            int bits[2];
            qpsk_demod(costas_frame[i], bits);              /* the reference discards these, qpsk.c:209 */
            dibit[i] = (uint8_t)(bits[0] | (bits[1] << 1));
        }
    }
    if (phase_freq) { phase_freq[0] = get_phase(); phase_freq[1] = get_frequency(); }
}

/* batch: pcm[c][F*FRAME_SIZE]; each channel starts from ref_rx_reset() */
void ref_rx_run(const int16_t *pcm, int nchan, int nframes, float *fir, float *dec, float *costas,
                uint8_t *dibit, float *phase, float *freq) {
    const size_t N = FRAME_SIZE, S = FRAME_SIZE / CYCLES, F = (size_t)nframes;
    for (size_t c = 0; c < (size_t)nchan; c++) {
        ref_rx_reset();
        for (size_t f = 0; f < F; f++) {
            const size_t u = c * F + f;
            float pf[2];
            ref_rx_frame(pcm + u * N, fir ? fir + u * N * 2 : 0, dec ? dec + u * S * 2 : 0,
                         costas ? costas + u * S * 2 : 0, dibit ? dibit + u * S : 0, pf);
            if (phase) phase[u] = pf[0];
            if (freq) freq[u] = pf[1];
        }
    }
}

/* seconds of CPU time for `reps` passes of rx_frame over nframes frames (timing flavour) */
double ref_rx_time(const int16_t *pcm, int nframes, int reps) {
    struct timespec t0, t1;
    int16_t buf[FRAME_SIZE];
    ref_rx_reset();
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < reps; r++)
        for (int f = 0; f < nframes; f++) {
            memcpy(buf, pcm + (size_t)f * FRAME_SIZE, sizeof buf);
            rx_frame(buf);
        }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

void ref_rrc_fir(float *memory, float *sample, int length) {
    rrc_fir((complex float *)memory, (complex float *)sample, length);
}

float ref_phase_detector(float re, float im) { return phase_detector(re + im * I); }

void ref_qpsk_demod(float re, float im, int *bits) { qpsk_demod(re + im * I, bits); }

void ref_qpsk_mod(const int *bits, float *out) {
    int b[2] = { bits[0], bits[1] };
    complex float s = qpsk_mod(b);
    out[0] = crealf(s); out[1] = cimagf(s);
}

void ref_tx_reset(double carrier_hz) {
    memset(tx_filter, 0, sizeof tx_filter);
    fbb_tx_phase = cmplx(0.0f);                             /* qpsk.c:316 */
    fbb_tx_rect = cmplx(TAU * carrier_hz / FS);             /* qpsk.c:320 with (CENTER + 50.0) */
}

/* samples must hold length*CYCLES entries (main's own buffer is too small) */
int ref_packet_mod(int16_t *samples, int *tx_bits, int length) { return qpsk_packet_mod(samples, tx_bits, length); }

void ref_get_rx_consts(float *out) {
    out[0] = crealf(fbb_rx_rect); out[1] = cimagf(fbb_rx_rect);
    complex float r = cmplx(ROTATE45);
    out[2] = crealf(r); out[3] = cimagf(r);
}
