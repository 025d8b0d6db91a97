/*
 * dropin_demo.c -- a C program written only against include/qpsk_dropin.h (the reference's API):
 * repeats the reference's loop-back experiment (qpsk.c:289-359) with the RNG-independent bit
 * pattern of SURVEY.md Appendix B and dumps every observable to a binary file that
 * tests/test_dropin_gpu.py compares with the reference golden vectors.
 */
#include <complex.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qpsk_dropin.h"

#define TAU (2.0 * 3.14159265358979323846)

static void put(FILE *f, const void *p, size_t n) { if (fwrite(p, 1, n, f) != n) { perror("fwrite"); exit(2); } }

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s out.bin\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "wb");
    if (!f) { perror(argv[1]); return 2; }

    create_control_loop((TAU / 100.0f), -1.0f, 1.0f);       /* qpsk.c:302 */
    rrc_make(9600.0, 2400.0, .35f);                         /* qpsk.c:308 */
    qpsk_dropin_tx_reset(2400.0, 1550.0);                   /* qpsk.c:316-321 */
    qpsk_dropin_rx_reset(2400.0, 1500.0);                   /* qpsk.c:341-342 */

    /* transmit 8 packets of 256 symbols */
    static int16_t pcm[8 * 1024];
    for (unsigned p = 0; p < 8; p++) {
        int bits[512];
        for (unsigned i = 0; i < 512; i++) { unsigned k = p * 512 + i; bits[i] = (int)(((k * k + k / 3) >> 1) & 1); }
        int n = qpsk_packet_mod(&pcm[p * 1024], bits, 256);
        if (n != 1024) { fprintf(stderr, "qpsk_packet_mod returned %d\n", n); return 1; }
    }
    put(f, pcm, sizeof pcm);

    /* receive 16 frames */
    for (int k = 0; k < 16; k++) {
        rx_frame(&pcm[k * 512]);
        put(f, qpsk_dropin_costas_frame(), 128 * sizeof(complex float));
        float pf[3] = { get_phase(), get_frequency(), qpsk_dropin_offset_freq() };
        put(f, pf, sizeof pf);
        put(f, qpsk_dropin_rx_bits(), 256 * sizeof(int));
    }

    /* rrc_fir with a caller-owned delay line: impulse response */
    static complex float mem[NTAPS], x[200];
    x[0] = 1.0f;
    rrc_fir(mem, x, 200);
    put(f, x, sizeof x);
    put(f, mem, sizeof mem);

    /* fftn of the ramp 1..8 and ifft(fft(x)) at NFFT */
    complex double in8[8], out8[8];
    for (int i = 0; i < 8; i++) in8[i] = i + 1;
    fftn(in8, out8, 8);
    put(f, out8, sizeof out8);
    static complex double a[NFFT], b[NFFT], c[NFFT];
    for (int i = 0; i < NFFT; i++) a[i] = (double)((i * 37) % 101) / 50.0 - 1.0 + ((double)((i * 11) % 17) / 8.0 - 1.0) * I;
    fft(a, b);
    ifft(b, c);
    put(f, a, sizeof a); put(f, b, sizeof b); put(f, c, sizeof c);

    /* bit stages */
    uint16_t crc = crc16((const uint8_t *)"123456789", 9);
    put(f, &crc, sizeof crc);
    uint8_t dbg[8] = { 0xAA, 0xAA, 0xAA, 0xAA, 0, 0, 0, 0 };             /* interleave.c:105 */
    interleave(dbg, 8, INTERLEAVE);
    put(f, dbg, sizeof dbg);
    interleave(dbg, 8, DEINTERLEAVE);
    put(f, dbg, sizeof dbg);
    scramble_init(both);
    uint8_t ks[32];
    for (int i = 0; i < 32; i++) { ks[i] = 0; scramble(&ks[i], tx); }
    put(f, ks, sizeof ks);
    uint8_t d = 3;
    int rc = scramble(&d, both);                                          /* -1, leaves d untouched */
    put(f, &rc, sizeof rc);
    put(f, &d, 1);

    /* tx_frame with explicit complex symbols == qpsk_packet_mod on the same dibits */
    qpsk_dropin_tx_reset(2400.0, 1550.0);
    complex float sym[256];
    for (unsigned i = 0; i < 256; i++) {
        unsigned k0 = 2 * i, k1 = 2 * i + 1;
        int dibit[2] = { (int)(((k1 * k1 + k1 / 3) >> 1) & 1), (int)(((k0 * k0 + k0 / 3) >> 1) & 1) };
        sym[i] = qpsk_mod(dibit);
    }
    static int16_t pcm2[1024];
    if (tx_frame(pcm2, sym, 256) != 1024) return 1;
    put(f, pcm2, sizeof pcm2);
    int bits2[2];
    qpsk_demod(1.0f + 0.0f * I, bits2);
    put(f, bits2, sizeof bits2);
    fclose(f);
    return 0;
}
