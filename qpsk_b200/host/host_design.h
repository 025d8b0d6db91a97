/* host_design.h -- host-side (C11) design-time arithmetic of the B200 receiver: RRC taps,
 * Costas loop gains and the one-off rotators.  These run once per context on the CPU with the
 * host libm, exactly as the reference computes them (rrc_fir.c:32-76, costas_loop.c:31-54,
 * qpsk.h:35-36), and are uploaded as kernel constants.  Compiled with -ffp-contract=off. */
#ifndef QPSK_HOST_DESIGN_H
#define QPSK_HOST_DESIGN_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    float phase, freq, max_freq, min_freq, damping, loop_bw, alpha, beta;
} qpsk_host_loop;

/* rrc_fir.c:32-76 with NTAPS a parameter; taps[] receives ntaps floats */
void qpsk_host_rrc_make(float *taps, int ntaps, float fs, float rs, float alpha);
/* costas_loop.c:31-42 */
void qpsk_host_loop_create(qpsk_host_loop *l, float loop_bw, float min_freq, float max_freq);
void qpsk_host_loop_update_gains(qpsk_host_loop *l);
/* qpsk.h:35 cmplx(v) / qpsk.h:36 cmplxconj(v): out = {cosf(v), +-sinf(v)} with v rounded to float first */
void qpsk_host_cis(double v, int conjugate, float out[2]);

#ifdef __cplusplus
}
#endif
#endif
