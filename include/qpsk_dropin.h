/*
 * qpsk_dropin.h -- the reference's own single-channel C API, exported by libqpsk_b200.so.
 *
 * Every prototype below is the one the reference declares (file:line cited); a program written
 * against the reference's headers links against this library unchanged -- e.g. the reference's
 * qpsk.c compiled with its own qpsk.h/rrc_fir.h/costas_loop.h and linked with -lqpsk_b200 instead
 * of rrc_fir.c and costas_loop.c (tests/test_dropin_gpu.py does exactly that).
 *
 * The array-processing entry points (rrc_fir, rx_frame, tx_frame, qpsk_packet_mod, fft*, crc16,
 * interleave) run on the GPU as a batch of one channel / frame / burst; they have no CPU path and
 * abort with a message on stderr when no B200 is usable (the reference's signatures return void).
 * The scalar control functions (Costas setters/getters and per-symbol helpers, qpsk_mod/qpsk_demod,
 * scramble) are plain host C: they are the API's bookkeeping, the batched kernels fuse the same
 * arithmetic where throughput matters.
 *
 * C only: the prototypes use C99 complex types, like the reference headers.
 */
#ifndef QPSK_DROPIN_H
#define QPSK_DROPIN_H

#include <complex.h>
#include <stdint.h>

/* ---- rrc_fir.h:13-17 ------------------------------------------------------------------------ */
#ifndef NTAPS
#define NTAPS 127
#endif
void rrc_fir(complex float memory[], complex float sample[], int length);
void rrc_make(float fs, float rs, float alpha);

/* ---- costas_loop.h:16-43 -------------------------------------------------------------------- */
void  create_control_loop(float loop_bw, float min_freq, float max_freq);
float phase_detector(complex float sample);
void  update_gains(void);
void  advance_loop(float error);
void  phase_wrap(void);
void  frequency_limit(void);
void  set_loop_bandwidth(float);
void  set_damping_factor(float);
void  set_alpha(float);
void  set_beta(float);
void  set_frequency(float);
void  set_phase(float);
void  set_max_freq(float);
void  set_min_freq(float);
float get_loop_bandwidth(void);
float get_damping_factor(void);
float get_alpha(void);
float get_beta(void);
float get_frequency(void);
float get_phase(void);
float get_max_freq(void);
float get_min_freq(void);

/* ---- qpsk.c:24-29 (file-static in the reference; public here under the same names) ------------ */
#ifndef FRAME_SIZE
#define FRAME_SIZE 512
#endif
void          rx_frame(int16_t in[FRAME_SIZE]);
int           tx_frame(int16_t samples[], complex float symbol[], int length);
complex float qpsk_mod(int bits[2]);
void          qpsk_demod(complex float symbol, int bits[2]);
int           qpsk_packet_mod(int16_t samples[], int tx_bits[], int length);
/* what rx_frame leaves in the reference's globals (qpsk.c:41,51) */
const complex float *qpsk_dropin_costas_frame(void);   /* costas_frame[FRAME_SIZE / CYCLES] of the last call */
const int           *qpsk_dropin_rx_bits(void);        /* bits[0], bits[1] of every symbol of the last call (the reference discards them) */
float                qpsk_dropin_offset_freq(void);    /* fbb_offset_freq, qpsk.c:217 */
/* the state main() sets up by hand (qpsk.c:316-321, 341-342); rs selects 2400 or 1200 baud */
void qpsk_dropin_rx_reset(double rs, double center_hz);
void qpsk_dropin_tx_reset(double rs, double carrier_hz);

/* ---- algorithms/fft.h:44-49 ----------------------------------------------------------------- */
#ifndef NFFT
#define NFFT 512
#endif
void fft(complex double *in, complex double *out);
void fftn(complex double *in, complex double *out, int n);
void ifft(complex double *in, complex double *out);
void ifftn(complex double *in, complex double *out, int n);

/* ---- algorithms/crc16.h:10, interleave.h:10-13, bit-scramble.h:21-30 -------------------------- */
uint16_t crc16(const uint8_t *data, int length);
#define INTERLEAVE   0
#define DEINTERLEAVE 1
void interleave(uint8_t *inout, int nbytes, int dir);
typedef enum { tx, rx, both } SRegister;
void scramble_init(SRegister sr);
int  scramble(uint8_t *dibit, SRegister sr);

#endif
