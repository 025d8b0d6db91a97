/*
 * qpsk_oracle.h -- CPU restatement of the MonsieurETM/QPSK receive/transmit hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker.  The product path is qpsk_b200/csrc (CUDA) behind include/.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against the unmodified
 * reference compiled from /root/reference (oracle/_ref, recipe in oracle/Makefile) by
 * tests/test_oracle_vs_ref.py, against the one golden vector the reference ships
 * (interleave.c:97-103) and against the known answers of SURVEY.md Appendix B.  Exceptions,
 * which have no reference implementation and are therefore "parity unpinned":
 * orc_fft_argmax (|X|^2 first-strict-maximum) and the orc_frame_* bit-pipeline composition.
 *
 * All arithmetic is written component-wise on {re,im} floats so that every rounding step is
 * explicit; build with -ffp-contract=off and no -march so that no FMA
 * is ever emitted (see oracle/Makefile).
 */
#ifndef QPSK_ORACLE_H
#define QPSK_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_TAPS   512
#define ORC_MAX_FRAME  512
#define ORC_GAIN       1.85 /* rrc_fir.h:14 (a double literal) */

typedef struct { float re, im; } orc_cf;
typedef struct { double re, im; } orc_cd;

enum { ORC_UB_ALIAS = 0, /* Makefile (-O0) build: input_frame[512+k] aliases decimated_frame[k] */
       ORC_UB_CLAMP = 1, /* fenced variant: out-of-frame reads return input_frame[FRAME_SIZE-1] */
       ORC_UB_PHASE = 2, /* extension: sample i*CYCLES + index % CYCLES, never out of frame */
       ORC_UB_TAU = 3    /* extension: index = round(tau) mod CYCLES from the square-law timing sum (orc_tau_index) */ };

/* ---- RRC FIR: rrc_fir.c:17-76 -------------------------------------------------------- */
void orc_rrc_make(float *taps, int ntaps, float fs, float rs, float alpha);
void orc_rrc_fir(const float *taps, int ntaps, orc_cf *memory, orc_cf *sample, int length);

/* ---- Costas control loop: costas_loop.c:13-154 --------------------------------------- */
typedef struct {
    float phase, freq, max_freq, min_freq, damping, loop_bw, alpha, beta;
} orc_loop;
void  orc_loop_create(orc_loop *l, float loop_bw, float min_freq, float max_freq);
float orc_phase_detector(orc_cf s);
void  orc_loop_update_gains(orc_loop *l);
void  orc_loop_advance(orc_loop *l, float error);
void  orc_loop_phase_wrap(orc_loop *l);
void  orc_loop_frequency_limit(orc_loop *l);
void  orc_loop_set_frequency(orc_loop *l, float f);
void  orc_loop_set_phase(orc_loop *l, float p);

/* ---- modem profile + per-channel state: qpsk.c:33-53, qpsk.h:16-23 -------------------- */
typedef struct {
    int   sps;          /* CYCLES = (int)(FS/RS) */
    int   frame_size;   /* FRAME_SIZE */
    int   nsym;         /* FRAME_SIZE / CYCLES */
    int   ntaps;        /* NTAPS */
    int   ub_mode;      /* ORC_UB_* */
    float fs, rs, center;
    float taps[ORC_MAX_TAPS];
    orc_cf rx_rect;     /* cmplxconj(TAU*CENTER/FS), qpsk.c:342 */
    orc_cf rot45;       /* cmplx(ROTATE45), qpsk.c:75 */
    orc_loop loop0;     /* loop constants + initial phase/freq, qpsk.c:302 */
} orc_profile;

typedef struct {
    orc_cf rx_phase;                 /* fbb_rx_phase */
    orc_cf fir_mem[ORC_MAX_TAPS];    /* rx_filter[] */
    orc_cf dec[ORC_MAX_FRAME];       /* decimated_frame[] (2*nsym used) */
    float  phase, freq;              /* d_phase, d_freq */
} orc_rx_state;

typedef struct {           /* optional per-frame taps of every intermediate; NULL = skip */
    orc_cf  *fir;          /* [frame_size]  input_frame after rrc_fir, qpsk.c:125 */
    int32_t *index;        /* [1]           timing "index", qpsk.c:173-180 */
    orc_cf  *dec;          /* [nsym]        newly decimated symbols, qpsk.c:190 */
    orc_cf  *costas;       /* [nsym]        costas_frame, qpsk.c:197 */
    uint8_t *dibit;        /* [nsym]        bits[0] | bits[1]<<1 from qpsk_demod, qpsk.c:209 */
    float   *phase, *freq; /* [1]           loop state after the frame */
    float   *offset_hz;    /* [1]           fbb_offset_freq, qpsk.c:217 */
} orc_rx_trace;

void orc_profile_init(orc_profile *p, float fs, float rs, float center, float rrc_alpha,
                      int ntaps, int frame_size, float loop_bw, int ub_mode);
void orc_rx_state_init(const orc_profile *p, orc_rx_state *s);
void orc_rx_frame(const orc_profile *p, orc_rx_state *s, const int16_t *pcm, const orc_rx_trace *t);

/* batch driver: C channels x F frames, pcm[c][F*frame_size]; any output may be NULL.
 * fir[c][F*frame], index[c][F], dec[c][F*nsym], costas[c][F*nsym], dibit[c][F*nsym],
 * phase[c][F], freq[c][F].  states[c] carried in/out. */
void orc_rx_run(const orc_profile *p, orc_rx_state *states, const int16_t *pcm, int nchan, int nframes,
                orc_cf *fir, int32_t *index, orc_cf *dec, orc_cf *costas, uint8_t *dibit,
                float *phase, float *freq);

void orc_qpsk_demod(const orc_profile *p, orc_cf sym, int bits[2]);
/* extension: the generator's counter-based test noise, in place on one channel's PCM (csrc/channel.cuh) */
void orc_awgn(int16_t *pcm, long long nsamples, float sigma, uint64_t seed, long long first_sample, int channel);
/* extension, parity unpinned: square-law timing statistic of one filtered frame (see the .c file) */
void orc_timing_sum(const orc_cf *frame, int n, int sps, orc_cf *out);
int  orc_tau_index(orc_cf S, int sps);
void orc_profile_slice_diagonal(orc_profile *p, int on);

/* ---- transmit: qpsk.c:58-63,225-285 --------------------------------------------------- */
typedef struct {
    orc_cf tx_phase;                /* fbb_tx_phase */
    orc_cf tx_rect;                 /* fbb_tx_rect */
    orc_cf fir_mem[ORC_MAX_TAPS];   /* tx_filter[] */
} orc_tx_state;
void   orc_tx_state_init(orc_tx_state *s, float carrier_hz, float fs);
void   orc_tx_set_carrier(orc_tx_state *s, float carrier_hz, float fs);   /* qpsk.c:320 again, phase kept */
orc_cf orc_qpsk_mod(const int bits[2]);
int    orc_tx_frame(const orc_profile *p, orc_tx_state *s, int16_t *samples, const orc_cf *symbol, int length);
int    orc_qpsk_packet_mod(const orc_profile *p, orc_tx_state *s, int16_t *samples, const int *tx_bits, int length);

/* ---- FFT: algorithms/fft.c:38-136 (forward scaled by 1/n, inverse unscaled) ------------ */
void orc_fftn(const orc_cd *in, orc_cd *out, int n);
void orc_ifftn(const orc_cd *in, orc_cd *out, int n);
/* parity-unpinned estimator: first strict maximum of |X[k]|^2 over k = 0..n-1 */
int  orc_fft_argmax(const orc_cd *x, int n, double *mag2);

/* ---- bit stages: algorithms/{bit-scramble,interleave,crc16}.c -------------------------- */
#define ORC_SCRAMBLE_SEED 0x4A80
void     orc_scramble_dibit(uint16_t *reg, uint8_t *dibit);
void     orc_interleave(uint8_t *inout, int nbytes, int dir /* 0 = interleave, 1 = de-interleave */);
int      orc_interleave_prime(int nbytes);
uint16_t orc_crc16(const uint8_t *data, int length);

/* ---- frame format (parity unpinned composition of the pinned bit stages; see qpsk_oracle.c) */
void orc_frame_encode(const uint8_t *payload, int nbytes, uint8_t *dibits);
int  orc_frame_decode(const uint8_t *dibits, int nbytes, uint8_t *frame);
/* rotation resolved on the CRC: returns the quarter turns undone (0..3) or -1; see the .c file */
int  orc_frame_decode_rotated(const uint8_t *dibits, int nbytes, uint8_t *frame);
uint8_t orc_rotate_dibit(uint8_t d, int quarter_turns);

/* ---- glibc 2.39 sinf/cosf (x86-64 FMA ifunc variant), restated: the device NCO follows this */
float orc_glibc_sinf(float y);
float orc_glibc_cosf(float y);
long  orc_glibc_check(uint32_t lo_bits, uint32_t hi_bits, uint32_t stride);

#ifdef __cplusplus
}
#endif
#endif
