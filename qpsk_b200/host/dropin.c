/*
 * dropin.c -- the reference's single-channel API (include/qpsk_dropin.h) on top of the batch
 * C-ABI (include/qpsk_b200.h).  C11, because the reference's prototypes use C99 complex types
 * (costas_loop.h:17 takes `complex float` by value, fft.h:46-49 `complex double *`).
 *
 * Singletons, like the reference (rrc_fir.c:12, costas_loop.c:13-23, bit-scramble.c:41-42,
 * qpsk.c:36-53): one default context per subsystem, created on first use, batch of one.
 */
#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qpsk_b200.h"
#include "qpsk_dropin.h"
#include "host_design.h"

#define CYCLES_MAX 8
static const double kTau = 2.0 * 3.14159265358979323846;

static void die(const char *what, int rc) {
    fprintf(stderr, "qpsk_b200 drop-in: %s failed (%d): %s\n", what, rc, qpsk_b200_last_error());
    abort();   /* the reference's signatures are void: fail loudly, never fall back to the CPU */
}
#define MUST(call) do { int rc__ = (call); if (rc__ != 0) die(#call, rc__); } while (0)

/* ============================================================================================
 * rrc_fir.h
 * ========================================================================================== */
static float g_taps[NTAPS];
static int g_taps_valid = 0;
static qpsk_b200_fir *g_fir = NULL;

void rrc_make(float fs, float rs, float alpha) {                     /* rrc_fir.c:32-76 */
    qpsk_host_rrc_make(g_taps, NTAPS, fs, rs, alpha);
    g_taps_valid = 1;
    if (g_fir) { qpsk_b200_fir_destroy(g_fir); g_fir = NULL; }       /* new coefficients */
}

void rrc_fir(complex float memory[], complex float sample[], int length) {   /* rrc_fir.c:17-30 */
    if (length <= 0) return;
    if (!g_taps_valid) memset(g_taps, 0, sizeof g_taps);             /* static coeffs[] start at zero */
    if (!g_fir) MUST(qpsk_b200_fir_create(g_taps, NTAPS, 1, QPSK_B200_MODE_EXACT, 0, &g_fir));
    MUST(qpsk_b200_fir_set_memory(g_fir, (const float *)memory));    /* the caller owns the delay line */
    MUST(qpsk_b200_fir_process_host(g_fir, (float *)sample, length));
    MUST(qpsk_b200_fir_get_memory(g_fir, (float *)memory));
}

/* ============================================================================================
 * costas_loop.h -- scalar control state (costas_loop.c:13-23)
 * ========================================================================================== */
static qpsk_host_loop g_loop;

void update_gains(void) { qpsk_host_loop_update_gains(&g_loop); }    /* :49-54 */

void phase_wrap(void) {                                              /* :61-67 */
    while ((double)g_loop.phase > kTau) g_loop.phase = (float)((double)g_loop.phase - kTau);
    while ((double)g_loop.phase < -kTau) g_loop.phase = (float)((double)g_loop.phase + kTau);
}
void frequency_limit(void) {                                         /* :69-74 */
    if (g_loop.freq > g_loop.max_freq) g_loop.freq = g_loop.max_freq;
    else if (g_loop.freq < g_loop.min_freq) g_loop.freq = g_loop.min_freq;
}
float phase_detector(complex float s) {                              /* :44-47 */
    return (crealf(s) > 0.0f ? 1.0f : -1.0f) * cimagf(s) - (cimagf(s) > 0.0f ? 1.0f : -1.0f) * crealf(s);
}
void advance_loop(float error) {                                     /* :56-59 */
    g_loop.freq = g_loop.freq + g_loop.beta * error;
    g_loop.phase = g_loop.phase + g_loop.freq + g_loop.alpha * error;
}
/* the setters' range checks are dead stores in the reference (:79-115): the argument is kept */
void set_loop_bandwidth(float bw) { g_loop.loop_bw = bw; update_gains(); }
void set_damping_factor(float df) { g_loop.damping = df; update_gains(); }
void set_alpha(float a) { g_loop.alpha = a; }
void set_beta(float b) { g_loop.beta = b; }
void set_frequency(float f) {                                        /* :117-125 */
    if (f > g_loop.max_freq) g_loop.freq = g_loop.max_freq;
    else if (f < g_loop.min_freq) g_loop.freq = g_loop.min_freq;
    else g_loop.freq = f;
}
void set_phase(float p) { g_loop.phase = p; phase_wrap(); }          /* :127-132 */
void set_max_freq(float f) { g_loop.max_freq = f; }
void set_min_freq(float f) { g_loop.min_freq = f; }
float get_loop_bandwidth(void) { return g_loop.loop_bw; }
float get_damping_factor(void) { return g_loop.damping; }
float get_alpha(void) { return g_loop.alpha; }
float get_beta(void) { return g_loop.beta; }
float get_frequency(void) { return g_loop.freq; }
float get_phase(void) { return g_loop.phase; }
float get_max_freq(void) { return g_loop.max_freq; }
float get_min_freq(void) { return g_loop.min_freq; }

void create_control_loop(float loop_bw, float min_freq, float max_freq) {   /* :31-42, same call order */
    set_phase(0.0f);
    set_frequency(0.0f);
    set_max_freq(max_freq);
    set_min_freq(min_freq);
    set_damping_factor(sqrtf(2.0f) / 2.0f);
    set_loop_bandwidth(loop_bw);
}

/* ============================================================================================
 * qpsk.c entry points
 * ========================================================================================== */
static qpsk_b200_rx *g_rx = NULL;
static qpsk_b200_tx *g_tx = NULL;
static double g_rs = 2400.0, g_center = 1500.0, g_tx_carrier = 1500.0, g_tx_rs = 2400.0;
static complex float g_costas_frame[FRAME_SIZE];
static int g_rx_bits[2 * FRAME_SIZE];
static float g_offset_freq;

void qpsk_dropin_rx_reset(double rs, double center_hz) {
    g_rs = rs; g_center = center_hz;
    if (g_rx) { qpsk_b200_rx_destroy(g_rx); g_rx = NULL; }
}
void qpsk_dropin_tx_reset(double rs, double carrier_hz) {
    g_tx_rs = rs; g_tx_carrier = carrier_hz;
    if (g_tx) { qpsk_b200_tx_destroy(g_tx); g_tx = NULL; }
}
const complex float *qpsk_dropin_costas_frame(void) { return g_costas_frame; }
const int *qpsk_dropin_rx_bits(void) { return g_rx_bits; }
float qpsk_dropin_offset_freq(void) { return g_offset_freq; }

void rx_frame(int16_t in[FRAME_SIZE]) {                              /* qpsk.c:88-218 */
    if (!g_rx) {
        qpsk_b200_rx_config cfg;
        qpsk_b200_rx_default_config(&cfg);
        cfg.rs = (float)g_rs;
        cfg.center = (float)g_center;
        cfg.flags = QPSK_B200_KEEP_SYMBOLS;
        MUST(qpsk_b200_rx_create(&cfg, 1, 1, &g_rx));
    }
    /* the loop singleton is the caller's to set: hand its gains and state to the device, take the state back */
    float st[2] = { g_loop.phase, g_loop.freq };
    MUST(qpsk_b200_rx_set_loop(g_rx, g_loop.alpha, g_loop.beta, g_loop.min_freq, g_loop.max_freq));
    MUST(qpsk_b200_rx_set_loop_state(g_rx, st));
    uint8_t packed[FRAME_SIZE / 4];
    MUST(qpsk_b200_rx_process_host(g_rx, in, 1, packed));
    const int nsym = (int)(qpsk_b200_rx_output_bytes(g_rx, QPSK_B200_OUT_DIBITS) * 4);
    MUST(qpsk_b200_rx_read(g_rx, QPSK_B200_OUT_SYMBOLS, g_costas_frame, (size_t)nsym * sizeof(complex float)));
    MUST(qpsk_b200_rx_get_loop_state(g_rx, st));
    g_loop.phase = st[0];
    g_loop.freq = st[1];
    for (int i = 0; i < nsym; i++) {
        const int d = (packed[i / 4] >> (2 * (i % 4))) & 3;
        g_rx_bits[2 * i] = d & 1;
        g_rx_bits[2 * i + 1] = d >> 1;
    }
    g_offset_freq = (float)((double)g_loop.freq * g_rs / kTau);      /* qpsk.c:217 */
}

complex float qpsk_mod(int bits[2]) {                                /* qpsk.c:58-63, 269-271 */
    static const float pts[4][2] = { { 1.0f, 0.0f }, { 0.0f, 1.0f }, { 0.0f, -1.0f }, { -1.0f, 0.0f } };
    const int k = (bits[1] << 1) | bits[0];
    return pts[k][0] + pts[k][1] * I;
}

void qpsk_demod(complex float symbol, int bits[2]) {                 /* qpsk.c:74-79 */
    float r[2];
    qpsk_host_cis(3.14159265358979323846 / 4.0, 0, r);
    const float sr = crealf(symbol), si = cimagf(symbol);
    const float re = sr * r[0] - si * r[1], im = sr * r[1] + si * r[0];
    bits[0] = re < 0.0f;
    bits[1] = im < 0.0f;
}

static void need_tx(void) {
    if (!g_tx) {
        const float carrier = (float)g_tx_carrier;
        /* one packet per call: the phasor is renormalised at the end of every tx_frame (qpsk.c:253), whatever its
         * length: tx_frame / qpsk_packet_mod end the packet themselves (qpsk_b200_tx_end_packet) */
        MUST(qpsk_b200_tx_create(9600.0f, (float)g_tx_rs, .35f, &carrier, 1, 1 << 20, 0, &g_tx));
    }
}

int tx_frame(int16_t samples[], complex float symbol[], int length) {        /* qpsk.c:225-264 */
    need_tx();
    MUST(qpsk_b200_tx_symbols_host(g_tx, (const float *)symbol, length, samples));
    MUST(qpsk_b200_tx_end_packet(g_tx));
    return length * (int)(9600.0 / g_tx_rs);
}

int qpsk_packet_mod(int16_t samples[], int tx_bits[], int length) {          /* qpsk.c:273-285 */
    if (length <= 0) return 0;
    uint8_t *idx = (uint8_t *)malloc((size_t)length);
    if (!idx) die("malloc", -1);
    for (int i = 0, s = 0; i < length; i++, s += 2)
        idx[i] = (uint8_t)(((tx_bits[s] & 1) << 1) | (tx_bits[s + 1] & 1));   /* dibit[1] = tx_bits[s], dibit[0] = tx_bits[s+1] */
    need_tx();
    const int rc = qpsk_b200_tx_process_host(g_tx, idx, length, samples);
    free(idx);
    if (rc) die("qpsk_b200_tx_process_host", rc);
    MUST(qpsk_b200_tx_end_packet(g_tx));
    return length * (int)(9600.0 / g_tx_rs);
}

/* ============================================================================================
 * algorithms/fft.h -- complex double in the reference, FP32 on the GPU (<= 1e-5)
 * ========================================================================================== */
static qpsk_b200_fft *g_fft = NULL;
static int g_fft_n = 0;

static void fft_any(complex double *in, complex double *out, int n, int inverse) {
    if (n == 1) { out[0] = in[0]; return; }
    if (n > 8192) {                                        /* the reference takes any power of two: four-step path */
        float *big = (float *)malloc(sizeof(float) * 2 * (size_t)n);
        if (!big) die("malloc", -1);
        for (int i = 0; i < n; i++) { big[2 * i] = (float)creal(in[i]); big[2 * i + 1] = (float)cimag(in[i]); }
        const int brc = qpsk_b200_fft_big_host(big, big, n, inverse, 0);
        if (brc) { free(big); die("qpsk_b200_fft_big_host", brc); }
        for (int i = 0; i < n; i++) out[i] = (double)big[2 * i] + (double)big[2 * i + 1] * I;
        free(big);
        return;
    }
    if (n != g_fft_n) {
        if (g_fft) { qpsk_b200_fft_destroy(g_fft); g_fft = NULL; }
        MUST(qpsk_b200_fft_create(n, 0, &g_fft));
        g_fft_n = n;
    }
    float *buf = (float *)malloc(sizeof(float) * 2 * (size_t)n);
    if (!buf) die("malloc", -1);
    for (int i = 0; i < n; i++) { buf[2 * i] = (float)creal(in[i]); buf[2 * i + 1] = (float)cimag(in[i]); }
    const int rc = qpsk_b200_fft_transform_host(g_fft, buf, buf, 1, inverse);
    if (rc) { free(buf); die("qpsk_b200_fft_transform_host", rc); }
    for (int i = 0; i < n; i++) out[i] = (double)buf[2 * i] + (double)buf[2 * i + 1] * I;
    free(buf);
}
void fft(complex double *in, complex double *out) { fft_any(in, out, NFFT, 0); }              /* fft.c:98-108 */
void fftn(complex double *in, complex double *out, int n) { fft_any(in, out, n, 0); }         /* :110-120 */
void ifft(complex double *in, complex double *out) { fft_any(in, out, NFFT, 1); }             /* :122-128 */
void ifftn(complex double *in, complex double *out, int n) { fft_any(in, out, n, 1); }        /* :130-136 */

/* ============================================================================================
 * algorithms/crc16.h, interleave.h, bit-scramble.h
 * ========================================================================================== */
uint16_t crc16(const uint8_t *data, int length) {                    /* crc16.c:11-23 */
    uint16_t crc = 0xFFFF;
    if (length <= 0) return crc;
    MUST(qpsk_b200_bits_crc16(data, length, 1, &crc, 0));
    return crc;
}

void interleave(uint8_t *inout, int nbytes, int dir) {               /* interleave.c:43-78 */
    if (nbytes <= 0) return;
    MUST(qpsk_b200_bits_interleave(inout, nbytes, 1, dir, 0));
}

static uint16_t g_scr_tx, g_scr_rx;                                  /* bit-scramble.c:41-42 */
void scramble_init(SRegister sr) {                                   /* :46-55 */
    if (sr == tx || sr == both) g_scr_tx = 0x4A80;
    if (sr == rx || sr == both) g_scr_rx = 0x4A80;
}
int scramble(uint8_t *input, SRegister sr) {                         /* :57-84 */
    uint16_t *m;
    if (sr == tx) m = &g_scr_tx; else if (sr == rx) m = &g_scr_rx; else return -1;
    for (int i = 0; i < 2; i++) {
        const uint16_t key = (uint16_t)(((*m >> 1) ^ *m) & 1u);
        *input = (uint8_t)(*input ^ (key << i));
        *m = (uint16_t)((*m >> 1) | (key << 14));
    }
    return 0;
}
