/*
 * qpsk_oracle.c -- CPU restatement of the MonsieurETM/QPSK hot path (see qpsk_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY; never linked into the product library.
 * Parity status: pinned against the compiled reference (oracle/_ref) -- see the header.
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (no -march, no -ffast-math): every float
 * operation below is then one IEEE-754 binary32 operation, exactly as in the reference's
 * Makefile build (Makefile:7, -std=c11 implies -ffp-contract=off).
 */
#include <math.h>
#include <string.h>

#include "qpsk_oracle.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define ORC_TAU (2.0 * M_PI) /* qpsk.h:30 */

/* --------------------------------------------------------------------------------------
 * complex helpers, spelled out the way GCC evaluates C99 complex arithmetic on finite data
 * (SURVEY.md Appendix A): complex*complex = (ar*br - ai*bi, ar*bi + ai*br), four rounded
 * products and two rounded sums; complex*real and complex/real act per component.
 * ------------------------------------------------------------------------------------ */
static inline orc_cf cf_mul(orc_cf a, orc_cf b) {
    float ac = a.re * b.re, bd = a.im * b.im, ad = a.re * b.im, bc = a.im * b.re;
    orc_cf r = { ac - bd, ad + bc };
    return r;
}

/* qpsk.h:35 cmplx(v) and qpsk.h:36 cmplxconj(v): the argument is converted to float first */
static inline orc_cf cf_cis(float v) { orc_cf r = { cosf(v), sinf(v) }; return r; }
static inline orc_cf cf_cis_conj(float v) { orc_cf r = { cosf(v), -sinf(v) }; return r; }

/* ======================================================================================
 * RRC taps -- rrc_fir.c:32-76.  Float expressions with double sub-expressions wherever the
 * reference mixes in M_PI or GAIN (both double).
 * ==================================================================================== */
void orc_rrc_make(float *taps, int ntaps, float fs, float rs, float alpha) {
    const float spb = fs / rs;                                   /* rrc_fir.c:34 */
    float scale = 0.f;

    for (int i = 0; i < ntaps; i++) {
        const float xindx = (float)(i - ntaps / 2);              /* :39 integer subtract, then to float */
        const float x1 = (float)(M_PI * (double)xindx / (double)spb);  /* :40 double product and quotient */
        float x2 = 4.f * alpha * xindx / spb;                    /* :41 */
        float x3 = x2 * x2 - 1.f;                                /* :42 */
        float num, den;

        if (fabsf(x3) >= 0.000001f) {                            /* :44 regular branch */
            const float c = cosf((1.f + alpha) * x1);
            if (i != ntaps / 2)                                  /* :45-47 */
                num = c + sinf((1.f - alpha) * x1) / (4.f * alpha * xindx / spb);
            else                                                 /* :49 centre tap: double tail */
                num = (float)((double)c + (double)(1.f - alpha) * M_PI / (double)(4.f * alpha));
            den = (float)((double)x3 * M_PI);                    /* :51 */
        } else {                                                 /* :52 singular points */
            if (alpha == 1.f) {                                  /* :53-57 */
                taps[i] = -1.f;
                scale += taps[i];
                continue;
            }
            x3 = (1.f - alpha) * x1;                             /* :59 */
            x2 = (1.f + alpha) * x1;                             /* :60 */
            const double t1 = (double)(sinf(x2) * (1.f + alpha)) * M_PI;                 /* :62 */
            const double t2 = (double)cosf(x3) * ((double)(1.f - alpha) * M_PI * (double)spb)
                              / (double)(4.f * alpha * xindx);                            /* :63 */
            const float  t3 = sinf(x3) * spb * spb / (4.f * alpha * xindx * xindx);       /* :64 */
            num = (float)(t1 - t2 + (double)t3);
            den = (float)((double)-32.f * M_PI * (double)alpha * (double)alpha * (double)xindx / (double)spb); /* :66 */
        }
        taps[i] = 4.f * alpha * num / den;                       /* :69 */
        scale += taps[i];                                        /* :70 */
    }
    for (int i = 0; i < ntaps; i++)                              /* :73-75 GAIN is double */
        taps[i] = (float)(((double)taps[i] * ORC_GAIN) / (double)scale);
}

/* ======================================================================================
 * FIR -- rrc_fir.c:17-30.  memory[ntaps-1] is the newest input; the sum runs oldest tap
 * first from +0; the output gain is a double multiply per component.  In place: the
 * delay line only ever holds inputs, so outputs never feed back.
 * ==================================================================================== */
void orc_rrc_fir(const float *taps, int ntaps, orc_cf *memory, orc_cf *sample, int length) {
    for (int j = 0; j < length; j++) {
        memmove(&memory[0], &memory[1], (size_t)(ntaps - 1) * sizeof(orc_cf));
        memory[ntaps - 1] = sample[j];
        float yr = 0.0f, yi = 0.0f;
        for (int i = 0; i < ntaps; i++) {
            yr += memory[i].re * taps[i];
            yi += memory[i].im * taps[i];
        }
        sample[j].re = (float)((double)yr * ORC_GAIN);
        sample[j].im = (float)((double)yi * ORC_GAIN);
    }
}

/* ======================================================================================
 * Costas control loop -- costas_loop.c.  The setters' range checks are dead stores
 * (costas_loop.c:79-115): the effective behaviour is "store the argument".
 * ==================================================================================== */
void orc_loop_update_gains(orc_loop *l) {                        /* :49-54 */
    const float denom = (1.0f + (2.0f * l->damping * l->loop_bw)) + (l->loop_bw * l->loop_bw);
    l->alpha = (4.0f * l->damping * l->loop_bw) / denom;
    l->beta = (4.0f * l->loop_bw * l->loop_bw) / denom;
}

void orc_loop_phase_wrap(orc_loop *l) {                          /* :61-67 double compare and subtract */
    while ((double)l->phase > ORC_TAU) l->phase = (float)((double)l->phase - ORC_TAU);
    while ((double)l->phase < -ORC_TAU) l->phase = (float)((double)l->phase + ORC_TAU);
}

void orc_loop_frequency_limit(orc_loop *l) {                     /* :69-74 */
    if (l->freq > l->max_freq) l->freq = l->max_freq;
    else if (l->freq < l->min_freq) l->freq = l->min_freq;
}

void orc_loop_set_frequency(orc_loop *l, float f) {              /* :117-125 */
    if (f > l->max_freq) l->freq = l->max_freq;
    else if (f < l->min_freq) l->freq = l->min_freq;
    else l->freq = f;
}

void orc_loop_set_phase(orc_loop *l, float p) { l->phase = p; orc_loop_phase_wrap(l); } /* :127-132 */

void orc_loop_create(orc_loop *l, float loop_bw, float min_freq, float max_freq) { /* :31-42 */
    /* the reference's statics are zero-initialised, and set_frequency(0) runs before the limits exist */
    memset(l, 0, sizeof *l);
    orc_loop_set_phase(l, 0.0f);
    orc_loop_set_frequency(l, 0.0f);
    l->max_freq = max_freq;
    l->min_freq = min_freq;
    l->damping = sqrtf(2.0f) / 2.0f;
    orc_loop_update_gains(l);          /* set_damping_factor calls update_gains with loop_bw still 0 */
    l->loop_bw = loop_bw;
    orc_loop_update_gains(l);
}

float orc_phase_detector(orc_cf s) {                             /* :44-47; zero maps to -1 */
    return (s.re > 0.0f ? 1.0f : -1.0f) * s.im - (s.im > 0.0f ? 1.0f : -1.0f) * s.re;
}

void orc_loop_advance(orc_loop *l, float error) {                /* :56-59 */
    l->freq = l->freq + l->beta * error;
    l->phase = l->phase + l->freq + l->alpha * error;
}

/* ======================================================================================
 * modem profile / state
 * ==================================================================================== */
void orc_profile_init(orc_profile *p, float fs, float rs, float center, float rrc_alpha,
                      int ntaps, int frame_size, float loop_bw, int ub_mode) {
    memset(p, 0, sizeof *p);
    p->fs = fs; p->rs = rs; p->center = center;
    p->sps = (int)((double)fs / (double)rs);                     /* qpsk.h:21 */
    p->frame_size = frame_size;
    p->nsym = frame_size / p->sps;
    p->ntaps = ntaps;
    p->ub_mode = ub_mode;
    orc_rrc_make(p->taps, ntaps, fs, rs, rrc_alpha);             /* qpsk.c:308 */
    p->rx_rect = cf_cis_conj((float)(ORC_TAU * (double)center / (double)fs)); /* qpsk.c:342 */
    p->rot45 = cf_cis((float)(M_PI / 4.0));                      /* qpsk.c:75, qpsk.h:31 */
    orc_loop_create(&p->loop0, loop_bw, -1.0f, 1.0f);            /* qpsk.c:302 */
}

/* Extension (not the reference): slice the loop output where phase_detector locks it, on the
 * diagonals, i.e. without qpsk_demod's extra 45 degrees that leaves one bit on a decision boundary
 * (SURVEY finding 3).  Multiplying by 1+0i keeps orc_qpsk_demod's arithmetic shape. */
void orc_profile_slice_diagonal(orc_profile *p, int on) {
    if (on) { p->rot45.re = 1.0f; p->rot45.im = 0.0f; }
    else p->rot45 = cf_cis((float)(M_PI / 4.0));
}

void orc_rx_state_init(const orc_profile *p, orc_rx_state *s) {
    memset(s, 0, sizeof *s);
    s->rx_phase = cf_cis(0.0f);                                  /* qpsk.c:341 */
    s->phase = p->loop0.phase;
    s->freq = p->loop0.freq;
}

void orc_qpsk_demod(const orc_profile *p, orc_cf sym, int bits[2]) { /* qpsk.c:74-79 */
    const orc_cf r = cf_mul(sym, p->rot45);
    bits[0] = r.re < 0.0f;
    bits[1] = r.im < 0.0f;
}

/* ======================================================================================
 * rx_frame -- qpsk.c:88-218
 * ==================================================================================== */
void orc_rx_frame(const orc_profile *p, orc_rx_state *s, const int16_t *pcm, const orc_rx_trace *t) {
    const int N = p->frame_size, sps = p->sps, nsym = p->nsym;
    orc_cf frame[ORC_MAX_FRAME];

    /* 1. mixer, qpsk.c:114-120 */
    for (int i = 0; i < N; i++) {
        s->rx_phase = cf_mul(s->rx_phase, p->rx_rect);
        const float v = (float)pcm[i] / 16384.0f;
        frame[i].re = s->rx_phase.re * v;
        frame[i].im = s->rx_phase.im * v;
    }
    {   /* cabsf -> glibc hypotf == (float)sqrt(re^2+im^2 in double) for finite input */
        const float mag = hypotf(s->rx_phase.re, s->rx_phase.im);
        s->rx_phase.re = s->rx_phase.re / mag;
        s->rx_phase.im = s->rx_phase.im / mag;
    }

    /* 2. matched filter, qpsk.c:125 */
    orc_rrc_fir(p->taps, p->ntaps, s->fir_mem, frame, N);
    if (t && t->fir) memcpy(t->fir, frame, (size_t)N * sizeof(orc_cf));

    /* 3. amplitude histograms, qpsk.c:131-180.  The running means are NOT reset per symbol. */
    float max_i = 0.0f, max_q = 0.0f, av_i = 0.0f, av_q = 0.0f;
    int hist_i[8] = { 0 }, hist_q[8] = { 0 };
    for (int i = 0; i < N; i += sps) {
        for (int j = 0; j < sps; j++) {
            av_i += fabsf(frame[i + j].re);
            av_q += fabsf(frame[i + j].im);
        }
        av_i /= (float)sps;
        av_q /= (float)sps;
        if (av_i > max_i) max_i = av_i;
        if (av_q > max_q) max_q = av_q;
        const float hv_i = max_i / 8.0f, hv_q = max_q / 8.0f;
        for (int k = 1; k < 8; k++) if (av_i <= hv_i * (float)k) { hist_i[k]++; break; }
        for (int k = 1; k < 8; k++) if (av_q <= hv_q * (float)k) { hist_q[k]++; break; }
    }
    int hmax = 0, index = 0;
    for (int k = 0; k < 8; k++) {
        const int h = hist_i[k] + hist_q[k];
        if (h > hmax) { hmax = h; index = k; }
    }
    if (p->ub_mode == ORC_UB_TAU) {                              /* extension: the sample nearest to the eye's maximum */
        orc_cf S;
        orc_timing_sum(frame, N, sps, &S);
        index = orc_tau_index(S, sps);
    }
    if (t && t->index) *t->index = index;

    /* 4. decimate with a one-frame delay, qpsk.c:186-191.  For sps == 4 and index >= 4 the
     * last symbol reads past input_frame[]; in the Makefile build that memory is
     * decimated_frame[index-4], which this very loop has already refreshed (SURVEY finding 1). */
    for (int i = 0; i < nsym; i++) {
        s->dec[i] = s->dec[nsym + i];
        /* ORC_UB_PHASE (extension, not the reference): the histogram picks a sampling phase only */
        const int j = i * sps + (p->ub_mode == ORC_UB_PHASE ? index % sps : index);
        if (j < N) s->dec[nsym + i] = frame[j];
        else if (p->ub_mode == ORC_UB_ALIAS) s->dec[nsym + i] = s->dec[j - N];
        else s->dec[nsym + i] = frame[N - 1];
    }
    if (t && t->dec) memcpy(t->dec, &s->dec[nsym], (size_t)nsym * sizeof(orc_cf));

    /* 5. Costas loop over the PREVIOUS frame's symbols, qpsk.c:196-212 */
    orc_loop l = p->loop0;
    l.phase = s->phase; l.freq = s->freq;
    for (int i = 0; i < nsym; i++) {
        const orc_cf c = cf_mul(s->dec[i], cf_cis_conj(l.phase));
        if (t && t->costas) t->costas[i] = c;
        const float err = orc_phase_detector(c);
        orc_loop_advance(&l, err);
        orc_loop_phase_wrap(&l);
        orc_loop_frequency_limit(&l);
        if (t && t->dibit) {
            int b[2];
            orc_qpsk_demod(p, c, b);
            t->dibit[i] = (uint8_t)(b[0] | (b[1] << 1));
        }
    }
    s->phase = l.phase; s->freq = l.freq;
    if (t && t->phase) *t->phase = l.phase;
    if (t && t->freq) *t->freq = l.freq;
    if (t && t->offset_hz) *t->offset_hz = (float)((double)l.freq * (double)p->rs / ORC_TAU); /* :217 */
}

void orc_rx_run(const orc_profile *p, orc_rx_state *states, const int16_t *pcm, int nchan, int nframes,
                orc_cf *fir, int32_t *index, orc_cf *dec, orc_cf *costas, uint8_t *dibit,
                float *phase, float *freq) {
    const size_t N = (size_t)p->frame_size, S = (size_t)p->nsym, F = (size_t)nframes;
    for (size_t c = 0; c < (size_t)nchan; c++) {
        for (size_t f = 0; f < F; f++) {
            orc_rx_trace t = { 0 };
            const size_t u = c * F + f;
            if (fir) t.fir = fir + u * N;
            if (index) t.index = index + u;
            if (dec) t.dec = dec + u * S;
            if (costas) t.costas = costas + u * S;
            if (dibit) t.dibit = dibit + u * S;
            if (phase) t.phase = phase + u;
            if (freq) t.freq = freq + u;
            orc_rx_frame(p, &states[c], pcm + u * N, &t);
        }
    }
}

/* Extension (not in the reference): the test channel's noise, restated from qpsk_b200/csrc/channel.cuh.
 * Philox-4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11), counter =
 * (sample >> 1, channel, sample >> 33, 0), key = seed; four 16-bit halves per sample -> Irwin-Hall(4). */
static void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_awgn(int16_t *pcm, long long nsamples, float sigma, uint64_t seed, long long first_sample, int channel) {
    for (long long t = 0; t < nsamples; t++) {
        const uint64_t n = (uint64_t)(first_sample + t);
        uint32_t w[4];
        orc_philox((uint32_t)(n >> 1), (uint32_t)channel, (uint32_t)(n >> 33), 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
        const uint32_t w0 = w[2 * (n & 1)], w1 = w[2 * (n & 1) + 1];
        const int s = (int)(w0 & 0xffffu) + (int)(w0 >> 16) + (int)(w1 & 0xffffu) + (int)(w1 >> 16);
        const float g = (float)(s - 131070) * (1.0f / 37837.0f);
        float v = (float)pcm[t] + sigma * g;
        v = truncf(v);
        if (v < -32768.0f) v = -32768.0f;
        if (v > 32767.0f) v = 32767.0f;
        pcm[t] = (int16_t)(int)v;
    }
}

/* Extension (PARITY UNPINNED, not in the reference): spectral-line timing statistic of one
 * filtered frame, S = sum_n y_n^2 e^{-2 pi i n / sps} taken per component and added at the end
 * (Oerder & Meyr's square-law estimator; tau = -arg(S) sps / (2 pi) samples).  Single rounded
 * float operations in exactly the order of the CUDA timing warps (rx_front.cuh). */
static void timing_half(const float *y, int stride, int n, int sps, float *re_out, float *im_out) {
    float re = 0.0f, im = 0.0f;
    for (int k = 0; k < n; k++) {
        const float v = y[(size_t)k * stride];
        const float p = v * v;
        const int j = k % sps;
        if (sps == 4) {
            if (j == 0) re = re + p; else if (j == 1) im = im - p; else if (j == 2) re = re - p; else im = im + p;
        } else {
            const float t = p * 0.70710678118654752f;
            switch (j) {
                case 0: re = re + p; break;
                case 1: re = re + t; im = im - t; break;
                case 2: im = im - p; break;
                case 3: re = re - t; im = im - t; break;
                case 4: re = re - p; break;
                case 5: re = re - t; im = im + t; break;
                case 6: im = im + p; break;
                default: re = re + t; im = im + t; break;
            }
        }
    }
    *re_out = re; *im_out = im;
}

/* round(tau) mod sps, tau = -arg(S) sps / (2 pi), as the sector of S: comparisons and (at 8 samples per symbol) one rounded
 * multiply, in the order of tau_index() in rx_front.cuh */
int orc_tau_index(orc_cf S, int sps) {
    const float ax = fabsf(S.re), ay = fabsf(S.im);
    if (sps == 4) {
        if (ax >= ay) return S.re >= 0.0f ? 0 : 2;
        return S.im < 0.0f ? 1 : 3;
    }
    const float T = 0.41421356237309503f;
    if (ay <= T * ax) return S.re >= 0.0f ? 0 : 4;
    if (ax <= T * ay) return S.im < 0.0f ? 2 : 6;
    if (S.im < 0.0f) return S.re > 0.0f ? 1 : 3;
    return S.re > 0.0f ? 7 : 5;
}

void orc_timing_sum(const orc_cf *frame, int n, int sps, orc_cf *out) {
    float ri, ii, rq, iq;
    timing_half(&frame[0].re, 2, n, sps, &ri, &ii);
    timing_half(&frame[0].im, 2, n, sps, &rq, &iq);
    out->re = ri + rq;
    out->im = ii + iq;
}

/* ======================================================================================
 * transmit -- qpsk.c:58-63 (constellation), :225-264 (tx_frame), :269-285
 * ==================================================================================== */
void orc_tx_state_init(orc_tx_state *s, float carrier_hz, float fs) {
    memset(s, 0, sizeof *s);
    s->tx_phase = cf_cis(0.0f);                                              /* qpsk.c:316 */
    s->tx_rect = cf_cis((float)(ORC_TAU * (double)carrier_hz / (double)fs)); /* qpsk.c:320 */
}

/* a new carrier from the next packet on, phase kept: the assignment of qpsk.c:320 again */
void orc_tx_set_carrier(orc_tx_state *s, float carrier_hz, float fs) {
    s->tx_rect = cf_cis((float)(ORC_TAU * (double)carrier_hz / (double)fs));
}

orc_cf orc_qpsk_mod(const int bits[2]) {
    static const orc_cf points[4] = { { 1.0f, 0.0f }, { 0.0f, 1.0f }, { 0.0f, -1.0f }, { -1.0f, 0.0f } };
    return points[(bits[1] << 1) | bits[0]];
}

int orc_tx_frame(const orc_profile *p, orc_tx_state *s, int16_t *samples, const orc_cf *symbol, int length) {
    const int sps = p->sps, n = length * sps;
    orc_cf signal[n > 0 ? n : 1];
    for (int i = 0; i < length; i++) {                          /* :232-238 zero stuffing */
        signal[i * sps] = symbol[i];
        for (int j = 1; j < sps; j++) { signal[i * sps + j].re = 0.0f; signal[i * sps + j].im = 0.0f; }
    }
    orc_rrc_fir(p->taps, p->ntaps, s->fir_mem, signal, n);      /* :243 */
    for (int i = 0; i < n; i++) {                               /* :248-251 */
        s->tx_phase = cf_mul(s->tx_phase, s->tx_rect);
        signal[i] = cf_mul(signal[i], s->tx_phase);
    }
    {
        const float mag = hypotf(s->tx_phase.re, s->tx_phase.im);   /* :253 */
        s->tx_phase.re = s->tx_phase.re / mag;
        s->tx_phase.im = s->tx_phase.im / mag;
    }
    for (int i = 0; i < n; i++)                                 /* :259-261 truncation toward zero */
        samples[i] = (int16_t)(int32_t)(signal[i].re * 16384.0f);
    return n;
}

int orc_qpsk_packet_mod(const orc_profile *p, orc_tx_state *s, int16_t *samples, const int *tx_bits, int length) {
    orc_cf symbol[length > 0 ? length : 1];
    for (int i = 0, k = 0; i < length; i++, k += 2) {           /* :277-282 */
        int dibit[2] = { tx_bits[k + 1] & 1, tx_bits[k] & 1 };
        symbol[i] = orc_qpsk_mod(dibit);
    }
    return orc_tx_frame(p, s, samples, symbol, length);
}

/* ======================================================================================
 * FFT -- algorithms/fft.c.  The reference recurses (even/odd split, then butterflies with a
 * freshly evaluated double twiddle cos/sin(TAU*m/n)).  The iterative decimation-in-time form
 * below performs exactly the same butterflies on exactly the same operands, level by level,
 * so results are bit-identical; n must be a power of two.
 * ==================================================================================== */
static void fft_core(orc_cd *v, int n, int inverse) {
    /* bit reversal = the composition of the reference's even/odd splits */
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { orc_cd tmp = v[i]; v[i] = v[j]; v[j] = tmp; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len / 2;
        for (int base = 0; base < n; base += len) {
            for (int m = 0; m < half; m++) {
                const double ang = ORC_TAU * (double)m / (double)len;   /* fft.c:55, :85 */
                const double wr = cos(ang);
                const double wi = inverse ? sin(ang) : -sin(ang);
                const orc_cd e = v[base + m], o = v[base + m + half];
                const double zr = wr * o.re - wi * o.im;                /* fft.c:57-58 */
                const double zi = wr * o.im + wi * o.re;
                v[base + m].re = e.re + zr;         v[base + m].im = e.im + zi;
                v[base + m + half].re = e.re - zr;  v[base + m + half].im = e.im - zi;
            }
        }
    }
}

void orc_fftn(const orc_cd *in, orc_cd *out, int n) {            /* fft.c:110-120 */
    if (out != in) memmove(out, in, (size_t)n * sizeof(orc_cd));
    fft_core(out, n, 0);
    for (int i = 0; i < n; i++) { out[i].re = out[i].re / (double)n; out[i].im = out[i].im / (double)n; }
}

void orc_ifftn(const orc_cd *in, orc_cd *out, int n) {           /* fft.c:130-136 */
    if (out != in) memmove(out, in, (size_t)n * sizeof(orc_cd));
    fft_core(out, n, 1);
}

int orc_fft_argmax(const orc_cd *x, int n, double *mag2) {       /* unpinned; mirrors qpsk.c:176 */
    int best = 0; double bm = -1.0;
    for (int k = 0; k < n; k++) {
        const double m = x[k].re * x[k].re + x[k].im * x[k].im;
        if (m > bm) { bm = m; best = k; }
    }
    if (mag2) *mag2 = bm;
    return best;
}

/* ======================================================================================
 * bit stages
 * ==================================================================================== */
void orc_scramble_dibit(uint16_t *reg, uint8_t *dibit) {         /* bit-scramble.c:57-69, BITS = 2 */
    for (int i = 0; i < 2; i++) {
        const uint16_t key = (uint16_t)(((*reg >> 1) ^ *reg) & 1u);
        *dibit = (uint8_t)(*dibit ^ (key << i));
        *reg = (uint16_t)((*reg >> 1) | (key << 14));
    }
}

static const uint16_t orc_primes[] = {                           /* interleave.c:33-41 */
    2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97,
    101, 103, 107, 109, 113, 127, 131, 137, 139, 149, 151, 157, 163, 167, 173, 179, 181, 191, 193,
    197, 199, 211, 223, 227, 229, 233, 239, 241, 251, 257, 263, 269, 271, 277, 281, 283, 293, 307,
    311, 313, 317, 331, 337, 347
};

int orc_interleave_prime(int nbytes) {                           /* interleave.c:48-55 */
    const int imax = (int)(sizeof orc_primes / sizeof orc_primes[0]);
    const uint16_t nbits = (uint16_t)(nbytes * 8);
    int index = 1;
    /* the reference tests primes[index] before index < imax; past the table the search stops either way */
    while (index < imax && orc_primes[index] < nbits) index++;
    return orc_primes[index - 1];
}

void orc_interleave(uint8_t *inout, int nbytes, int dir) {       /* interleave.c:43-78 */
    uint8_t out[nbytes > 0 ? nbytes : 1];
    memset(out, 0, (size_t)nbytes);
    const uint16_t nbits = (uint16_t)(nbytes * 8);
    const uint32_t b = (uint32_t)orc_interleave_prime(nbytes);
    for (uint32_t n = 0; n < nbits; n++) {
        uint32_t i = n, j = (b * n) % nbits;
        if (dir == 1) { uint32_t tmp = j; j = i; i = tmp; }
        const uint32_t bit = (inout[i / 8] >> (i % 8)) & 1u;
        out[j / 8] |= (uint8_t)(bit << (j % 8));
    }
    memcpy(inout, out, (size_t)nbytes);
}

uint16_t orc_crc16(const uint8_t *data, int length) {            /* crc16.c:11-23 */
    uint16_t crc = 0xFFFF;
    while (length-- > 0) {
        uint8_t x = (uint8_t)((crc >> 8) ^ *data++);
        x ^= (uint8_t)(x >> 4);
        crc = (uint16_t)((crc << 8) ^ ((uint16_t)(x << 12)) ^ ((uint16_t)(x << 5)) ^ ((uint16_t)x));
    }
    return crc;
}

/* ======================================================================================
 * glibc 2.39 sinf/cosf, x86-64 "fma" ifunc variant (sysdeps/ieee754/flt-32/s_sinf.c,
 * s_cosf.c, sincosf.h, sincosf_table.c built with -mfma -mavx2).  glibc is a dependency of
 * the reference that is not under /root/reference; the algorithm is restated here from the
 * published sources: reduce by pi/2 in double, degree-7/8 polynomials in double, every
 * a + b*c contracted into one fused multiply-add.  Valid for |y| < 120, which covers the
 * Costas NCO (|phase| <= TAU).  tests/test_oracle_libm.py checks it against the host libm.
 * ==================================================================================== */
static const double SC_HPI_INV = 0x1.45F306DC9C883p+23;  /* 2/pi * 2^24 */
static const double SC_HPI = 0x1.921FB54442D18p0;
static const double SC_C0 = 0x1p0, SC_C1 = -0x1.ffffffd0c621cp-2, SC_C2 = 0x1.55553e1068f19p-5,
                    SC_C3 = -0x1.6c087e89a359dp-10, SC_C4 = 0x1.99343027bf8c3p-16;
static const double SC_S1 = -0x1.555545995a603p-3, SC_S2 = 0x1.1107605230bc4p-7, SC_S3 = -0x1.994eb3774cf24p-13;

static inline double sc_reduce(double x, int *np) {
    const double r = x * SC_HPI_INV;
    const int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return fma(-(double)n, SC_HPI, x);
}
static inline float sc_sin_poly(double xs, double x2) {
    const double x3 = xs * x2;
    const double s1 = fma(x2, SC_S3, SC_S2);
    const double x7 = x3 * x2;
    const double s = fma(x3, SC_S1, xs);
    return (float)fma(x7, s1, s);
}
static inline float sc_cos_poly(double x2, int negate) {
    const double sg = negate ? -1.0 : 1.0;     /* second table row = negated C coefficients */
    const double x4 = x2 * x2;
    const double c2 = fma(x2, sg * SC_C4, sg * SC_C3);
    const double c1 = fma(x2, sg * SC_C1, sg * SC_C0);
    const double x6 = x4 * x2;
    const double c = fma(x4, sg * SC_C2, c1);
    return (float)fma(x6, c2, c);
}
static inline int sc_tiny(float y) {          /* abstop12(y) < abstop12(0x1p-12f): sinf returns y, cosf returns 1 */
    uint32_t u;
    memcpy(&u, &y, sizeof u);
    return ((u >> 20) & 0x7ff) < 0x398;
}
float orc_glibc_sinf(float y) {
    int n;
    if (sc_tiny(y)) return y;
    const double x = sc_reduce((double)y, &n);
    const double x2 = x * x;
    if ((n & 1) == 0) return sc_sin_poly(((n & 3) == 1 || (n & 3) == 2) ? -x : x, x2);
    return sc_cos_poly(x2, (n & 2) != 0);
}
float orc_glibc_cosf(float y) {
    int n;
    if (sc_tiny(y)) return 1.0f;
    const double x = sc_reduce((double)y, &n);
    const double x2 = x * x;
    if (n & 1) return sc_sin_poly(((n & 3) == 1 || (n & 3) == 2) ? -x : x, x2);
    return sc_cos_poly(x2, (n & 2) != 0);
}

/* mismatch count of the restated sinf/cosf against the host libm over the float bit patterns
 * lo..hi (both signs), every `stride`-th one */
long orc_glibc_check(uint32_t lo, uint32_t hi, uint32_t stride) {
    long bad = 0;
    for (uint64_t b = lo; b <= hi; b += stride) {
        for (uint32_t sign = 0; sign < 2; sign++) {
            uint32_t u = (uint32_t)b | (sign << 31);
            float y;
            memcpy(&y, &u, sizeof y);
            float s0 = sinf(y), c0 = cosf(y), s1 = orc_glibc_sinf(y), c1 = orc_glibc_cosf(y);
            if (memcmp(&s0, &s1, 4) != 0 || memcmp(&c0, &c1, 4) != 0) bad++;
        }
    }
    return bad;
}

/* ======================================================================================
 * Frame format -- PARITY UNPINNED: the reference never chains its bit stages (algorithms/ is
 * not even linked), so this composition is this project's own, built
 * only from the pinned primitives above:
 *
 *   frame[nbytes] = payload[nbytes-2] | crc16(payload) high byte | low byte
 *   interleave(frame, nbytes, INTERLEAVE)                       (interleave.c:43-78)
 *   dibit k = bit 2k | bit 2k+1 << 1, bits taken LSB first      (bit-scramble.c:57-69 bit order)
 *   scramble every dibit, register reset to SEED per frame      (bit-scramble.c:11)
 *
 * and the inverse on the receive side.  One frame = the dibits of one rx_frame call
 * (FRAME_SIZE / CYCLES symbols = 32 bytes at 2400 baud, 16 bytes at 1200 baud).
 * ==================================================================================== */
void orc_frame_encode(const uint8_t *payload, int nbytes, uint8_t *dibits /* [4*nbytes] */) {
    uint8_t frame[nbytes];
    memcpy(frame, payload, (size_t)nbytes - 2);
    const uint16_t crc = orc_crc16(payload, nbytes - 2);
    frame[nbytes - 2] = (uint8_t)(crc >> 8);
    frame[nbytes - 1] = (uint8_t)(crc & 0xff);
    orc_interleave(frame, nbytes, 0);
    uint16_t reg = ORC_SCRAMBLE_SEED;
    for (int k = 0; k < 4 * nbytes; k++) {
        uint8_t d = (uint8_t)((frame[k / 4] >> (2 * (k % 4))) & 3);
        orc_scramble_dibit(&reg, &d);
        dibits[k] = d;
    }
}

/* returns 1 when the CRC matches; frame[nbytes] receives payload | crc */
int orc_frame_decode(const uint8_t *dibits, int nbytes, uint8_t *frame) {
    uint16_t reg = ORC_SCRAMBLE_SEED;
    memset(frame, 0, (size_t)nbytes);
    for (int k = 0; k < 4 * nbytes; k++) {
        uint8_t d = (uint8_t)(dibits[k] & 3);
        orc_scramble_dibit(&reg, &d);
        frame[k / 4] |= (uint8_t)(d << (2 * (k % 4)));
    }
    orc_interleave(frame, nbytes, 1);
    const uint16_t crc = orc_crc16(frame, nbytes - 2);
    return frame[nbytes - 2] == (uint8_t)(crc >> 8) && frame[nbytes - 1] == (uint8_t)(crc & 0xff);
}

/* 90-degree ambiguity (PARITY UNPINNED, this project's composition).  The constellation is
 * index d -> {1, j, -j, -1} (qpsk.c:58-63) and, at the intended phase, qpsk_demod (qpsk.c:74-79)
 * returns dibit == index.  Turning the received symbols a quarter turn ahead maps
 * 1 -> j -> -1 -> -j -> 1, i.e. d -> 1, 3, 0, 2 for d = 0, 1, 2, 3. */
uint8_t orc_rotate_dibit(uint8_t d, int quarter_turns) {
    static const uint8_t ahead[4] = { 1, 3, 0, 2 };
    d &= 3;
    for (int r = 0; r < (quarter_turns & 3); r++) d = ahead[d];
    return d;
}

/* try rotations 0..3 in order: undo r quarter turns on every dibit, decode, accept the first CRC
 * match.  No match: frame holds the rotation-0 decode and the result is -1. */
int orc_frame_decode_rotated(const uint8_t *dibits, int nbytes, uint8_t *frame) {
    uint8_t turned[4 * nbytes], trial[nbytes];
    for (int r = 0; r < 4; r++) {
        for (int k = 0; k < 4 * nbytes; k++) turned[k] = orc_rotate_dibit(dibits[k], 4 - r);
        const int ok = orc_frame_decode(turned, nbytes, r == 0 ? frame : trial);
        if (ok) {
            if (r) memcpy(frame, trial, (size_t)nbytes);
            return r;
        }
    }
    return -1;
}
