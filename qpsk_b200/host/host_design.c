/* host_design.c -- see host_design.h.  Build: gcc -std=c11 -O2 -ffp-contract=off (no -march). */
#include <math.h>
#include <string.h>

#include "host_design.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
static const double kGain = 1.85; /* rrc_fir.h:14, a double literal */

/* One tap before normalisation.  Sub-expressions that touch M_PI are double in the reference
 * (usual arithmetic conversions), everything else is float; each line cites what it follows. */
static float rrc_raw_tap(int i, int ntaps, float spb, float alpha, int *is_unit) {
    const float k = (float)(i - ntaps / 2);                               /* rrc_fir.c:39 */
    const float x1 = (float)(M_PI * (double)k / (double)spb);             /* :40 */
    const float x2 = 4.f * alpha * k / spb;                               /* :41 */
    const float x3 = x2 * x2 - 1.f;                                       /* :42 */
    float num, den;
    *is_unit = 0;
    if (fabsf(x3) >= 0.000001f) {                                         /* :44 */
        const float lead = cosf((1.f + alpha) * x1);
        if (i != ntaps / 2)
            num = lead + sinf((1.f - alpha) * x1) / (4.f * alpha * k / spb);                    /* :46-47 */
        else
            num = (float)((double)lead + (double)(1.f - alpha) * M_PI / (double)(4.f * alpha)); /* :49 */
        den = (float)((double)x3 * M_PI);                                 /* :51 */
    } else {
        if (alpha == 1.f) { *is_unit = 1; return -1.f; }                  /* :53-57 */
        const float a3 = (1.f - alpha) * x1, a2 = (1.f + alpha) * x1;     /* :59-60 */
        const double p = (double)(sinf(a2) * (1.f + alpha)) * M_PI;       /* :62 */
        const double q = (double)cosf(a3) * ((double)(1.f - alpha) * M_PI * (double)spb)
                         / (double)(4.f * alpha * k);                      /* :63 */
        const float r = sinf(a3) * spb * spb / (4.f * alpha * k * k);     /* :64 */
        num = (float)(p - q + (double)r);
        den = (float)((double)-32.f * M_PI * (double)alpha * (double)alpha * (double)k / (double)spb); /* :66 */
    }
    return 4.f * alpha * num / den;                                       /* :69 */
}

void qpsk_host_rrc_make(float *taps, int ntaps, float fs, float rs, float alpha) {
    const float spb = fs / rs;                                            /* :34 */
    float sum = 0.f;
    for (int i = 0; i < ntaps; i++) {
        int unit;
        taps[i] = rrc_raw_tap(i, ntaps, spb, alpha, &unit);
        sum += taps[i];                                                   /* :70 (and :55) */
    }
    for (int i = 0; i < ntaps; i++)
        taps[i] = (float)(((double)taps[i] * kGain) / (double)sum);       /* :73-75 */
}

void qpsk_host_loop_update_gains(qpsk_host_loop *l) {                     /* costas_loop.c:49-54 */
    const float denom = (1.0f + (2.0f * l->damping * l->loop_bw)) + (l->loop_bw * l->loop_bw);
    l->alpha = (4.0f * l->damping * l->loop_bw) / denom;
    l->beta = (4.0f * l->loop_bw * l->loop_bw) / denom;
}

void qpsk_host_loop_create(qpsk_host_loop *l, float loop_bw, float min_freq, float max_freq) {
    memset(l, 0, sizeof *l);               /* file-scope statics start at zero, costas_loop.c:13-23 */
    l->max_freq = max_freq;                /* :35-36 (phase and frequency were just set to 0, :32-33) */
    l->min_freq = min_freq;
    l->damping = sqrtf(2.0f) / 2.0f;       /* :38 */
    l->loop_bw = loop_bw;                  /* :41 */
    qpsk_host_loop_update_gains(l);
}

void qpsk_host_cis(double v, int conjugate, float out[2]) {
    const float a = (float)v;
    out[0] = cosf(a);
    out[1] = conjugate ? -sinf(a) : sinf(a);
}
