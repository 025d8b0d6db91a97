// channel.cuh -- channel impairments for the synthetic-signal generator (SURVEY 8(d) "synthetic input"): additive white
// Gaussian-like noise on int16 PCM from a counter-based generator, so that the same (seed, channel, sample) gives the same
// noise sample on the GPU and in the oracle's C restatement (orc_awgn_sample) -- no transcendental functions anywhere.
//
//   bits   = Philox-4x32-10(counter = (sample >> 1, channel, 0, 0), key = (seed lo, seed hi))   (Salmon et al., SC'11)
//   u_k    = the eight 16-bit halves of the four words, samples 2m and 2m+1 share one block (words 0-1 / 2-3 -> 4 halves each)
//   g      = (sum of FOUR u_k - 131070) * (1 / 37837.0f)        Irwin-Hall(4): mean 2 * 65535, sigma = sqrt(4 (2^32 - 1) / 12)
//   pcm'   = clamp(trunc((float)pcm + sigma * g), -32768, 32767) the reference's own float -> int16 conversion (qpsk.c:260)
// Every float step is one rounded operation.  Irwin-Hall(4) has light tails (|g| <= 3.46); fine for a 20 dB test channel.
#pragma once

#include "common.cuh"

#define QPSK_PHILOX_M0 0xD2511F53u
#define QPSK_PHILOX_M1 0xCD9E8D57u
#define QPSK_PHILOX_W0 0x9E3779B9u
#define QPSK_PHILOX_W1 0xBB67AE85u

__host__ __device__ inline void qpsk_philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned (&out)[4]) {
    for (int r = 0; r < 10; r++) {
        const unsigned long long p0 = (unsigned long long)QPSK_PHILOX_M0 * c0, p1 = (unsigned long long)QPSK_PHILOX_M1 * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += QPSK_PHILOX_W0; k1 += QPSK_PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__host__ __device__ inline float qpsk_awgn_unit(unsigned w0, unsigned w1) {      // one noise sample of unit variance from two words
    const int s = (int)(w0 & 0xffffu) + (int)(w0 >> 16) + (int)(w1 & 0xffffu) + (int)(w1 >> 16);
    return (float)(s - 131070) * (1.0f / 37837.0f);
}

struct AwgnArgs {
    int16_t* pcm;              // [C][T] in place
    const float* sigma;        // [C] noise standard deviation per channel, in PCM units
    int C;
    long long T;
    long long first_sample;    // stream position of pcm[.][0] (even): a stream may be generated in pieces
    unsigned seed_lo, seed_hi;
    int chan_base;             // channel number of row 0 (so shards of one channel set draw the same noise)
};

__global__ void __launch_bounds__(256) awgn_kernel(const AwgnArgs a) {
    const long long pairs = (a.T + 1) / 2;
    const unsigned blocks_per_row = (unsigned)((pairs + 255) / 256);
    const int c = (int)(blockIdx.x / blocks_per_row);
    const long long i = (long long)(blockIdx.x % blocks_per_row) * blockDim.x + threadIdx.x;   // sample pair inside the row
    if (i >= pairs) return;
    const long long n0 = a.first_sample + 2 * i;
    unsigned w[4];
    qpsk_philox4x32_10((unsigned)((unsigned long long)n0 >> 1), (unsigned)(a.chan_base + c), (unsigned)((unsigned long long)n0 >> 33), 0u, a.seed_lo, a.seed_hi, w);
    const float sg = a.sigma[c];
    int16_t* row = a.pcm + (size_t)c * a.T;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const long long t = 2 * i + h;
        if (t < a.T) {
            const float g = qpsk_awgn_unit(w[2 * h], w[2 * h + 1]);
            float v = __fadd_rn((float)row[t], __fmul_rn(sg, g));
            v = truncf(v);
            v = fminf(fmaxf(v, -32768.0f), 32767.0f);
            row[t] = (int16_t)(int)v;
        }
    }
}
