// bits.cuh -- K4: the bitwise stages after the slicer, batched over frames.
//
//   scramble   bit-scramble.c:46-84   additive DVB LFSR 1 + x^14 + x^15, seed 0x4A80, 2 bits per call
//   interleave interleave.c:43-78     j = (b*i) mod nbits, b = largest table prime < nbits, LSB-first bits
//   crc16      crc16.c:11-23          CRC-16/CCITT-FALSE, nibble-folded byte update
//
// Two forms: generic primitives over row-major frames [nframes][nbytes] (any nbytes < 8192,
// mirroring the three reference functions one to one), and the receiver-attached fused decode /
// encode (descramble -> de-interleave -> CRC check) that reads the slicer's packed dibits in the
// channel-fastest layout the Costas kernel writes, one thread per (frame, channel), with the
// keystream and the bit permutation resolved at compile time.
#pragma once

#include "common.cuh"

#define QPSK_SCRAMBLE_SEED 0x4A80u   // bit-scramble.h:13

__host__ __device__ inline uint16_t crc16_update(uint16_t crc, uint8_t byte) {   // crc16.c:15-20
    uint8_t x = (uint8_t)((crc >> 8) ^ byte);
    x ^= (uint8_t)(x >> 4);
    return (uint16_t)((crc << 8) ^ ((uint16_t)(x << 12)) ^ ((uint16_t)(x << 5)) ^ ((uint16_t)x));
}

// one LFSR step: returns the keystream bit and advances the register (bit-scramble.c:59-67)
__host__ __device__ inline unsigned lfsr_step(uint16_t& reg) {
    const unsigned key = ((reg >> 1) ^ reg) & 1u;
    reg = (uint16_t)((reg >> 1) | (key << 14));
    return key;
}

static const uint16_t kInterleavePrimes[] = {   // interleave.c:33-41
    2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97,
    101, 103, 107, 109, 113, 127, 131, 137, 139, 149, 151, 157, 163, 167, 173, 179, 181, 191, 193,
    197, 199, 211, 223, 227, 229, 233, 239, 241, 251, 257, 263, 269, 271, 277, 281, 283, 293, 307,
    311, 313, 317, 331, 337, 347
};

// interleave.c:48-55.  Past the end of the table the reference's search reads primes[69] (out of
// bounds) and stops either way in its -O0/-O1 builds, leaving b = 347.
static inline int interleave_prime(int nbits) {
    const int imax = (int)(sizeof kInterleavePrimes / sizeof kInterleavePrimes[0]);
    int index = 1;
    while (index < imax && kInterleavePrimes[index] < nbits) index++;
    return kInterleavePrimes[index - 1];
}
__host__ __device__ constexpr int interleave_prime_ct(int nbits) {
    constexpr int tab[] = { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97,
                            101, 103, 107, 109, 113, 127, 131, 137, 139, 149, 151, 157, 163, 167, 173, 179, 181, 191, 193,
                            197, 199, 211, 223, 227, 229, 233, 239, 241, 251, 257, 263, 269, 271, 277, 281, 283, 293, 307,
                            311, 313, 317, 331, 337, 347 };
    int index = 1;
    while (index < 69 && tab[index] < nbits) index++;
    return tab[index - 1];
}

// ---------------------------------------------------------------------------------------------
// generic primitives, row-major frames
// ---------------------------------------------------------------------------------------------
__global__ void crc16_rows_kernel(const uint8_t* __restrict__ data, int nbytes, int nframes, uint16_t* __restrict__ crc_out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    const uint8_t* row = data + (size_t)f * nbytes;
    uint16_t crc = 0xFFFF;
    for (int i = 0; i < nbytes; i++) crc = crc16_update(crc, row[i]);
    crc_out[f] = crc;
}

// one warp per frame; bits are scattered with OR (exactly the reference's `out[jbyte] |= ...`,
// which matters when b divides nbits and the map is not a bijection)
__global__ void interleave_rows_kernel(uint8_t* __restrict__ data, int nbytes, int nframes, int b, int dir) {
    extern __shared__ unsigned sm_words[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * (blockDim.x >> 5) + warp;
    const int nwords = (nbytes + 3) / 4;
    unsigned* out = sm_words + (size_t)warp * nwords;
    const unsigned nbits = (unsigned)(nbytes * 8) & 0xFFFFu;          // uint16_t nbits, interleave.c:49
    if (f < nframes) {
        for (int i = lane; i < nwords; i += 32) out[i] = 0u;
        __syncwarp();
        uint8_t* row = data + (size_t)f * nbytes;
        for (unsigned n = lane; n < nbits; n += 32) {
            unsigned i = n, j = ((unsigned)b * n) % nbits;
            if (dir == 1) { const unsigned t = j; j = i; i = t; }
            const unsigned bit = (row[i >> 3] >> (i & 7)) & 1u;
            if (bit) atomicOr(&out[j >> 5], 1u << (j & 31));
        }
        __syncwarp();
        for (int i = lane; i < nbytes; i += 32) row[i] = (uint8_t)(out[i >> 2] >> (8 * (i & 3)));
    }
}

// dibits [nframes][ndibits], one per byte (low 2 bits); the keystream restarts from SEED in every
// row (the reference resets the register per frame, bit-scramble.c:11) and is data independent
__global__ void scramble_rows_kernel(uint8_t* __restrict__ dibits, const uint8_t* __restrict__ keystream, int ndibits, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    dibits[i] = (uint8_t)(dibits[i] ^ keystream[i % ndibits]);
}

// ---------------------------------------------------------------------------------------------
// receiver-attached fused decode: packed dibits uint32 [F][W][Cpad] (W = NBYTES/4 words per frame,
// 16 dibits per word, LSB first) -> de-scrambled, de-interleaved frame words [F][W][Cpad],
// CRC verdict [F][Cpad] and pass/fail counters.
// ---------------------------------------------------------------------------------------------
// The scrambler keystream of one frame (packed like the dibits) travels in the kernel arguments: it is data
// independent, 64 bytes at most, and per launch -- nothing process-wide to upload per device or to order
// against kernels in flight.
struct Keystream {
    unsigned w[8];
};

struct FrameDecodeArgs {
    const unsigned* dibits_t;   // [F][W][Cpad]
    unsigned* frames_t;         // [F][W][Cpad] payload | crc, byte k of the frame at bits 8*(k%4) of word k/4
    uint8_t* crc_ok_t;          // [F][Cpad]
    uint8_t* rotation_t;        // [F][Cpad] or null: resolve the 90-degree ambiguity of the loop on the CRC (see below)
    unsigned long long* counters;   // [0] frames examined, [1] CRC passes
    int C, Cpad, F;
    int c0;                     // first channel of this launch; C is its end
    Keystream ks;
};

template <int NBYTES, int DIR>
__device__ __forceinline__ void permute_frame(const unsigned (&in)[NBYTES / 4], unsigned (&out)[NBYTES / 4]) {
    constexpr int NBITS = NBYTES * 8;
    constexpr int B = interleave_prime_ct(NBITS);
#pragma unroll
    for (int w = 0; w < NBYTES / 4; w++) out[w] = 0u;
#pragma unroll
    for (int n = 0; n < NBITS; n++) {
        const int m = (B * n) % NBITS;
        const int i = DIR == 0 ? n : m, j = DIR == 0 ? m : n;        // out bit j <- in bit i (interleave.c:57-75)
        out[j >> 5] |= ((in[i >> 5] >> (i & 31)) & 1u) << (j & 31);
    }
}

// A Costas loop locks on any of four phases 90 degrees apart.  With the reference's mapping (constellation
// index d <-> {1, j, -j, -1}, qpsk.c:58-63, and the slicer's dibit = b0 | b1 << 1, qpsk.c:74-79) a stream received a
// quarter turn ahead carries rho(d) = (b0' = !b1, b1' = b0) in place of d: 0 -> 1 -> 3 -> 2 -> 0.  This undoes one
// quarter turn on 16 packed dibits at once: b0 = b1', b1 = !b0'.
__device__ __forceinline__ unsigned unrotate_dibits(unsigned w) {
    const unsigned b0 = w & 0x55555555u, b1 = (w >> 1) & 0x55555555u;
    return b1 | ((~b0 & 0x55555555u) << 1);
}

template <int NBYTES>
__global__ void __launch_bounds__(128) frame_decode_kernel(const FrameDecodeArgs a) {
    constexpr int W = NBYTES / 4;
    const int c = a.c0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    unsigned ok = 0;
    if (c < a.C) {
        unsigned raw[W], in[W], out[W], first[W];
#pragma unroll
        for (int w = 0; w < W; w++) raw[w] = a.dibits_t[((size_t)f * W + w) * a.Cpad + c];
        // rotation 0 is the plain decode; with rotation_t the other three quarter turns are tried in order and
        // the first whose CRC matches is kept (no match: the rotation-0 decode is stored, rotation = 255)
        const int nrot = a.rotation_t ? 4 : 1;
        int rot = 0;
#pragma unroll 1
        for (; rot < nrot; rot++) {
#pragma unroll
            for (int w = 0; w < W; w++) in[w] = raw[w] ^ a.ks.w[w];
            permute_frame<NBYTES, 1>(in, out);
            uint16_t crc = 0xFFFF;
#pragma unroll
            for (int k = 0; k < NBYTES - 2; k++) crc = crc16_update(crc, (uint8_t)(out[k >> 2] >> (8 * (k & 3))));
            const unsigned hi = (out[(NBYTES - 2) >> 2] >> (8 * ((NBYTES - 2) & 3))) & 0xffu;
            const unsigned lo = (out[(NBYTES - 1) >> 2] >> (8 * ((NBYTES - 1) & 3))) & 0xffu;
            ok = (hi == (unsigned)(crc >> 8) && lo == (unsigned)(crc & 0xff)) ? 1u : 0u;
            if (rot == 0) {
#pragma unroll
                for (int w = 0; w < W; w++) first[w] = out[w];
            }
            if (ok) break;
#pragma unroll
            for (int w = 0; w < W; w++) raw[w] = unrotate_dibits(raw[w]);
        }
#pragma unroll
        for (int w = 0; w < W; w++) a.frames_t[((size_t)f * W + w) * a.Cpad + c] = ok ? out[w] : first[w];
        a.crc_ok_t[(size_t)f * a.Cpad + c] = (uint8_t)ok;
        if (a.rotation_t) a.rotation_t[(size_t)f * a.Cpad + c] = ok ? (uint8_t)rot : (uint8_t)255;
    }
    // one atomic per warp: ballot the verdicts
    const unsigned live = __ballot_sync(0xffffffffu, c < a.C);
    const unsigned pass = __ballot_sync(0xffffffffu, ok != 0);
    if ((threadIdx.x & 31) == 0 && live) {
        atomicAdd(&a.counters[0], (unsigned long long)__popc(live));
        atomicAdd(&a.counters[1], (unsigned long long)__popc(pass));
    }
}

// transmit side of the same format: payload words [F][W][Cpad] (last two bytes ignored) ->
// CRC appended, interleaved, scrambled, packed dibits [F][W][Cpad]
struct FrameEncodeArgs {
    const unsigned* payload_t;
    unsigned* dibits_t;
    int C, Cpad, F;
    Keystream ks;
};

template <int NBYTES>
__global__ void __launch_bounds__(128) frame_encode_kernel(const FrameEncodeArgs a) {
    constexpr int W = NBYTES / 4;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (c >= a.C) return;
    unsigned in[W], out[W];
#pragma unroll
    for (int w = 0; w < W; w++) in[w] = a.payload_t[((size_t)f * W + w) * a.Cpad + c];
    uint16_t crc = 0xFFFF;
#pragma unroll
    for (int k = 0; k < NBYTES - 2; k++) crc = crc16_update(crc, (uint8_t)(in[k >> 2] >> (8 * (k & 3))));
    // last word: two payload bytes | crc high | crc low (NBYTES is a multiple of 4)
    in[W - 1] = (in[W - 1] & 0x0000ffffu) | ((unsigned)(crc >> 8) << 16) | ((unsigned)(crc & 0xff) << 24);
    permute_frame<NBYTES, 0>(in, out);
#pragma unroll
    for (int w = 0; w < W; w++) a.dibits_t[((size_t)f * W + w) * a.Cpad + c] = out[w] ^ a.ks.w[w];
}
