/*
 * qpsk_b200.h -- batch C-ABI of the B200-native QPSK receiver (libqpsk_b200.so).
 *
 * Plain C: opaque handles, plain pointers and sizes, int status returns (0 = success, negative
 * = error, text from qpsk_b200_last_error()).  A context is not thread-safe and its calls must be
 * stream-ordered; contexts are independent of each other (taps, keystream and every other constant
 * travel with each kernel launch, nothing is process-wide), so several contexts -- on one device
 * or on several -- may be driven from one host thread or from one thread each.  The intended
 * deployment is one process per GPU.  No CPU fallback exists: every entry point that
 * computes fails with QPSK_B200_ERR_CUDA when no sm_100 device is usable.
 *
 * The reference (MonsieurETM/QPSK) handles exactly one channel through file-scope singletons;
 * this interface is the same pipeline over many independent channels:
 *
 *   qpsk_b200_rx_process_*   <->  rx_frame()           qpsk.c:88-218   (static in the reference)
 *        mixer                    qpsk.c:114-120
 *        matched filter           rrc_fir()            rrc_fir.c:17-30
 *        timing histogram/index   qpsk.c:131-180
 *        decimation + delay       qpsk.c:186-191
 *        Costas loop              qpsk.c:196-207, costas_loop.c:44-74
 *        slicer                   qpsk_demod()         qpsk.c:74-79
 *   qpsk_b200_fir_*          <->  rrc_fir()/rrc_make() rrc_fir.h:16-17
 *   qpsk_b200_fft_*          <->  fftn()/ifftn()       algorithms/fft.h:46-49 (+ |X|^2 argmax)
 *   qpsk_b200_bits_*         <->  scramble()/interleave()/crc16()  algorithms/{bit-scramble,interleave,crc16}.h
 *   qpsk_b200_tx_*           <->  qpsk_packet_mod()/tx_frame()     qpsk.c:225-285
 *
 * The single-channel drop-in symbols of the reference headers (rrc_fir, rrc_make, the 22
 * costas_loop functions, fft/fftn/ifft/ifftn, crc16, interleave, scramble, rx_frame, ...) are
 * exported by the same library; see include/qpsk_dropin.h and INTEGRATION.md.
 */
#ifndef QPSK_B200_H
#define QPSK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    QPSK_B200_OK = 0,
    QPSK_B200_ERR_ARG = -1,      /* bad argument / unsupported configuration */
    QPSK_B200_ERR_CUDA = -2,     /* CUDA runtime or device failure (no fallback) */
    QPSK_B200_ERR_STATE = -3     /* call sequence error (e.g. output not enabled) */
};

enum { QPSK_B200_MODE_EXACT = 0,   /* reference arithmetic, bit-exact decisions */
       QPSK_B200_MODE_FAST = 1 };  /* fused multiply-add FIR (<= 1e-5 relative), not bit-exact */
enum { QPSK_B200_UB_ALIAS = 0,     /* reproduce the Makefile-build out-of-frame read of qpsk.c:190 */
       QPSK_B200_UB_CLAMP = 1,     /* fenced: out-of-frame reads return the last sample of the frame */
       QPSK_B200_UB_PHASE = 2,     /* extension, not the reference: the timing index only picks the sampling phase
                                      (sample i*CYCLES + index % CYCLES), so no symbol is ever taken from outside the frame */
       QPSK_B200_UB_TAU = 3 };     /* extension ("sample at tau"): the histogram index is replaced by round(tau) mod CYCLES, tau the
                                      square-law timing estimate of the frame (OUT_TIMING_TAU): the sample nearest to the eye's
                                      maximum, stable from frame to frame.  OUT_INDEX then reports that phase */

enum {                              /* cfg.flags */
    QPSK_B200_KEEP_FIR = 1,         /* keep the matched-filter output (16 B/sample!) for parity checks */
    QPSK_B200_KEEP_SYMBOLS = 2,     /* keep the derotated symbols (costas_frame) */
    QPSK_B200_DECODE_FRAMES = 4,    /* run descramble -> de-interleave -> CRC16 on every frame's dibits */
    QPSK_B200_NO_FUSE = 8,          /* always run the Costas loop as its own kernel (it is fused into the front end
                                       whenever a CTA owns whole streams, i.e. when channels are plentiful) */
    QPSK_B200_SLICE_DIAGONAL = 32,  /* extension, not the reference: slice the loop output on the diagonals where phase_detector
                                       locks it, without qpsk_demod's extra 45 degrees (qpsk.c:75) that leaves bits[0] on a
                                       decision boundary; with UB_PHASE + RESOLVE_ROTATION this makes framed loop-back decodable */
    QPSK_B200_ESTIMATE_OFFSET = 64, /* run the FFT frequency estimator inside every process call: 4th power of the call's first
                                       n decimated symbols per channel (n = the largest power of two <= min(1024, symbols of
                                       the call)) -> batched n-point FFT -> |X|^2 argmax, results in HBM (OUT_OFFSET_BIN/_HZ).
                                       The reference has fftn() but never calls it: an extension output, the
                                       receive decisions do not depend on it */
    QPSK_B200_ESTIMATE_TIMING = 128,/* extension, not the reference: per frame the symbol-rate line of the squared matched-filter output,
                                       S = sum_n y_n^2 e^{-2 pi i n / CYCLES} (Oerder & Meyr), accumulated by the timing warps next to the
                                       reference's amplitude histogram; OUT_TIMING_SUM / OUT_TIMING_TAU.  The decisions do not use it */
    QPSK_B200_RESOLVE_ROTATION = 16,/* with DECODE_FRAMES: a frame whose CRC fails is retried with its dibits turned back
                                       by 90, 180 and 270 degrees (the loop's phase ambiguity); first match wins */
    QPSK_B200_PREROTATE_OFFSET = 512, /* extension: in the first call after create / reset the FFT frequency estimator (4th power of the
                                       call's first symbols -> FFT -> argmax, see ESTIMATE_OFFSET, which this flag implies) runs BEFORE the
                                       Costas loop and seeds every channel's d_freq with its estimate, set_frequency(TAU * offset_hz / RS):
                                       the loop starts locked instead of pulling in.  Later calls are unchanged */
    QPSK_B200_TRANSIENT_SYMBOLS = 1024, /* the decimated symbols are a hand-off between the timing stage and the loop, not an output:
                                       when the loop rides along in the front-end kernel, a ring slot is dropped from L2 without write-back
                                       (discard.global.L2) as soon as the loop has consumed it, which halves the kernel's HBM write traffic.
                                       QPSK_B200_OUT_DEC is then unavailable; every other output is unchanged */
    QPSK_B200_NO_CHUNK = 256        /* never cut a call into frame chunks (with few channels a long call is processed as
                                       chunks of frames so that the Costas loop of one chunk runs under the front end of
                                       the next; results are identical either way) */
};

typedef struct {
    float fs;          /* FS      qpsk.h:16   9600 */
    float rs;          /* RS      qpsk.h:17   2400 (1200 for the 10 m profile) */
    float center;      /* CENTER  qpsk.h:18   1500 */
    float rrc_alpha;   /* qpsk.c:308          0.35 */
    float loop_bw;     /* qpsk.c:302          (float)(TAU/100) */
    int   ntaps;       /* NTAPS   rrc_fir.h:13 127 */
    int   frame_size;  /* FRAME_SIZE qpsk.h:23 512 */
    int   mode;        /* QPSK_B200_MODE_* */
    int   ub_mode;     /* QPSK_B200_UB_* */
    int   flags;       /* QPSK_B200_KEEP_* */
    int   device;      /* CUDA device ordinal */
} qpsk_b200_rx_config;

typedef struct qpsk_b200_rx qpsk_b200_rx;

/* which array qpsk_b200_rx_read() downloads; all are channel-major on the host side */
enum {
    QPSK_B200_OUT_DIBITS = 0,   /* uint8  [C][F*nsym/4]  4 dibits per byte, symbol i at bits 2*(i%4); dibit = bits[0] | bits[1]<<1 */
    QPSK_B200_OUT_INDEX = 1,    /* int32  [C][F]         timing index per frame */
    QPSK_B200_OUT_TRACK = 2,    /* float  [C][F][2]      (d_phase, d_freq) after each frame */
    QPSK_B200_OUT_DEC = 3,      /* float2 [C][F*nsym]    decimated symbols produced by each frame */
    QPSK_B200_OUT_SYMBOLS = 4,  /* float2 [C][F*nsym]    costas_frame (needs KEEP_SYMBOLS) */
    QPSK_B200_OUT_FIR = 5,      /* float2 [C][F*N]       matched-filter output (needs KEEP_FIR) */
    QPSK_B200_OUT_TAPS = 6,     /* float  [ntaps] */
    QPSK_B200_OUT_FRAMES = 7,   /* uint8  [C][F*nbytes]  de-scrambled, de-interleaved frames: payload | crc16 hi | lo (needs DECODE_FRAMES) */
    QPSK_B200_OUT_CRC_OK = 8,   /* uint8  [C][F]         1 where the frame's CRC16 matched (needs DECODE_FRAMES) */
    QPSK_B200_OUT_ROTATION = 9, /* uint8  [C][F]         quarter turns undone before the CRC matched, 0..3; 255 = no match
                                                          (needs DECODE_FRAMES | RESOLVE_ROTATION) */
    QPSK_B200_OUT_OFFSET_BIN = 10, /* int32 [C]          argmax bin of the last call's 4th-power spectrum (needs ESTIMATE_OFFSET) */
    QPSK_B200_OUT_OFFSET_HZ = 11,  /* float [C]          the same as a carrier offset: signed bin * rs / (4 n) */
    QPSK_B200_OUT_TIMING_SUM = 12, /* float [C][F][2]    (Re S, Im S) per frame (needs ESTIMATE_TIMING) */
    QPSK_B200_OUT_TIMING_TAU = 13  /* float [C][F]       sampling phase of the eye in samples, -arg(S) CYCLES / (2 pi) in [0, CYCLES) */
};

const char *qpsk_b200_last_error(void);
int qpsk_b200_device_count(void);

/* fills the reference's constants (2400 baud profile) */
void qpsk_b200_rx_default_config(qpsk_b200_rx_config *cfg);

/* nchan independent channels, at most max_frames frames per process call */
int qpsk_b200_rx_create(const qpsk_b200_rx_config *cfg, int nchan, int max_frames, qpsk_b200_rx **out);
int qpsk_b200_rx_destroy(qpsk_b200_rx *rx);
/* stream start for every channel: zero delay lines, loop at rest, fbb_rx_phase = cmplx(0) */
int qpsk_b200_rx_reset(qpsk_b200_rx *rx);

/* PCM already in HBM: d_pcm is int16 [C][nframes*frame_size], 16-byte aligned.  Asynchronous on
 * `cuda_stream` (a cudaStream_t, NULL = the context's own stream).  Channel state carries over
 * to the next call, so a stream may be fed in several calls. */
int qpsk_b200_rx_process_device(qpsk_b200_rx *rx, const int16_t *d_pcm, int nframes, void *cuda_stream);
/* PCM in host memory (pinned for best speed): copies in, runs, copies the packed dibits out
 * (h_dibits may be NULL) and returns when done. */
int qpsk_b200_rx_process_host(qpsk_b200_rx *rx, const int16_t *h_pcm, int nframes, uint8_t *h_dibits);
/* The same call split in two, for a continuous receiver (the reference's read loop, qpsk.c:339-354, reads the next
 * 512 samples only after rx_frame returned): submit enqueues the copies and kernels of one batch and returns at once,
 * wait blocks until the OLDEST submitted batch is complete (its dibits are in h_dibits).  Up to two batches may be in
 * flight, so the host can read batch k+1 from its file or socket while the GPU works on batch k.  h_pcm and h_dibits
 * must stay valid (and should be page-locked) until the matching wait.  After an error the context needs
 * qpsk_b200_rx_reset(): some channels may have advanced and others not. */
int qpsk_b200_rx_submit_host(qpsk_b200_rx *rx, const int16_t *h_pcm, int nframes, uint8_t *h_dibits);
int qpsk_b200_rx_wait(qpsk_b200_rx *rx);
/* measurement aid: the host<->device copies of qpsk_b200_rx_process_host (same slices, streams and events) with no
 * kernel launched -- the copy ceiling an end-to-end rate is quoted against.  Channel state is not touched. */
int qpsk_b200_rx_probe_copy_host(qpsk_b200_rx *rx, const int16_t *h_pcm, int nframes, uint8_t *h_scratch_out);
int qpsk_b200_rx_sync(qpsk_b200_rx *rx);

/* page-locked host memory for the buffers of the host entry points (cudaHostAlloc / cudaFreeHost) */
int qpsk_b200_host_alloc(size_t bytes, void **out);
int qpsk_b200_host_free(void *p);

/* Continuous receiver over raw s16le PCM files, one file per channel -- the reference's on-disk format (TX_FILENAME,
 * qpsk.h:14) and read loop (qpsk.c:339-354: 512 samples per fread, stop at the first short read), for every channel of
 * `rx` at once.  Batches of frames_per_batch frames (<= the receiver's max_frames) are read by a few host threads
 * into page-locked buffers while the GPU works on the previous batch; `sink` receives every completed batch in order:
 * packed dibits uint8 [C][nframes * nsym / 4] (valid until it returns; return non-zero to stop).  Files are cut to the
 * shortest one; a tail that does not fill a frame is ignored.  frame_size and nsym are the receiver's (512 and
 * 512 / CYCLES). */
typedef struct qpsk_b200_stream qpsk_b200_stream;
typedef int (*qpsk_b200_stream_sink)(void *user, long long first_frame, int nframes, const uint8_t *dibits);
typedef struct {
    long long frames;        /* frames per channel processed */
    double seconds;          /* wall time of the run */
    double read_seconds;     /* of which: reading the files */
    double wait_seconds;     /* of which: waiting for the GPU */
    int readers;             /* reader threads */
} qpsk_b200_stream_stats;
int qpsk_b200_stream_open(qpsk_b200_rx *rx, const char *const *paths, int nchan, int frames_per_batch, int frame_size, int nsym,
                          qpsk_b200_stream **out);
long long qpsk_b200_stream_frames(const qpsk_b200_stream *st);
int qpsk_b200_stream_run(qpsk_b200_stream *st, qpsk_b200_stream_sink sink, void *user, qpsk_b200_stream_stats *stats);
int qpsk_b200_stream_close(qpsk_b200_stream *st);

/* the loop's gains and limits (set_alpha/set_beta/set_min_freq/set_max_freq of costas_loop.h:26-32), all channels */
int qpsk_b200_rx_set_loop(qpsk_b200_rx *rx, float alpha, float beta, float min_freq, float max_freq);
/* per-channel (d_phase, d_freq) pairs, float [C][2] (set_phase/set_frequency/get_phase/get_frequency) */
int qpsk_b200_rx_get_loop_state(qpsk_b200_rx *rx, float *h_phase_freq);
int qpsk_b200_rx_set_loop_state(qpsk_b200_rx *rx, const float *h_phase_freq);

/* download one output of the most recent process call, channel-major, `bytes` = exact size */
int qpsk_b200_rx_read(qpsk_b200_rx *rx, int what, void *h_dst, size_t bytes);
size_t qpsk_b200_rx_output_bytes(const qpsk_b200_rx *rx, int what);
/* device-resident packed dibits of the most recent call, internal layout
 * uint32 [F][nsym/16][Cpad] (channel-fastest); for consumers that stay on the GPU */
int qpsk_b200_rx_device_dibits(qpsk_b200_rx *rx, const uint32_t **d_ptr, int *cpad);
/* kernels launched by this context so far (bench bookkeeping) */
long long qpsk_b200_rx_launch_count(const qpsk_b200_rx *rx);
/* device milliseconds of the front-end kernel in the most recent call (CUDA events on its stream) */
int qpsk_b200_rx_last_kernel_ms(qpsk_b200_rx *rx, float *front_ms, float *costas_ms);
/* the launch plan the most recent device-resident call (or the last job of a host call) ran under: frame chunks of the call,
 * frame blocks per channel group of its last front-end launch, and where the Costas loop (qpsk.c:196-212) ran */
#define QPSK_B200_LOOP_STANDALONE 0   /* costas_kernel behind the front end (per call or per chunk) */
#define QPSK_B200_LOOP_FUSED      1   /* in the spare warp of whole-stream front-end CTAs */
#define QPSK_B200_LOOP_RELAYED    2   /* in the front-end CTAs of frame blocks, state relayed from block to block */
#define QPSK_B200_LOOP_CHASING    3   /* one kernel on SMs of its own that chases the frame chunks */
#define QPSK_B200_LOOP_FOLLOWING  4   /* one-warp CTAs beside a frame-blocked front end (QPSK_B200_FOLLOW=1) */
int qpsk_b200_rx_last_plan(const qpsk_b200_rx *rx, int *frame_chunks, int *frame_blocks, int *loop_mode);
/* the launch policy alone (no device touched): the plan of a device-resident call of nchan channels x nframes frames at symbol
 * rate rs (2400 or 1200) on a GPU of nsm SMs, default configuration; sm_clock_khz <= 0 = 1,965,000 */
int qpsk_b200_debug_plan(int nsm, int sm_clock_khz, int nchan, int nframes, double rs, int *frame_chunks, int *frame_blocks, int *loop_mode);

/* ------------------------------------------------------------------------------------------
 * Channel-batched rrc_fir()/rrc_make()  (rrc_fir.h:16-17, rrc_fir.c:17-76)
 * ---------------------------------------------------------------------------------------- */
typedef struct qpsk_b200_fir qpsk_b200_fir;

/* rrc_make(fs, rs, alpha) for any NTAPS: taps[ntaps] on the host, bit-identical to the reference */
int qpsk_b200_rrc_make(float *taps, int ntaps, float fs, float rs, float alpha);
/* ntaps in {127, 256}; taps[ntaps] as produced by qpsk_b200_rrc_make (any real taps work) */
int qpsk_b200_fir_create(const float *taps, int ntaps, int nchan, int mode, int device, qpsk_b200_fir **out);
int qpsk_b200_fir_destroy(qpsk_b200_fir *f);
/* zero every channel's delay line (the reference's `memory[]` starts as zero-initialised globals) */
int qpsk_b200_fir_reset(qpsk_b200_fir *f);
/* rrc_fir(memory, sample, length) for every channel: d_samples is complex float [C][nsamples]
 * (re,im interleaved) in HBM, filtered in place; the delay lines persist between calls */
int qpsk_b200_fir_process_device(qpsk_b200_fir *f, float *d_samples, int nsamples, void *cuda_stream);
int qpsk_b200_fir_process_host(qpsk_b200_fir *f, float *h_samples, int nsamples);
/* read / write the delay lines, complex float [C][ntaps], oldest input first (== rrc_fir's memory[]) */
int qpsk_b200_fir_get_memory(qpsk_b200_fir *f, float *h_memory);
int qpsk_b200_fir_set_memory(qpsk_b200_fir *f, const float *h_memory);
int qpsk_b200_fir_last_kernel_ms(qpsk_b200_fir *f, float *ms);

/* ------------------------------------------------------------------------------------------
 * Batched FFT + |X|^2 argmax  (algorithms/fft.h:46-49; fft.c:98-136)
 * FP32 on the GPU (the reference is complex double): agrees within 1e-5 max-norm relative.
 * Forward is scaled by 1/n and the inverse is unscaled, as in the reference.
 * ---------------------------------------------------------------------------------------- */
typedef struct qpsk_b200_fft qpsk_b200_fft;

/* n: a power of two, 2..8192 */
int qpsk_b200_fft_create(int n, int device, qpsk_b200_fft **out);
int qpsk_b200_fft_destroy(qpsk_b200_fft *f);
/* estimator: d_in complex float [nbursts][n] in HBM -> d_bin int32 [nbursts] (first strict maximum
 * of |X[k]|^2, k = 0..n-1) and d_mag2 float [nbursts] (may be NULL).  Forward transform. */
int qpsk_b200_fft_argmax_device(qpsk_b200_fft *f, const float *d_in, int nbursts, int32_t *d_bin, float *d_mag2, void *cuda_stream);
int qpsk_b200_fft_argmax_host(qpsk_b200_fft *f, const float *h_in, int nbursts, int32_t *h_bin, float *h_mag2);
/* fftn (inverse = 0) / ifftn (inverse = 1) over nbursts transforms; d_out may equal d_in */
int qpsk_b200_fft_transform_device(qpsk_b200_fft *f, const float *d_in, float *d_out, int nbursts, int inverse, void *cuda_stream);
int qpsk_b200_fft_transform_host(qpsk_b200_fft *f, const float *h_in, float *h_out, int nbursts, int inverse);
int qpsk_b200_fft_last_kernel_ms(qpsk_b200_fft *f, float *ms);
/* one transform longer than a CTA holds, n a power of two in 16384 .. 2^26 (the reference's fftn / ifftn take any power of
 * two, fft.c:110-136): four-step decomposition over the batched kernel; h_out may equal h_in */
int qpsk_b200_fft_big_host(const float *h_in, float *h_out, int n, int inverse, int device);

/* ------------------------------------------------------------------------------------------
 * Bit stages  (algorithms/bit-scramble.h:29-30, interleave.h:13, crc16.h:10), batched over frames.
 * Host-buffer entry points; rows are frames.  The *_rx_* form runs on the slicer output in HBM.
 *
 * Frame format used by DECODE_FRAMES / qpsk_b200_frames_encode (this project's composition; the
 * reference never chains its bit stages): frame = payload[nbytes-2] | crc16(payload) hi | lo ->
 * interleave(INTERLEAVE) -> dibits LSB first -> scramble with the register reset to SEED per
 * frame; nbytes = FRAME_SIZE/CYCLES/4 (32 at 2400 baud, 16 at 1200 baud), one frame per rx_frame.
 * ---------------------------------------------------------------------------------------- */
/* crc16() of every row: h_data uint8 [nframes][nbytes] -> h_crc uint16 [nframes] */
int qpsk_b200_bits_crc16(const uint8_t *h_data, int nbytes, int nframes, uint16_t *h_crc, int device);
/* interleave(row, nbytes, dir) of every row in place; dir 0 = INTERLEAVE, 1 = DEINTERLEAVE; nbytes < 8192 */
int qpsk_b200_bits_interleave(uint8_t *h_data, int nbytes, int nframes, int dir, int device);
/* scramble() over every row of dibits (one dibit per byte, low two bits), register = SEED at the start of each row */
int qpsk_b200_bits_scramble(uint8_t *h_dibits, int ndibits, int nframes, int device);
/* payload uint8 [C][F][nbytes] (last two bytes of each frame ignored) -> packed dibits uint8 [C][F*nbytes]
 * in the layout QPSK_B200_OUT_DIBITS uses; nbytes in {16, 32} */
int qpsk_b200_frames_encode(const uint8_t *h_payload, int nbytes, int nchan, int nframes, uint8_t *h_dibits, int device);
/* inverse of the above on host buffers: packed dibits -> frames uint8 [C][F][nbytes] and CRC verdicts uint8 [C][F] */
int qpsk_b200_frames_decode(const uint8_t *h_dibits, int nbytes, int nchan, int nframes, uint8_t *h_frames, uint8_t *h_crc_ok, int device);
/* the same with the loop's 90-degree ambiguity resolved on the CRC: constellation index d <-> {1, j, -j, -1}
 * (qpsk.c:58-63), so a stream received r quarter turns ahead carries rho^r(d), rho = 0->1->3->2->0.  Rotations
 * 0..3 are tried in order; h_rotation uint8 [C][F] receives the first that matched or 255, in which case the
 * frame stored is the rotation-0 decode. */
int qpsk_b200_frames_decode_rotated(const uint8_t *h_dibits, int nbytes, int nchan, int nframes, uint8_t *h_frames,
                                    uint8_t *h_crc_ok, uint8_t *h_rotation, int device);
/* counters accumulated by DECODE_FRAMES since create/reset: frames examined and CRC passes */
int qpsk_b200_rx_crc_counters(qpsk_b200_rx *rx, unsigned long long *frames, unsigned long long *passes);

/* ------------------------------------------------------------------------------------------
 * Batched transmit path  (qpsk_packet_mod / tx_frame / qpsk_mod, qpsk.c:58-63, 225-285)
 * Bit-exact PCM.  Also the on-device synthetic-signal generator for the receiver benchmarks.
 * ---------------------------------------------------------------------------------------- */
typedef struct qpsk_b200_tx qpsk_b200_tx;

/* carrier_hz[nchan]: each channel's fbb_tx_rect = cmplx(TAU * carrier / FS) (qpsk.c:320 uses CENTER + 50).
 * packet_symbols: symbols per qpsk_packet_mod call (qpsk.c:329 uses FRAME_SIZE/2 = 256); the up-mix
 * phasor is renormalised at every packet end (qpsk.c:253).  rs selects 2400 (sps 4) or 1200 (sps 8). */
int qpsk_b200_tx_create(float fs, float rs, float rrc_alpha, const float *carrier_hz, int nchan, int packet_symbols,
                        int device, qpsk_b200_tx **out);
int qpsk_b200_tx_destroy(qpsk_b200_tx *tx);
int qpsk_b200_tx_reset(qpsk_b200_tx *tx);
/* symbols: uint8 [C][nsym], constellation index per symbol = (tx_bits[2k] << 1) | tx_bits[2k+1] (qpsk.c:270,278-279);
 * pcm: int16 [C][nsym*sps], 8-byte aligned.  Any nsym >= 1, as tx_frame takes any length (qpsk.c:225-264).  Device pointers,
 * asynchronous. */
int qpsk_b200_tx_process_device(qpsk_b200_tx *tx, const uint8_t *d_symbols, int nsym, int16_t *d_pcm, void *cuda_stream);
int qpsk_b200_tx_process_host(qpsk_b200_tx *tx, const uint8_t *h_symbols, int nsym, int16_t *h_pcm);
/* new carriers from the next call on: fbb_tx_rect = cmplx(TAU * carrier / FS) again (qpsk.c:320), the up-mix phasor keeps
 * its phase -- a Doppler ramp is a sequence of calls with stepped carriers */
int qpsk_b200_tx_set_carrier(qpsk_b200_tx *tx, const float *carrier_hz);
/* Test-channel noise on int16 PCM in HBM, in place: d_pcm int16 [nchan][nsamples], h_sigma float [nchan] (PCM units).
 * Counter-based (Philox-4x32-10 keyed by seed, counter = sample pair and first_channel + row) Irwin-Hall(4) noise with
 * no transcendental functions, so the oracle reproduces every sample (csrc/channel.cuh); first_sample (even) is the
 * stream position of column 0.  Extension: the reference has no channel model. */
int qpsk_b200_channel_awgn_device(int16_t *d_pcm, int nchan, long long nsamples, const float *h_sigma, unsigned long long seed,
                                  long long first_sample, int first_channel, int device, void *cuda_stream);
/* end the current tx_frame call now: renormalise fbb_tx_phase (qpsk.c:253) and restart the packet position */
int qpsk_b200_tx_end_packet(qpsk_b200_tx *tx);
/* tx_frame(samples, symbol, length) itself: arbitrary complex symbols, float [C][nsym][2] in host memory */
int qpsk_b200_tx_symbols_host(qpsk_b200_tx *tx, const float *h_symbols, int nsym, int16_t *h_pcm);

/* Extension, not part of the reference's pipeline (the reference has no frequency estimator): the
 * carrier offset of every channel from the first 2^log2n decimated symbols of the most recent process call --
 * 4th power (strips the QPSK modulation), batched FFT, |X|^2 argmax.  Resolution rs / (4 * 2^log2n) Hz, range
 * +-rs/8.  h_bin (may be NULL) receives the raw argmax bins. */
int qpsk_b200_rx_estimate_offset(qpsk_b200_rx *rx, int log2n, float *h_offset_hz, int32_t *h_bin);

/* measurement aid: the FP32 pipe's own ceiling on `device`, in complex tap-updates per second machine-wide (one rounded
 * multiply + one rounded add per component, rrc_fir.c:22-26), for the exact (FMUL2 + FADD2, fused = 0) or the fast
 * (FFMA2, fused = 1) formulation of the filters -- the denominator of the FP32 roofline bench.py reports */
int qpsk_b200_probe_fp32(int device, int fused, double *tap_updates_per_s, float *kernel_ms);

/* test hook: the device NCO (the restated glibc sinf/cosf the Costas kernel uses) evaluated over n host floats */
int qpsk_b200_debug_nco(const float *h_in, float *h_sin, float *h_cos, int n, int device);

#ifdef __cplusplus
}
#endif
#endif
