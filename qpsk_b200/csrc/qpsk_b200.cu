// qpsk_b200.cu -- C-ABI (include/qpsk_b200.h) over the sm_100a receiver kernels.
// Build: see qpsk_b200/Makefile (nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <new>

#include "../../include/qpsk_b200.h"
#include "../host/host_design.h"
#include "common.cuh"
#include "rx_front.cuh"
#include "rx_front2.cuh"
#include "channel.cuh"
#include "rx_costas.cuh"
#include "fir.cuh"
#include "fft.cuh"
#include "bits.cuh"
#include "tx.cuh"
#include "probe.cuh"

// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail(QPSK_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),  \
                        __FILE__, __LINE__);                                                          \
    } while (0)

extern "C" const char* qpsk_b200_last_error(void) { return g_err; }

// for the C host files of the library (host/stream.c): one error text per thread, whoever set it
extern "C" int qpsk_b200_stream_set_error(const char* text) {
    snprintf(g_err, sizeof g_err, "%s", text ? text : "");
    return 0;
}

extern "C" int qpsk_b200_host_alloc(size_t bytes, void** out) {
    if (!out || bytes == 0) return fail(QPSK_B200_ERR_ARG, "bad argument");
    *out = nullptr;
    CU(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_host_free(void* p) {
    if (p) CU(cudaFreeHost(p));
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

struct qpsk_b200_fft;

static const double kTau = 2.0 * 3.14159265358979323846;

static long long g_next_id = 0;
struct DevBuf {      // scoped device allocation for the one-shot entry points
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
};
#define QPSK_MAX_CHUNKS 32      // frame chunks per call (rx_plan_chunks)
#define QPSK_FOLLOW_WARPS_PER_SM 4   // costas_follow_kernel: one-warp CTAs of 64 registers that fit beside two front-end CTAs (9,216 free registers / 2,048)

// the taps of a context as the kernels take them: a __grid_constant__ parameter (see TapBank, rx_front.cuh)
template <int NTAPS>
static TapBank<NTAPS> tap_bank(const float* taps) {
    TapBank<NTAPS> tb;
    for (int i = 0; i < NTAPS; i++) tb.t[i] = make_float2(taps[i], taps[i]);
    tb.one = make_float2(1.0f, 1.0f);
    return tb;
}

// ---------------------------------------------------------------------------------------------
struct qpsk_b200_rx {
    qpsk_b200_rx_config cfg;
    long long id;
    int C, Cpad, maxF, N, sps, nsym, nslots, slot_base, lastF;
    float taps[QPSK_MAX_TAPS];
    float2 rect, rot45;
    qpsk_host_loop loop;
    cudaStream_t stream;
    cudaEvent_t ev_fr[2 * QPSK_MAX_CHUNKS], ev_lp[2 * QPSK_MAX_CHUNKS];   // around K1 / K3 of every frame chunk of the last call
    int timed_chunks, timed_loop_chunks;  bool timed_loop;
    bool timed, last_fused, no_fuse;
    int dephase_cycles;     // see rx_front_kernel
    bool front_v1;          // rx_front_kernel (default) or, with QPSK_B200_FRONT=2 in the environment, rx_front2_kernel
    long long launches;
    // device state
    int16_t* d_pcm_tail;    // [Cpad][128]
    float2* d_phasor2[2];   // two tables [128 + maxF*N] (current call / next call, evaluated ahead on s_k0)
    float2* d_ph_state2;    // [2] mixer phasor after the frames of each table
    int ph_cur, ph_cur_frames;          // slot and frame count of the table the last call used
    bool ph_spec_valid; int ph_spec_frames;   // a table for a next call of ph_spec_frames frames is in flight in the other slot
    cudaStream_t s_k0;
    cudaEvent_t ev_k0_done, ev_call_start;
    float2* d_dec_ring;     // [nslots][nsym][Cpad]
    int* d_index_t;         // [maxF][Cpad]
    float2* d_loop_state;   // [Cpad]
    unsigned* d_dibits_t;   // [maxF][nsym/16][Cpad]
    float2* d_track_t;      // [maxF][Cpad]
    float2* d_fir_dbg;      // [C][maxF*N] or null
    float2* d_costas_dbg;   // [maxF][nsym][Cpad] or null
    unsigned* d_frames_t;   // [maxF][W][Cpad] decoded frames (DECODE_FRAMES)
    uint8_t* d_crc_ok_t;    // [maxF][Cpad]
    uint8_t* d_rotation_t;  // [maxF][Cpad] (RESOLVE_ROTATION)
    float2* d_timing_t;     // [maxF][Cpad] (ESTIMATE_TIMING)
    bool est_on;            // ESTIMATE_OFFSET: the estimator runs inside every process call
    float2* d_est_bursts;   // [C][est_call_n] 4th-power bursts
    int* d_est_bins;        // [C]
    float* d_est_mag;       // [C]
    int est_call_n;         // burst length of the last call
    unsigned long long* d_counters;   // [2]
    int16_t* d_pcm_stage2[2];   // host path: double-buffered PCM slices (lazy)
    unsigned* d_out_stage2[2];  // host path: double-buffered transposed dibit slices
    size_t stage_pcm_bytes, stage_out_bytes;
    cudaStream_t s_in, s_out;
    cudaEvent_t ev_in[2], ev_cmp[2], ev_res[2], ev_out[2], ev_done[2];
    unsigned long long job_seq, call_seq;   // host path: jobs and calls enqueued so far (staging buffer = seq & 1)
    int inflight;                           // submitted host calls not yet waited for (at most 2)
    bool needs_reset;                       // a host call failed half way
    bool no_chunk;                          // QPSK_B200_NO_CHUNK
    bool prerotate, loop_seeded;            // QPSK_B200_PREROTATE_OFFSET; the first call after a reset has seeded the loop
    cudaStream_t s_loop;                    // the loop of frame chunk k runs here, under the front end of chunk k+1
    int* d_chunk_flags;                     // [QPSK_MAX_CHUNKS] ticket of the last call whose chunk k has left the front end (costas_chase_kernel)
    int chase_ticket;                       // tickets handed out so far
    unsigned long long* d_block_progress;   // [front-end CTAs] progress words of a followed call (costas_follow_kernel)
    size_t block_progress_n;
    bool follow_now;                        // rx_launch_front: publish progress words for this launch
    int follow_fblocks;                     // ... and split the frames into this many blocks (decided by rx_run_call)
    int follow_mode;                        // QPSK_B200_FOLLOW in the environment: 0 = never (default: measured slower, profiles/r02_notes.md), 1 = when the cost model says so
    int follow_fb_forced;                   // QPSK_B200_FOLLOW_FB=n: follow every eligible call with n frame blocks (tests, sweeps)
    int host_slice_halfwaves;               // QPSK_B200_HOST_SLICE_HALFWAVES: channel slice of a frame-chunked host call, in half waves of front-end CTAs (default 1 = one CTA per SM)
    int host_tail_chunks;                   // QPSK_B200_HOST_CHUNKS=n: frame chunks per multi-slice host call (default 4; 1 = whole calls per slice)
    int chunk_div;                          // a chunked call is cut into this many frame chunks of at least 8 frames (QPSK_B200_CHUNK_DIV, <= QPSK_MAX_CHUNKS)
    bool no_est_emit;                       // QPSK_B200_NO_EST_EMIT=1: always the separate pass (A/B measurements)
    bool est_emit; int est_emit_n;          // this call's loop launches leave the estimator's 4th-power bursts behind (costas_emit_power4)
    int plan_chunks, plan_fblocks, plan_loop;   // qpsk_b200_rx_last_plan
    int relay_mode;                         // QPSK_B200_RELAY in the environment: 0 = never, 1 = when the cost model says so (default), n > 1 = n frame blocks whenever legal
    int chase_smem;                         // dynamic shared memory a chasing loop CTA asks for and never touches: keeps front-end CTAs off its SM
    bool no_chase;                          // QPSK_B200_NO_CHASE=1 in the environment: one loop kernel per chunk, as in round 2's first sessions
    cudaEvent_t ev_front;
    int nsm, sm_clock_khz;                  // launch policy inputs, read from the device
    size_t l2_persist_bytes, l2_window_bytes;   // persisting-L2 carve-out set aside for the frame scratch, largest access-policy window
    float* d_front_scratch; // front-end frame scratch: nsm * QPSK_SCRATCH_SLOTS shared slots, then one private region per CTA (grow-only)
    size_t front_scratch_bytes;
    int* d_scratch_slots;   // [nsm * QPSK_SCRATCH_SLOTS] claim flags of the shared slots
    qpsk_b200_fft* est_fft; int est_fft_n;   // estimator extension (lazy)
    void* d_scratch;        // transposed download staging (lazy)
    size_t scratch_bytes;
};

extern "C" void qpsk_b200_rx_default_config(qpsk_b200_rx_config* cfg) {
    memset(cfg, 0, sizeof *cfg);
    cfg->fs = 9600.0f;                        // qpsk.h:16
    cfg->rs = 2400.0f;                        // qpsk.h:17
    cfg->center = 1500.0f;                    // qpsk.h:18
    cfg->rrc_alpha = .35f;                    // qpsk.c:308
    cfg->loop_bw = (float)(kTau / 100.0f);    // qpsk.c:302
    cfg->ntaps = 127;                         // rrc_fir.h:13
    cfg->frame_size = 512;                    // qpsk.h:23
    cfg->mode = QPSK_B200_MODE_EXACT;
    cfg->ub_mode = QPSK_B200_UB_ALIAS;
}

__global__ void save_pcm_tail_kernel(const int16_t* __restrict__ pcm, int16_t* __restrict__ tail, int C, size_t row_stride, size_t row_len) {
    // last 128 of the row_len samples at the start of every channel row; 16 threads x 16 bytes per channel
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = t >> 4, q = t & 15;
    if (c >= C) return;
    const uint4* src = reinterpret_cast<const uint4*>(pcm + (size_t)c * row_stride + row_len - QPSK_CHUNK) + q;
    reinterpret_cast<uint4*>(tail + (size_t)c * QPSK_CHUNK)[q] = *src;
}

extern "C" int qpsk_b200_fft_destroy(qpsk_b200_fft* f);

static int rx_free(qpsk_b200_rx* rx) {
    if (!rx) return 0;
    cudaSetDevice(rx->cfg.device);
    if (rx->est_fft) qpsk_b200_fft_destroy(rx->est_fft);
    void* ptrs[] = { rx->d_pcm_tail, rx->d_phasor2[0], rx->d_phasor2[1], rx->d_ph_state2, rx->d_dec_ring, rx->d_index_t,
                     rx->d_loop_state, rx->d_dibits_t, rx->d_track_t, rx->d_fir_dbg, rx->d_costas_dbg,
                     rx->d_frames_t, rx->d_crc_ok_t, rx->d_rotation_t, rx->d_counters, rx->d_est_bursts, rx->d_est_bins, rx->d_est_mag, rx->d_timing_t,
                     rx->d_pcm_stage2[0], rx->d_pcm_stage2[1], rx->d_out_stage2[0], rx->d_out_stage2[1], rx->d_scratch, rx->d_front_scratch,
                     rx->d_scratch_slots, rx->d_chunk_flags, rx->d_block_progress };
    for (void* p : ptrs) if (p) cudaFree(p);
    for (auto& e : rx->ev_fr) if (e) cudaEventDestroy(e);
    for (auto& e : rx->ev_lp) if (e) cudaEventDestroy(e);
    for (int b = 0; b < 2; b++) {
        if (rx->ev_in[b]) cudaEventDestroy(rx->ev_in[b]);
        if (rx->ev_cmp[b]) cudaEventDestroy(rx->ev_cmp[b]);
        if (rx->ev_res[b]) cudaEventDestroy(rx->ev_res[b]);
        if (rx->ev_done[b]) cudaEventDestroy(rx->ev_done[b]);
        if (rx->ev_out[b]) cudaEventDestroy(rx->ev_out[b]);
    }
    if (rx->ev_k0_done) cudaEventDestroy(rx->ev_k0_done);
    if (rx->ev_call_start) cudaEventDestroy(rx->ev_call_start);
    if (rx->ev_front) cudaEventDestroy(rx->ev_front);
    if (rx->s_loop) cudaStreamDestroy(rx->s_loop);
    if (rx->s_k0) cudaStreamDestroy(rx->s_k0);
    if (rx->s_in) cudaStreamDestroy(rx->s_in);
    if (rx->s_out) cudaStreamDestroy(rx->s_out);
    if (rx->stream) cudaStreamDestroy(rx->stream);
    delete rx;
    return 0;
}

extern "C" int qpsk_b200_rx_destroy(qpsk_b200_rx* rx) { return rx_free(rx); }

template <int NTAPS, int SPS, int MODE>
static cudaError_t launch_front(const RxFrontArgs& a, const float* taps, int grid, cudaStream_t s, size_t persist_bytes, size_t persist_window) {
    const size_t smem = sizeof(RxFrontSmem<SPS>);
    cudaError_t e = cudaFuncSetAttribute(rx_front_kernel<NTAPS, SPS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(rx_front_kernel<NTAPS, SPS, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    // The per-CTA frame scratch is rewritten every frame and re-read within the frame: its lines are marked persisting in L2
    // for this launch (an access-policy window as a launch attribute: nothing is changed on the caller's stream), so that
    // the PCM streaming through does not send them on a round trip to HBM.
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(QPSK_FRONT_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    int nattr = 0;
    if (persist_bytes > 0 && a.scratch != nullptr) {
        const size_t live = (size_t)a.scratch_nslots * QPSK_SCRATCH_REGION_FLOATS * sizeof(float);      // the shared slots, not the fallback regions
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = a.scratch;
        attr[0].val.accessPolicyWindow.num_bytes = live < persist_window ? live : persist_window;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        nattr = 1;
    }
    cfg.attrs = attr; cfg.numAttrs = nattr;
    return cudaLaunchKernelEx(&cfg, rx_front_kernel<NTAPS, SPS, MODE>, a, tap_bank<NTAPS>(taps));
}

#ifdef QPSK_FRONT_PROF
extern "C" int qpsk_b200_debug_front_trace(long long* trace, unsigned long long* hdr, int* count) {
    if (cudaMemcpyFromSymbol(trace, g_front_trace, sizeof(long long) * QPSK_FRONT_TRACE_SLOTS * 256 * 10 * 2) != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(hdr, g_front_trace_hdr, sizeof(unsigned long long) * QPSK_FRONT_TRACE_SLOTS * 4) != cudaSuccess) return -1;
    return cudaMemcpyFromSymbol(count, g_front_trace_count, sizeof(int)) == cudaSuccess ? 0 : -1;
}
extern "C" int qpsk_b200_debug_front_prof(unsigned long long* dst) {
    return cudaMemcpyFromSymbol(dst, g_front_prof, sizeof(unsigned long long) * QPSK_FRONT_PROF_ROWS * 48) == cudaSuccess ? 0 : -1;
}
#endif

// the second-generation front end (rx_front2.cuh); same arguments, same outputs
template <int NTAPS, int SPS, int MODE>
static cudaError_t launch_front2(const RxFrontArgs& a, const float* taps, int grid, cudaStream_t s, size_t persist_bytes, size_t persist_window) {
    const size_t smem = sizeof(RxFront2Smem);
    cudaError_t e = cudaFuncSetAttribute(rx_front2_kernel<NTAPS, SPS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(rx_front2_kernel<NTAPS, SPS, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(QPSK_F2_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    int nattr = 0;
    if (persist_bytes > 0 && a.scratch != nullptr) {
        const size_t live = (size_t)a.scratch_nslots * QPSK_SCRATCH_REGION_FLOATS * sizeof(float);      // the shared slots, not the fallback regions
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = a.scratch;
        attr[0].val.accessPolicyWindow.num_bytes = live < persist_window ? live : persist_window;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        nattr = 1;
    }
    cfg.attrs = attr; cfg.numAttrs = nattr;
    TapBank2 tb;
    for (int i = 0; i < 127; i++) tb.t[i] = make_float2(taps[i], taps[i]);
    tb.t[127] = make_float2(0.0f, 0.0f);
    tb.one = make_float2(1.0f, 1.0f);
    return cudaLaunchKernelEx(&cfg, rx_front2_kernel<NTAPS, SPS, MODE>, a, tb);
}

extern "C" int qpsk_b200_rx_create(const qpsk_b200_rx_config* cfg, int nchan, int max_frames, qpsk_b200_rx** out) {
    if (!cfg || !out) return fail(QPSK_B200_ERR_ARG, "null argument");
    *out = nullptr;
    if (nchan < 1 || max_frames < 1) return fail(QPSK_B200_ERR_ARG, "nchan and max_frames must be positive");
    if (cfg->frame_size != 512) return fail(QPSK_B200_ERR_ARG, "frame_size %d unsupported (FRAME_SIZE is 512, qpsk.h:23)", cfg->frame_size);
    if (cfg->ntaps != 127) return fail(QPSK_B200_ERR_ARG, "receiver supports ntaps 127 (rrc_fir.h:13); got %d (use qpsk_b200_fir_* for other lengths)", cfg->ntaps);
    const int sps = (int)((double)cfg->fs / (double)cfg->rs);   // CYCLES, qpsk.h:21
    if (sps != 4 && sps != 8) return fail(QPSK_B200_ERR_ARG, "samples/symbol %d unsupported (4 = 2400 baud, 8 = 1200 baud)", sps);
    if (cfg->mode != QPSK_B200_MODE_EXACT && cfg->mode != QPSK_B200_MODE_FAST) return fail(QPSK_B200_ERR_ARG, "bad mode");
    if (cfg->ub_mode < QPSK_B200_UB_ALIAS || cfg->ub_mode > QPSK_B200_UB_TAU) return fail(QPSK_B200_ERR_ARG, "bad ub_mode %d", cfg->ub_mode);
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(QPSK_B200_ERR_CUDA, "CUDA device %d not present (%d devices)", cfg->device, ndev);
    CU(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return fail(QPSK_B200_ERR_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only", cfg->device, prop.major, prop.minor);

    qpsk_b200_rx* rx = new (std::nothrow) qpsk_b200_rx();
    if (!rx) return fail(QPSK_B200_ERR_ARG, "out of host memory");
    memset(rx, 0, sizeof *rx);
    rx->cfg = *cfg;
    rx->no_fuse = (cfg->flags & QPSK_B200_NO_FUSE) != 0;
    rx->dephase_cycles = 0;
    if (const char* dp = getenv("QPSK_B200_DEPHASE")) rx->dephase_cycles = atoi(dp);
    // QPSK_B200_FRONT=2 selects the barrier-free front end (rx_front2.cuh): same results bit for bit, within 1 % of the same speed
    // (profiles/r02_notes.md, "front-end timeline"); the default stays the kernel every round-1 and round-2 number was taken with
    rx->front_v1 = true;
    if (const char* fv = getenv("QPSK_B200_FRONT")) rx->front_v1 = atoi(fv) != 2;
    rx->no_chunk = (cfg->flags & QPSK_B200_NO_CHUNK) != 0;
    rx->nsm = prop.multiProcessorCount;
    rx->sm_clock_khz = 0;
    cudaDeviceGetAttribute(&rx->sm_clock_khz, cudaDevAttrClockRate, cfg->device);      // the device's own boost clock
    if (rx->sm_clock_khz <= 0) rx->sm_clock_khz = 1965000;
    {
        // room in L2 for the resident CTAs' frame scratch (2 per SM x 128 KB), if the device offers a persisting carve-out
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, cfg->device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, cfg->device);
        size_t want = (size_t)2 * rx->nsm * QPSK_SCRATCH_REGION_FLOATS * sizeof(float);
        if ((size_t)max_persist < want) want = (size_t)max_persist;
        if (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
            rx->l2_persist_bytes = want;
            rx->l2_window_bytes = (size_t)max_window;
        } else {
            cudaGetLastError();
        }
    }
    rx->id = g_next_id++;
    rx->C = nchan;
    rx->Cpad = (nchan + QPSK_GROUP - 1) / QPSK_GROUP * QPSK_GROUP;
    rx->maxF = max_frames;
    rx->N = cfg->frame_size;
    rx->sps = sps;
    rx->nsym = rx->N / sps;
    rx->nslots = max_frames + 1;
    // design-time constants on the host, with the host libm, as the reference computes them
    qpsk_host_rrc_make(rx->taps, cfg->ntaps, cfg->fs, cfg->rs, cfg->rrc_alpha);           // qpsk.c:308
    float t2[2];
    qpsk_host_cis(kTau * (double)cfg->center / (double)cfg->fs, 1, t2);                    // qpsk.c:342
    rx->rect = make_float2(t2[0], t2[1]);
    qpsk_host_cis(3.14159265358979323846 / 4.0, 0, t2);                                    // qpsk.c:75
    rx->rot45 = make_float2(t2[0], t2[1]);
    if (cfg->flags & QPSK_B200_SLICE_DIAGONAL) rx->rot45 = make_float2(1.0f, 0.0f);        // extension: qpsk_demod without its 45 degrees
    qpsk_host_loop_create(&rx->loop, cfg->loop_bw, -1.0f, 1.0f);                           // qpsk.c:302

    const size_t Cp = rx->Cpad, F = max_frames, N = rx->N, S = rx->nsym;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    if (cudaStreamCreateWithFlags(&rx->stream, cudaStreamNonBlocking) != cudaSuccess) e = cudaGetLastError();
    for (auto& ev : rx->ev_fr) if (e == cudaSuccess) e = cudaEventCreate(&ev);
    for (auto& ev : rx->ev_lp) if (e == cudaSuccess) e = cudaEventCreate(&ev);
    alloc((void**)&rx->d_pcm_tail, Cp * QPSK_CHUNK * sizeof(int16_t));
    alloc((void**)&rx->d_phasor2[0], (QPSK_CHUNK + F * N) * sizeof(float2));
    alloc((void**)&rx->d_phasor2[1], (QPSK_CHUNK + F * N) * sizeof(float2));
    alloc((void**)&rx->d_ph_state2, 2 * sizeof(float2));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&rx->s_k0, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&rx->ev_k0_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&rx->ev_call_start, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&rx->s_loop, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void**)&rx->d_chunk_flags, (QPSK_MAX_CHUNKS + 1) * sizeof(int));      // + the chasing loop's watchdog word
    if (e == cudaSuccess) e = cudaMemset(rx->d_chunk_flags, 0, (QPSK_MAX_CHUNKS + 1) * sizeof(int));
    rx->chase_ticket = 0;
    rx->no_chase = false;
    if (e == cudaSuccess) {
        // CUDA loads a kernel's code at its first launch, and that load can wait for the device to drain: a first launch made
        // while the chasing loop spins for its result would never start.  Everything a chunked call launches is loaded here.
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, phasor_table_kernel);
        cudaFuncGetAttributes(&fa, save_pcm_tail_kernel);
        cudaFuncGetAttributes(&fa, chunk_signal_kernel);
        cudaFuncGetAttributes(&fa, costas_chase_kernel);
        cudaFuncGetAttributes(&fa, costas_follow_kernel);
        cudaFuncSetAttribute(costas_follow_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncGetAttributes(&fa, costas_kernel);
        if (rx->sps == 4) {
            cudaFuncGetAttributes(&fa, rx_front_kernel<127, 4, QPSK_MODE_EXACT>); cudaFuncGetAttributes(&fa, rx_front_kernel<127, 4, QPSK_MODE_FAST>);
            cudaFuncGetAttributes(&fa, rx_front2_kernel<127, 4, QPSK_MODE_EXACT>); cudaFuncGetAttributes(&fa, rx_front2_kernel<127, 4, QPSK_MODE_FAST>);
        } else {
            cudaFuncGetAttributes(&fa, rx_front_kernel<127, 8, QPSK_MODE_EXACT>); cudaFuncGetAttributes(&fa, rx_front_kernel<127, 8, QPSK_MODE_FAST>);
            cudaFuncGetAttributes(&fa, rx_front2_kernel<127, 8, QPSK_MODE_EXACT>); cudaFuncGetAttributes(&fa, rx_front2_kernel<127, 8, QPSK_MODE_FAST>);
        }
        cudaGetLastError();
    }
    rx->chase_smem = 0;
    if (e == cudaSuccess) {
        int kb = 120;                       // 228 KB per SM - 120 - 1 < 110 + 1: no room for a front-end CTA
        if (const char* ck = getenv("QPSK_B200_CHASE_SMEM_KB")) kb = atoi(ck);
        if (kb > 0 && cudaFuncSetAttribute(costas_chase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024) == cudaSuccess) rx->chase_smem = kb * 1024;
        else cudaGetLastError();
    }
    if (const char* nc = getenv("QPSK_B200_NO_CHASE")) rx->no_chase = atoi(nc) != 0;
    rx->follow_now = false; rx->follow_fblocks = 0; rx->follow_mode = 0; rx->follow_fb_forced = 0;
    if (const char* fo = getenv("QPSK_B200_FOLLOW")) rx->follow_mode = atoi(fo) != 0;
    if (const char* fo = getenv("QPSK_B200_FOLLOW_FB")) rx->follow_fb_forced = atoi(fo);
    rx->relay_mode = 1;
    rx->chunk_div = QPSK_MAX_CHUNKS;
    if (const char* cd = getenv("QPSK_B200_CHUNK_DIV")) { const int v = atoi(cd); if (v >= 1 && v <= QPSK_MAX_CHUNKS) rx->chunk_div = v; }
    rx->est_emit = false; rx->est_emit_n = 0; rx->no_est_emit = false;
    if (const char* ne = getenv("QPSK_B200_NO_EST_EMIT")) rx->no_est_emit = atoi(ne) != 0;
    rx->plan_chunks = 0; rx->plan_fblocks = 0; rx->plan_loop = QPSK_B200_LOOP_STANDALONE;
    rx->host_tail_chunks = 4;
    rx->host_slice_halfwaves = 1;
    if (const char* hw = getenv("QPSK_B200_HOST_SLICE_HALFWAVES")) { const int v = atoi(hw); if (v >= 1 && v <= 64) rx->host_slice_halfwaves = v; }
    if (const char* hc = getenv("QPSK_B200_HOST_CHUNKS")) rx->host_tail_chunks = atoi(hc);
    if (const char* re = getenv("QPSK_B200_RELAY")) rx->relay_mode = atoi(re);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&rx->ev_front, cudaEventDisableTiming);
    // the front end's per-CTA frame scratch, for the largest grid a call can ask for (every group x every frame, capped
    // at eight waves: the policy never cuts finer than that), so nothing is allocated inside a stream-ordered call
    if (e == cudaSuccess) {
        long long ctas = (long long)(rx->Cpad / QPSK_GROUP) * max_frames;
        const long long cap = (long long)(rx->Cpad / QPSK_GROUP) > 16LL * rx->nsm ? (long long)(rx->Cpad / QPSK_GROUP) : 16LL * rx->nsm;
        if (ctas > cap) ctas = cap;
        rx->front_scratch_bytes = (size_t)(ctas + (long long)rx->nsm * QPSK_SCRATCH_SLOTS) * QPSK_SCRATCH_REGION_FLOATS * sizeof(float);
        e = cudaMalloc((void**)&rx->d_front_scratch, rx->front_scratch_bytes);
        if (e != cudaSuccess) rx->front_scratch_bytes = 0;
        if (e == cudaSuccess) {
            rx->block_progress_n = (size_t)ctas + 1;
            e = cudaMalloc((void**)&rx->d_block_progress, rx->block_progress_n * sizeof(unsigned long long));
            if (e == cudaSuccess) e = cudaMemset(rx->d_block_progress, 0, rx->block_progress_n * sizeof(unsigned long long));
        }
        if (e == cudaSuccess) e = cudaMalloc((void**)&rx->d_scratch_slots, (size_t)rx->nsm * (QPSK_SCRATCH_SLOTS + 1) * sizeof(int));   // + the first wave's arrival counters
    }
    alloc((void**)&rx->d_dec_ring, (F + 1) * S * Cp * sizeof(float2));
    alloc((void**)&rx->d_index_t, F * Cp * sizeof(int));
    alloc((void**)&rx->d_loop_state, Cp * sizeof(float2));
    alloc((void**)&rx->d_dibits_t, F * (S / 16) * Cp * sizeof(unsigned));
    alloc((void**)&rx->d_track_t, F * Cp * sizeof(float2));
    if (cfg->flags & QPSK_B200_KEEP_FIR) alloc((void**)&rx->d_fir_dbg, (size_t)nchan * F * N * sizeof(float2));
    if (cfg->flags & QPSK_B200_KEEP_SYMBOLS) alloc((void**)&rx->d_costas_dbg, F * S * Cp * sizeof(float2));
    if (cfg->flags & QPSK_B200_DECODE_FRAMES) {
        alloc((void**)&rx->d_frames_t, F * (S / 16) * Cp * sizeof(unsigned));
        alloc((void**)&rx->d_crc_ok_t, F * Cp);
        if (cfg->flags & QPSK_B200_RESOLVE_ROTATION) alloc((void**)&rx->d_rotation_t, F * Cp);
        alloc((void**)&rx->d_counters, 2 * sizeof(unsigned long long));
    }
    if (cfg->flags & QPSK_B200_ESTIMATE_TIMING) alloc((void**)&rx->d_timing_t, F * Cp * sizeof(float2));
    rx->prerotate = (cfg->flags & QPSK_B200_PREROTATE_OFFSET) != 0;
    if (cfg->flags & (QPSK_B200_ESTIMATE_OFFSET | QPSK_B200_PREROTATE_OFFSET)) {
        rx->est_on = true;
        alloc((void**)&rx->d_est_bursts, (size_t)nchan * 1024 * sizeof(float2));
        alloc((void**)&rx->d_est_bins, (size_t)nchan * sizeof(int));
        alloc((void**)&rx->d_est_mag, (size_t)nchan * sizeof(float));
    }
    if (e != cudaSuccess) {
        rx_free(rx);
        return fail(QPSK_B200_ERR_CUDA, "allocating receiver state failed: %s", cudaGetErrorString(e));
    }
    int rc = qpsk_b200_rx_reset(rx);
    if (rc != 0) { rx_free(rx); return rc; }
    *out = rx;
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_rx_reset(qpsk_b200_rx* rx) {
    if (!rx) return fail(QPSK_B200_ERR_ARG, "null receiver");
    CU(cudaSetDevice(rx->cfg.device));
    CU(cudaDeviceSynchronize());                 // nothing of an earlier call may still be in flight on any stream
    const size_t Cp = rx->Cpad, S = rx->nsym;
    cudaStream_t s = rx->stream;
    CU(cudaMemsetAsync(rx->d_pcm_tail, 0, Cp * QPSK_CHUNK * sizeof(int16_t), s));
    CU(cudaMemsetAsync(rx->d_dec_ring, 0, (size_t)rx->nslots * S * Cp * sizeof(float2), s));
    // slot 0 plays "the table of a previous call with zero frames": its first 128 entries are the phasors of the
    // samples before the stream start (they only ever multiply zero PCM; keep them finite) and its state is cmplx(0)
    CU(cudaStreamSynchronize(rx->s_k0));
    float2 ones[QPSK_CHUNK];
    for (auto& v : ones) v = make_float2(1.0f, 0.0f);
    CU(cudaMemcpyAsync(rx->d_phasor2[0], ones, sizeof ones, cudaMemcpyHostToDevice, s));
    float c0[2];
    qpsk_host_cis(0.0, 0, c0);                                                             // qpsk.c:341
    const float2 ph0 = make_float2(c0[0], c0[1]);
    CU(cudaMemcpyAsync(rx->d_ph_state2, &ph0, sizeof ph0, cudaMemcpyHostToDevice, s));
    rx->ph_cur = 0; rx->ph_cur_frames = 0; rx->ph_spec_valid = false;
    // d_phase = d_freq = 0 (costas_loop.c:32-33)
    CU(cudaMemsetAsync(rx->d_loop_state, 0, Cp * sizeof(float2), s));
    if (rx->d_counters) CU(cudaMemsetAsync(rx->d_counters, 0, 2 * sizeof(unsigned long long), s));
    CU(cudaMemsetAsync(rx->d_scratch_slots, 0, (size_t)rx->nsm * QPSK_SCRATCH_SLOTS * sizeof(int), s));      // every slot free
    CU(cudaStreamSynchronize(s));
    rx->slot_base = 0;
    rx->lastF = 0;
    rx->loop_seeded = false;
    rx->inflight = 0;
    rx->needs_reset = false;
    return QPSK_B200_OK;
}



// ---- bit-stage helpers shared by the receiver and the standalone entry points -----------------
static Keystream make_keystream(int nbytes) {
    Keystream ks;
    memset(&ks, 0, sizeof ks);
    uint16_t reg = (uint16_t)QPSK_SCRAMBLE_SEED;                 // bit-scramble.c:46-55, reset per frame
    for (int bit = 0; bit < nbytes * 8; bit++) ks.w[bit >> 5] |= lfsr_step(reg) << (bit & 31);
    return ks;
}

static int launch_frame_decode(int nbytes, const unsigned* dibits_t, unsigned* frames_t, uint8_t* crc_ok_t, uint8_t* rotation_t,
                               unsigned long long* counters, int c0, int C, int Cpad, int F, cudaStream_t s) {
    if (nbytes != 16 && nbytes != 32) return fail(QPSK_B200_ERR_ARG, "frame decode supports 16- and 32-byte frames, got %d", nbytes);
    FrameDecodeArgs a;
    a.ks = make_keystream(nbytes);
    a.dibits_t = dibits_t; a.frames_t = frames_t; a.crc_ok_t = crc_ok_t; a.rotation_t = rotation_t; a.counters = counters; a.C = C; a.Cpad = Cpad; a.F = F; a.c0 = c0;
    dim3 grid((C - c0 + 127) / 128, F);
    if (nbytes == 32) frame_decode_kernel<32><<<grid, 128, 0, s>>>(a);
    else frame_decode_kernel<16><<<grid, 128, 0, s>>>(a);
    CU(cudaGetLastError());
    return 0;
}

// ---- one process call = frame chunks x channel slices; every (chunk, slice) is one RxJob ----------
// K0: the mixer phasors of the next `F` frames (see phasor_table_kernel).  If the previous chunk already evaluated a
// table for this frame count on the side stream, just wait for it; otherwise evaluate it now.
static int rx_begin_chunk(qpsk_b200_rx* rx, int F, cudaStream_t s) {
    const int nxt = rx->ph_cur ^ 1;
    if (rx->ph_spec_valid && rx->ph_spec_frames == F) {
        CU(cudaStreamWaitEvent(s, rx->ev_k0_done, 0));
    } else {
        if (rx->ph_spec_valid) CU(cudaStreamSynchronize(rx->s_k0));       // a table nobody wants is being written into `nxt`
        phasor_table_kernel<<<1, QPSK_CHUNK, 0, s>>>(rx->d_phasor2[rx->ph_cur], rx->ph_cur_frames, rx->d_ph_state2 + rx->ph_cur,
                                                     rx->d_phasor2[nxt], rx->d_ph_state2 + nxt, rx->rect, F, rx->N);
        CU(cudaGetLastError());
        rx->launches += 1;
    }
    rx->ph_cur = nxt; rx->ph_cur_frames = F; rx->ph_spec_valid = false;
    // the table of a next chunk with the same frame count, evaluated on the side stream while this one runs; it
    // overwrites the other slot, which the previous chunk's kernels (all enqueued before this point on `s`) still read
    CU(cudaEventRecord(rx->ev_call_start, s));
    CU(cudaStreamWaitEvent(rx->s_k0, rx->ev_call_start, 0));
    phasor_table_kernel<<<1, QPSK_CHUNK, 0, rx->s_k0>>>(rx->d_phasor2[rx->ph_cur], F, rx->d_ph_state2 + rx->ph_cur,
                                                        rx->d_phasor2[rx->ph_cur ^ 1], rx->d_ph_state2 + (rx->ph_cur ^ 1), rx->rect, F, rx->N);
    CU(cudaGetLastError());
    CU(cudaEventRecord(rx->ev_k0_done, rx->s_k0));
    rx->ph_spec_valid = true; rx->ph_spec_frames = F;
    rx->launches += 1;
    return 0;
}

static void rx_end_chunk(qpsk_b200_rx* rx, int F) { rx->slot_base = (rx->slot_base + F) % rx->nslots; }

__global__ void symbol_power4_kernel(const float2* __restrict__ ring, float2* __restrict__ bursts, int c_begin, int c_end, int Cpad, int nsym,
                                     int nslots, int first_slot, int n);
extern "C" int qpsk_b200_fft_create(int n, int device, qpsk_b200_fft** out);
extern "C" int qpsk_b200_fft_argmax_device(qpsk_b200_fft* f, const float* d_in, int nbursts, int32_t* d_bin, float* d_mag2, void* cuda_stream);

// burst length of the in-call estimator: the largest power of two <= min(1024, symbols of the call)
static int est_burst_length(const qpsk_b200_rx* rx, int F) {
    int n = 32;
    while (2 * n <= 1024 && 2 * n <= F * rx->nsym) n *= 2;
    return n;
}

// Launch policy, in units of "one frame of one resident front-end CTA".  Nothing here is a measured constant of one
// particular box: the unit follows from the filter's arithmetic (2 CTAs share an SM's FP32 pipe at ~80 % of 32 / 64
// complex tap-updates per clock) and the device's own SM count and clock, read at context creation.
struct RxCostModel {
    int slots;          // resident front-end CTAs: 2 per SM
    double unit_us;     // one frame (512 samples x 32 channels x 127 taps) for one of two CTAs sharing an SM
    double sym_us;      // latency of one Costas-loop symbol for one stream (~540 dependent cycles)
    double loop_us_per_sym_chan;   // throughput cost of the stand-alone loop kernel per symbol and channel, machine-wide
};
static RxCostModel rx_cost_model(const qpsk_b200_rx* rx) {
    RxCostModel m;
    const double clk_hz = rx->sm_clock_khz * 1e3;
    const double taps_per_clk_sm = (rx->cfg.mode == QPSK_B200_MODE_FAST ? 64.0 : 32.0) * 0.8;
    m.slots = 2 * rx->nsm;
    m.unit_us = 2.0 * QPSK_GROUP * 512.0 * 127.0 / (taps_per_clk_sm * clk_hz) * 1e6;
    m.sym_us = 540.0 / clk_hz * 1e6;
    // ~105 instructions per symbol and lane, one warp instruction per scheduler and clock, 4 schedulers per SM, ~50 % issue
    // efficiency for a latency-bound kernel: 105 / (32 lanes x 4 x nsm x 0.5) clocks per symbol and channel
    m.loop_us_per_sym_chan = 105.0 / (32.0 * 4.0 * rx->nsm * 0.5) / clk_hz * 1e6;
    return m;
}

struct RxJob {
    const int16_t* d_pcm;   // rows of channels [c0, c0 + nc), the job's F frames at the start of each row
    size_t pcm_row;         // row stride in samples
    int c0, nc;             // channel slice; c0 is a multiple of 32
    int F, f_off;           // frames of this job and their position in the call (outputs are indexed by call frame)
};

// How many frame blocks per channel group?  One block (the CTA owns whole streams) lets the Costas loop ride along in
// the CTA's spare warp; more blocks fill the machine when channels are few, or when the groups are an awkward number of
// waves (16,384 channels = 512 CTAs = 1.73 waves of 2 x nsm), at the price of the loop as its own kernel: a block start
// costs about a third of a frame, the front end without the loop runs ~3 % faster, the stand-alone loop needs
// max(latency of one stream, its share of the machine).
static int rx_frame_blocks(const qpsk_b200_rx* rx, int ngroups, int nc, int F, bool loop_overlapped, double margin = 0.98) {
    const RxCostModel m = rx_cost_model(rx);
    int fblocks = 1;
    const double loop_units = loop_overlapped ? 0.0
        : fmax((double)F * rx->nsym * m.sym_us, m.loop_us_per_sym_chan * (double)nc * F * rx->nsym) / m.unit_us;
    double best = rx->no_fuse ? 1e30 : (double)((ngroups + m.slots - 1) / m.slots) * F;
    for (int fb = rx->no_fuse ? 1 : 2; fb <= F; fb++) {              // fb = 1 with the loop fused is `best` already
        const int fpb = (F + fb - 1) / fb, nb = (F + fpb - 1) / fpb;
        if (nb != fb) continue;                                     // same split as a smaller fb
        const double waves = (double)(((long long)ngroups * nb + m.slots - 1) / m.slots);
        // (a front end without the loop warp runs ~3 % faster -- unless the loop runs beside it as its own kernel)
        const double t = waves * (fpb + 0.3) * (loop_overlapped ? 1.0 : 0.97) + loop_units;
        if (t < best * margin) { best = t; fblocks = nb; }          // hysteresis in favour of fewer blocks (2 % unless the caller asks for more)
    }
    return fblocks;
}

// Frame blocks of a RELAYED launch: the CTAs own frame blocks AND keep the loop in their spare warp; the loop state travels from
// the CTA of block k to the CTA of block k + 1 of the same channels through a relay word (rx_front_kernel, fuse_costas == 2).
// Same instructions as the fused kernel, but the machine fills evenly when the channel groups are an awkward number of waves
// of whole-stream CTAs: 16,384 channels are 512 CTAs = 1.73 waves (the second leaves 80 SMs with one CTA), 8,192 are 0.86
// (40 SMs with one CTA, and every CTA as long as the call).  Measured (tools/strong_time.py, 64 frames): 8,192 channels
// 5.75 -> 4.89 ms, 16,384 11.06 -> 9.56 ms, 32,768 19.81 (frame chunks) -> 18.95 ms, 65,536 unchanged within 0.3 %.
// Model, in frame units: whole-stream CTAs run in lock step, waves x F; blocks are short and numerous, so they cost their
// total work over the resident CTAs plus half a block of tail, a block start costing ~0.15 frame; the relayed loop of one
// channel group stays a serial chain (~0.9 unit per 128-symbol frame between the filter warps).  Few channels (under half a
// wave of groups) are left to the frame-chunk plan, whose loop runs on SMs of its own.  0 = the plain fused kernel.
static int rx_relay_blocks(const qpsk_b200_rx* rx, int ngroups, int F) {
    if (rx->relay_mode == 0 || rx->no_fuse || !rx->front_v1 || !rx->d_block_progress || F < 2) return 0;
    const RxCostModel m = rx_cost_model(rx);
    if ((size_t)(rx->Cpad / QPSK_GROUP) > rx->block_progress_n) return 0;
    if (rx->relay_mode > 1) {                                                  // forced (tests, sweeps): legal for any shape, only not fast
        const int fpb = (F + rx->relay_mode - 1) / rx->relay_mode;
        const int nb = (F + fpb - 1) / fpb;
        return nb > 1 ? nb : 0;
    }
    if (2 * ngroups < m.slots) return 0;
    const double chain = 0.9 * F * rx->nsym / 128.0;
    double best = (double)((ngroups + m.slots - 1) / m.slots) * F * 0.95;      // 5 % in favour of whole-stream CTAs
    int fblocks = 0;
    for (int fb = 2; fb <= F; fb++) {
        const int fpb = (F + fb - 1) / fb, nb = (F + fpb - 1) / fpb;
        if (nb != fb) continue;
        const double t = fmax((double)ngroups * nb * (fpb + 0.15) / m.slots + 0.5 * fpb, chain);
        if (t < best) { best = t * 0.99; fblocks = nb; }                       // 1 % in favour of fewer blocks
    }
    return fblocks;
}

// Frame blocks of a FOLLOWED call (the loop beside the front end, costas_follow_kernel), 0 = the fused kernel is the better
// plan.  The front end is waves x (frames per block + a third of a frame per block start); the loop is one dependency chain
// per channel group that, issued between the filter warps of two co-resident front-end CTAs, takes ~1.27 front-end frame
// units per frame (measured: 6.7 ms for 64 frames beside a 4.5 ms front end; the ratio is clock-independent), on
// QPSK_FOLLOW_WARPS_PER_SM x nsm warps.  8,192 channels (256 groups) are bound by that chain and stay fused; 16,384 and
// 32,768 are 1.73 and 3.46 waves as whole streams, 6.92 and 6.92 as blocks of 16 and 32 frames.
static int rx_follow_blocks(const qpsk_b200_rx* rx, int ngroups, int F) {
    const RxCostModel m = rx_cost_model(rx);
    const double warps = (double)QPSK_FOLLOW_WARPS_PER_SM * rx->nsm;
    const double loop_floor = 1.27 * F * ceil(ngroups / warps);     // the busiest warp: its groups one after the other, every frame block
    const double fused = (double)((ngroups + m.slots - 1) / m.slots) * F;
    double best = fused * 0.95;                                      // 5 % in favour of the fused kernel
    int fblocks = 0;
    for (int fb = 2; fb <= F; fb++) {
        const int fpb = (F + fb - 1) / fb, nb = (F + fpb - 1) / fpb;
        if (nb != fb) continue;
        const double waves = (double)(((long long)ngroups * nb + m.slots - 1) / m.slots);
        const double t = fmax(waves * (fpb + 0.3), loop_floor);
        if (t < best) { best = t; fblocks = nb; }
    }
    return fblocks;
}

static int rx_ensure_front_scratch(qpsk_b200_rx* rx, int grid) {
    // per-CTA frame scratch (512 samples x 2 components x 32 lanes of float); rewritten every frame, so it lives in L2.
    // Sized once for the largest grid any call can ask for (every channel group x every frame), so no allocation ever
    // happens in the middle of a stream-ordered call.
    const size_t need = (size_t)(grid + rx->nsm * QPSK_SCRATCH_SLOTS) * QPSK_SCRATCH_REGION_FLOATS * sizeof(float);
    if (rx->front_scratch_bytes >= need) return 0;
    CU(cudaDeviceSynchronize());
    if (rx->d_front_scratch) { cudaFree(rx->d_front_scratch); rx->d_front_scratch = nullptr; rx->front_scratch_bytes = 0; }
    CU(cudaMalloc((void**)&rx->d_front_scratch, need));
    rx->front_scratch_bytes = need;
    return 0;
}

// The in-call estimator's bursts (4th power of the call's first n symbols) are left behind by the loop itself when every one
// of those symbols is consumed within the call: frame m of the call is consumed by loop frame m + 1, so n symbols need
// n / nsym + 1 frames.  Otherwise (short calls; the seeding call, whose loop runs after the estimate) symbol_power4_kernel
// makes a pass over the ring as before.
static void rx_plan_est_emit(qpsk_b200_rx* rx, int F, bool seeding) {
    const int n = est_burst_length(rx, F);
    rx->est_emit = rx->est_on && !seeding && rx->d_est_bursts != nullptr && !rx->no_est_emit && n % rx->nsym == 0 && n / rx->nsym + 1 <= F;
    rx->est_emit_n = rx->est_emit ? n : 0;
}

static CostasArgs rx_costas_args(const qpsk_b200_rx* rx, const RxJob& j) {
    const size_t Cp = rx->Cpad, S = rx->nsym, fo = j.f_off;
    CostasArgs ca;
    ca.dec_ring = rx->d_dec_ring; ca.index_t = rx->d_index_t + fo * Cp; ca.loop_state = rx->d_loop_state;
    ca.dibits_t = rx->d_dibits_t + fo * (S / 16) * Cp;
    ca.costas_dbg = rx->d_costas_dbg ? rx->d_costas_dbg + fo * S * Cp : nullptr;
    ca.track_t = rx->d_track_t + fo * Cp;
    ca.C = rx->C; ca.Cpad = rx->Cpad; ca.F = j.F; ca.nsym = rx->nsym; ca.sps = rx->sps; ca.N = rx->N;
    ca.c0 = j.c0; ca.c1 = (j.c0 + j.nc < rx->C) ? j.c0 + j.nc : rx->C;
    ca.slot_base = rx->slot_base; ca.nslots = rx->nslots; ca.ub_mode = rx->cfg.ub_mode;
    // TRANSIENT_SYMBOLS: the estimator reads the call's first symbols after the kernel, those slots stay
    ca.est_bursts = rx->est_emit ? rx->d_est_bursts : nullptr; ca.est_n = rx->est_emit_n; ca.est_f_off = (int)fo;
    ca.discard_from = -1;
    if (rx->cfg.flags & QPSK_B200_TRANSIENT_SYMBOLS)
        ca.discard_from = rx->est_on ? (est_burst_length(rx, j.F + j.f_off) + rx->nsym - 1) / rx->nsym + 1 - j.f_off : 1;
    if (ca.discard_from >= 0 && ca.discard_from < 1) ca.discard_from = 1;
    ca.alpha = rx->loop.alpha; ca.beta = rx->loop.beta; ca.max_freq = rx->loop.max_freq; ca.min_freq = rx->loop.min_freq;
    ca.rot45 = rx->rot45;
    return ca;
}

// K1 (+ the fused loop) and the PCM tail of one job on stream s.  *fused tells the caller whether K3 is still to run.
static int rx_launch_front(qpsk_b200_rx* rx, const RxJob& j, bool loop_overlapped, cudaStream_t s, bool* fused_out, int timed_chunk = -1) {
    const int N = rx->N;
    const size_t Cp = rx->Cpad;
    RxFrontArgs fa;
    fa.pcm = j.d_pcm; fa.pcm_row = j.pcm_row; fa.pcm_tail = rx->d_pcm_tail; fa.phasor = rx->d_phasor2[rx->ph_cur];
    fa.dec_ring = rx->d_dec_ring; fa.index_t = rx->d_index_t + (size_t)j.f_off * Cp; fa.fir_dbg = rx->d_fir_dbg;
    fa.timing_t = rx->d_timing_t ? rx->d_timing_t + (size_t)j.f_off * Cp : nullptr;
    fa.C = rx->C; fa.Cpad = rx->Cpad; fa.F = j.F; fa.N = N; fa.chan_base = j.c0; fa.chan_count = j.nc;
    fa.slot_base = rx->slot_base; fa.nslots = rx->nslots; fa.ub_mode = rx->cfg.ub_mode;
    const int ngroups = (j.nc + QPSK_GROUP - 1) / QPSK_GROUP;
    const int relay_fb = (rx->follow_now || loop_overlapped) ? 0 : rx_relay_blocks(rx, ngroups, j.F);
    int fblocks = rx->follow_now ? rx->follow_fblocks : relay_fb > 1 ? relay_fb : rx_frame_blocks(rx, ngroups, j.nc, j.F, loop_overlapped);
    fa.frames_per_block = (j.F + fblocks - 1) / fblocks;
    fblocks = (j.F + fa.frames_per_block - 1) / fa.frames_per_block;
    const int grid = ngroups * fblocks;
    int rc = rx_ensure_front_scratch(rx, grid);      // a no-op after creation (sized for Cpad/32 x maxF CTAs)
    if (rc) return rc;
    fa.scratch = rx->d_front_scratch;
    fa.scratch_slots = rx->d_scratch_slots;
    fa.scratch_nslots = rx->nsm * QPSK_SCRATCH_SLOTS;
    fa.block_progress = nullptr; fa.progress_ticket = 0;
    if (rx->follow_now) {
        if ((size_t)grid > rx->block_progress_n) return fail(QPSK_B200_ERR_STATE, "internal: a followed call with more CTAs than progress words");
        fa.block_progress = rx->d_block_progress;
        fa.progress_ticket = (unsigned long long)rx->chase_ticket << 32;
    }
    fa.dephase_counters = rx->d_scratch_slots + rx->nsm * QPSK_SCRATCH_SLOTS;
    fa.dephase_cycles = rx->dephase_cycles;
    if (fa.dephase_cycles > 0) CU(cudaMemsetAsync(fa.dephase_counters, 0, (size_t)rx->nsm * sizeof(int), s));
    const bool relayed = relay_fb > 1 && fblocks > 1;
    const bool fused = ((fblocks == 1) && !rx->no_fuse) || relayed;
    fa.fuse_costas = relayed ? 2 : fused ? 1 : 0;
    fa.relay_group_base = j.c0 / QPSK_GROUP; fa.relay_watchdog = rx->d_chunk_flags + QPSK_MAX_CHUNKS; fa.relay_ticket = 0;
    if (relayed) {
        fa.relay_ticket = ++rx->chase_ticket;
        fa.block_progress = rx->d_block_progress;
        fa.progress_ticket = (unsigned long long)fa.relay_ticket << 32;
    }
    fa.costas = rx_costas_args(rx, j);
    cudaError_t e;
    const bool fast = rx->cfg.mode == QPSK_B200_MODE_FAST;
    if (timed_chunk >= 0) CU(cudaEventRecord(rx->ev_fr[2 * timed_chunk], s));
    const size_t pb = rx->l2_persist_bytes, pw = rx->l2_window_bytes;
    if (rx->front_v1) {
        if (rx->sps == 4) e = fast ? launch_front<127, 4, QPSK_MODE_FAST>(fa, rx->taps, grid, s, pb, pw) : launch_front<127, 4, QPSK_MODE_EXACT>(fa, rx->taps, grid, s, pb, pw);
        else              e = fast ? launch_front<127, 8, QPSK_MODE_FAST>(fa, rx->taps, grid, s, pb, pw) : launch_front<127, 8, QPSK_MODE_EXACT>(fa, rx->taps, grid, s, pb, pw);
    } else {
        if (rx->sps == 4) e = fast ? launch_front2<127, 4, QPSK_MODE_FAST>(fa, rx->taps, grid, s, pb, pw) : launch_front2<127, 4, QPSK_MODE_EXACT>(fa, rx->taps, grid, s, pb, pw);
        else              e = fast ? launch_front2<127, 8, QPSK_MODE_FAST>(fa, rx->taps, grid, s, pb, pw) : launch_front2<127, 8, QPSK_MODE_EXACT>(fa, rx->taps, grid, s, pb, pw);
    }
    if (e != cudaSuccess) return fail(QPSK_B200_ERR_CUDA, "front-end kernel launch failed: %s", cudaGetErrorString(e));
    if (timed_chunk >= 0) CU(cudaEventRecord(rx->ev_fr[2 * timed_chunk + 1], s));
    // carry the last 128 PCM samples of every channel (the next chunk's filter history)
    const int live = fa.costas.c1 - j.c0;
    save_pcm_tail_kernel<<<(live * 16 + 255) / 256, 256, 0, s>>>(j.d_pcm, rx->d_pcm_tail + (size_t)j.c0 * QPSK_CHUNK, live, j.pcm_row, (size_t)j.F * N);
    CU(cudaGetLastError());
    rx->launches += 2;
    rx->last_fused = fused;
    rx->plan_chunks = 1; rx->plan_fblocks = fblocks;
    rx->plan_loop = relayed ? QPSK_B200_LOOP_RELAYED : fused ? QPSK_B200_LOOP_FUSED : rx->follow_now ? QPSK_B200_LOOP_FOLLOWING : QPSK_B200_LOOP_STANDALONE;
    *fused_out = fused;
    return 0;
}

// K3 (when it did not ride along in K1) and K4 of one job on stream s
static int rx_launch_loop_and_decode(qpsk_b200_rx* rx, const RxJob& j, bool fused, cudaStream_t s, int timed_chunk = -1) {
    const CostasArgs ca = rx_costas_args(rx, j);
    const int live = ca.c1 - j.c0;
    if (!fused) {
        if (timed_chunk >= 0) CU(cudaEventRecord(rx->ev_lp[2 * timed_chunk], s));
        costas_kernel<<<(live + 127) / 128, 128, 0, s>>>(ca);
        CU(cudaGetLastError());
        if (timed_chunk >= 0) CU(cudaEventRecord(rx->ev_lp[2 * timed_chunk + 1], s));
        rx->launches += 1;
    }
    if (rx->d_frames_t) {   // K4: descramble -> de-interleave -> CRC16 per frame
        const size_t Cp = rx->Cpad, W = rx->nsym / 16, fo = j.f_off;
        int rc = launch_frame_decode(rx->nsym / 4, rx->d_dibits_t + fo * W * Cp, rx->d_frames_t + fo * W * Cp, rx->d_crc_ok_t + fo * Cp,
                                     rx->d_rotation_t ? rx->d_rotation_t + fo * Cp : nullptr, rx->d_counters, j.c0, ca.c1, rx->Cpad, j.F, s);
        if (rc) return rc;
        rx->launches += 1;
    }
    return 0;
}

// the FFT frequency estimator over channels [c0, c1) of a call of F frames whose first frame sits in ring slot first_slot
static int rx_launch_estimator(qpsk_b200_rx* rx, int c0, int c1, int F, int first_slot, cudaStream_t s) {
    const int n = est_burst_length(rx, F);
    if (!rx->est_fft || rx->est_fft_n != n) {
        if (rx->est_fft) { CU(cudaDeviceSynchronize()); qpsk_b200_fft_destroy(rx->est_fft); rx->est_fft = nullptr; }
        int rc = qpsk_b200_fft_create(n, rx->cfg.device, &rx->est_fft);
        if (rc) return rc;
        rx->est_fft_n = n;
    }
    rx->est_call_n = n;
    if (!(rx->est_emit && rx->est_emit_n == n)) {          // the loop has not left the bursts behind: one pass over the ring
        dim3 grid((c1 - c0 + 31) / 32, (n + 31) / 32), block(32, 8);
        symbol_power4_kernel<<<grid, block, 0, s>>>(rx->d_dec_ring, rx->d_est_bursts, c0, c1, rx->Cpad, rx->nsym, rx->nslots, first_slot, n);
        CU(cudaGetLastError());
        rx->launches += 1;
    }
    int rc = qpsk_b200_fft_argmax_device(rx->est_fft, reinterpret_cast<const float*>(rx->d_est_bursts + (size_t)c0 * n), c1 - c0,
                                         rx->d_est_bins + c0, rx->d_est_mag + c0, s);
    if (rc) return rc;
    rx->launches += 1;
    return 0;
}

// PREROTATE_OFFSET: d_freq of every channel from the estimator's bin, as set_frequency(TAU * offset_hz / RS) would
// (costas_loop.c:117-125 clamps to [min_freq, max_freq]); offset_hz = signed bin * rs / (4 n), rounded to float as OUT_OFFSET_HZ is
__global__ void seed_loop_freq_kernel(float2* __restrict__ loop_state, const int* __restrict__ bins, int c0, int c1, int n, float rs,
                                      float min_freq, float max_freq) {
    const int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    const int b = bins[c], sb = b >= n / 2 ? b - n : b;
    const float hz = __double2float_rn(__ddiv_rn(__dmul_rn((double)sb, (double)rs), (double)(4 * n)));
    float f = __double2float_rn(__ddiv_rn(__dmul_rn(6.283185307179586476925286766559, (double)hz), (double)rs));
    f = fminf(fmaxf(f, min_freq), max_freq);
    loop_state[c].y = f;
}

static int rx_seed_loop(qpsk_b200_rx* rx, int c0, int c1, int F, int first_slot, cudaStream_t s) {
    int rc = rx_launch_estimator(rx, c0, c1, F, first_slot, s);
    if (rc) return rc;
    seed_loop_freq_kernel<<<(c1 - c0 + 127) / 128, 128, 0, s>>>(rx->d_loop_state, rx->d_est_bins, c0, c1, rx->est_call_n, rx->cfg.rs,
                                                                  rx->loop.min_freq, rx->loop.max_freq);
    CU(cudaGetLastError());
    rx->launches += 1;
    return 0;
}

// Frame chunks of a call.  With few channels the loop cannot ride along in the front end (a CTA would have to own whole
// streams and there are too few CTAs for that), and as one kernel after the front end it is a pure latency chain:
// 1,024 streams x 16,384 symbols x ~540 cycles = 4.7 ms behind a 2.3 ms front end.  Cutting the call into frame chunks
// lets the loop of chunk k run on a second stream under the front end of chunks k+1.., and lets the host path copy
// chunk k+1 in and chunk k-1 out meanwhile.  State carries between chunks exactly as between calls.
static int rx_plan_chunks(const qpsk_b200_rx* rx, int nc, int F, int div = 0) {
    if (div <= 0) div = rx->chunk_div;
    if (rx->d_fir_dbg || rx->d_costas_dbg || rx->no_chunk) return F;         // debug taps are indexed by the call's frames
    if (rx->prerotate && !rx->loop_seeded) return F;                          // the seeding call: front end, estimator, then the loop
    const int ngroups = (nc + QPSK_GROUP - 1) / QPSK_GROUP;
    if (F < 32) return F;
    if (rx_relay_blocks(rx, ngroups, F) > 1) return F;                        // frame blocks with the loop relayed from CTA to CTA
    if (rx_frame_blocks(rx, ngroups, nc, F, false) == 1) return F;            // the fused kernel is the better plan
    int fc = (F + div - 1) / div;
    if (fc < 8) fc = 8;
    return fc;
}

// Is the loop of a chunked call (fc < F frames per chunk) ONE kernel that chases the chunks?  (the reasons are with its use in rx_run_call)
static bool rx_chase_plan(const qpsk_b200_rx* rx, int F, int fc, bool seeding) {
    const int ngroups_all = (rx->C + QPSK_GROUP - 1) / QPSK_GROUP;
    const int nchunks = (F + fc - 1) / fc;
    return fc < F && !rx->no_chase && !seeding && nchunks <= QPSK_MAX_CHUNKS && rx->chase_smem > 0
           && ((rx->C + 127) / 128) * 8 <= rx->nsm
           && rx_frame_blocks(rx, ngroups_all, rx->C, fc, true) > 1 && rx_frame_blocks(rx, ngroups_all, rx->C, F - (nchunks - 1) * fc, true) > 1;
}

// The device-resident call: frame chunks on `s`, the loop of a chunked call on the loop stream.  CUDA events bracket K1
// and K3 of every chunk (qpsk_b200_rx_last_kernel_ms sums them).
static int rx_run_call(qpsk_b200_rx* rx, const int16_t* d_pcm, size_t pcm_row, int F, cudaStream_t s) {
    const int fc = rx_plan_chunks(rx, rx->C, F);
    const bool chunked = fc < F;
    const int first_slot = (rx->slot_base + 1) % rx->nslots;
    // PREROTATE_OFFSET, first call after a reset: the loop runs as its own kernel, after the estimator has seeded it
    const bool seeding = rx->prerotate && !rx->loop_seeded;
    rx_plan_est_emit(rx, F, seeding);
    const bool saved_no_fuse = rx->no_fuse;
    if (seeding) rx->no_fuse = true;
    // chunked call: ONE loop kernel chases the chunks through flags (costas_chase_kernel) when no chunk's front end would carry
    // the loop itself; its arguments describe the whole call, taken before the chunks advance the ring
    const int ngroups_all = (rx->C + QPSK_GROUP - 1) / QPSK_GROUP;
    const int nchunks = (F + fc - 1) / fc;
    // ... and only when its CTAs (128 streams each) get SMs of their own, an eighth of the machine at most: launched first, a
    // loop CTA pins its SM's shared-memory carve-out at whatever it was placed with, and a chasing loop on EVERY SM (32,768
    // channels x 64 frames, the 2-GPU strong-scaling shape, planned as chunks) left no SM a 110 KB front-end CTA could start
    // on: the loop waited for a front end that waited for the loop, until the watchdog.  More channels than that run the
    // per-chunk loop kernels, which wait for nothing.
    const bool chase = rx_chase_plan(rx, F, fc, seeding);
    RxJob whole;
    whole.d_pcm = d_pcm; whole.pcm_row = pcm_row; whole.c0 = 0; whole.nc = rx->C; whole.F = F; whole.f_off = 0;
    const CostasArgs chase_args = rx_costas_args(rx, whole);
    // A call that is neither chunked nor worth chunking would run as whole-stream CTAs with the loop fused.  When its channel
    // groups are an awkward number of waves (16,384 channels = 1.73, 8,192 = 0.86 of 2 x nsm resident CTAs) the frames are cut
    // into blocks all the same, and the loop runs beside the front end as costas_follow_kernel: one warp per resident
    // front-end CTA, following its channel groups frame by frame through the progress words the front-end CTAs publish.
    int follow_fb = 0;
    if (!chunked && !seeding && (rx->follow_mode != 0 || rx->follow_fb_forced > 0) && !rx->no_fuse && !rx->d_fir_dbg && !rx->d_costas_dbg && rx->d_block_progress && F >= 2) {
        follow_fb = rx->follow_fb_forced > 0 ? (rx->follow_fb_forced < F ? rx->follow_fb_forced : F)
                                             : rx_follow_blocks(rx, ngroups_all, F);
        if (follow_fb > 1) {
            const int fpb = (F + follow_fb - 1) / follow_fb;
            follow_fb = (F + fpb - 1) / fpb;
        }
        if (follow_fb <= 1 || (size_t)ngroups_all * follow_fb > rx->block_progress_n) follow_fb = 0;
    }
    const int ticket = (chase || follow_fb) ? ++rx->chase_ticket : 0;
    bool loop_stream_used = false, chase_launched = false;
    int k = 0;
    rx->timed_loop = false;
    // a chasing loop that is already running must not be left waiting for chunks that will never come
    auto bail = [&](int rc) -> int {
        rx->no_fuse = saved_no_fuse;
        if (chase_launched) {
            int tickets[QPSK_MAX_CHUNKS];
            for (int i = 0; i < QPSK_MAX_CHUNKS; i++) tickets[i] = ticket;
            cudaMemcpyAsync(rx->d_chunk_flags, tickets, sizeof tickets, cudaMemcpyHostToDevice, rx->s_k0);
            cudaStreamSynchronize(rx->s_k0);
            cudaStreamSynchronize(rx->s_loop);
        }
        return rc;
    };
    if (chase) {
        // The chasing loop goes first: its few CTAs (128 streams each) are placed while the SMs are still empty and ask for so
        // much (unused) shared memory that no front-end CTA fits beside them -- the loop is a dependency chain, every cycle it
        // queues behind filter warps is a cycle of the call; the front end, which has time to spare here, runs on the other SMs.
        // It starts behind whatever the caller's stream holds (the previous call's last reader of the loop state).
        // Nothing between this launch and the last chunk's signal may synchronise with the device (the loop would wait for
        // signals the host has not enqueued yet): the one such place, the growth of the frame scratch, is done first.
        {
            const int fb = rx_frame_blocks(rx, ngroups_all, rx->C, fc, true);
            const int fpb = (fc + fb - 1) / fb;
            int rc = rx_ensure_front_scratch(rx, ngroups_all * ((fc + fpb - 1) / fpb));
            if (rc) return bail(rc);
        }
        cudaStream_t sl = rx->s_loop;
        cudaError_t e = cudaEventRecord(rx->ev_front, s);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(sl, rx->ev_front, 0);
        if (e == cudaSuccess) e = cudaEventRecord(rx->ev_lp[0], sl);
        if (e != cudaSuccess) return bail(fail(QPSK_B200_ERR_CUDA, "loop stream setup failed: %s", cudaGetErrorString(e)));
        const int live = chase_args.c1 - chase_args.c0;
        // ... as long as that costs the front end no more than an eighth of the machine
        const int loop_ctas = (live + 127) / 128;
        costas_chase_kernel<<<loop_ctas, 128, rx->chase_smem, sl>>>(chase_args, rx->d_chunk_flags, ticket, fc, QPSK_MAX_CHUNKS);
        if (cudaGetLastError() != cudaSuccess) return bail(fail(QPSK_B200_ERR_CUDA, "loop kernel launch failed"));
        chase_launched = true;
        cudaEventRecord(rx->ev_lp[1], sl);
        rx->launches += 1;
    }
    for (int f0 = 0; f0 < F; f0 += fc, k++) {
        RxJob j;
        j.d_pcm = d_pcm + (size_t)f0 * rx->N; j.pcm_row = pcm_row; j.c0 = 0; j.nc = rx->C; j.F = (F - f0 < fc) ? F - f0 : fc; j.f_off = f0;
        int rc = rx_begin_chunk(rx, j.F, s);
        if (rc) return bail(rc);
        bool fused = false;
        if (follow_fb) {
            int rc2 = rx_ensure_front_scratch(rx, ngroups_all * follow_fb);
            if (rc2) return bail(rc2);
            // the loop starts behind everything the caller's stream holds up to here (the previous call's readers of the loop state)
            cudaError_t e = cudaEventRecord(rx->ev_front, s);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(rx->s_loop, rx->ev_front, 0);
            if (e != cudaSuccess) return bail(fail(QPSK_B200_ERR_CUDA, "loop stream setup failed: %s", cudaGetErrorString(e)));
            rx->follow_now = true; rx->follow_fblocks = follow_fb;
        }
        rc = rx_launch_front(rx, j, chunked || follow_fb > 0, s, &fused, k);
        rx->follow_now = false;
        if (!rc && seeding) rc = rx_seed_loop(rx, 0, rx->C, F, first_slot, s);
        if (rc) return bail(rc);
        if (follow_fb && !fused) {
            // The following loop is launched AFTER the front end it follows: the front end never waits for the loop, so whatever
            // the block scheduler does with the loop's one-warp CTAs (two per SM fit beside two front-end CTAs: 2,560 registers
            // and 1 KB of shared memory each) the call cannot deadlock.  (Launched first, its CTAs pinned every SM's shared-memory
            // carve-out at the small setting they were placed with, and no 110 KB front-end CTA could start: measured, a
            // watchdog exit.)  Same carve-out preference as the front end for the same reason.
            const int fpb = (F + follow_fb - 1) / follow_fb;
            cudaStream_t sl = rx->s_loop;
            cudaError_t e = cudaEventRecord(rx->ev_lp[0], sl);
            if (e != cudaSuccess) return bail(fail(QPSK_B200_ERR_CUDA, "loop stream setup failed: %s", cudaGetErrorString(e)));
            const int loop_ctas = ngroups_all < QPSK_FOLLOW_WARPS_PER_SM * rx->nsm ? ngroups_all : QPSK_FOLLOW_WARPS_PER_SM * rx->nsm;
            costas_follow_kernel<<<loop_ctas, 32, 0, sl>>>(chase_args, rx->d_block_progress, (unsigned long long)ticket << 32, ngroups_all, fpb,
                                                           rx->d_chunk_flags + QPSK_MAX_CHUNKS, ticket);
            if (cudaGetLastError() != cudaSuccess) return bail(fail(QPSK_B200_ERR_CUDA, "loop kernel launch failed"));
            cudaEventRecord(rx->ev_lp[1], sl);
            rx->launches += 1;
        }
        if (follow_fb) {
            if (fused) return bail(fail(QPSK_B200_ERR_STATE, "internal: a followed call fused its loop"));
            rx_end_chunk(rx, j.F);
            continue;
        }
        if (chase) {
            if (fused) return bail(fail(QPSK_B200_ERR_STATE, "internal: a chunk of a chased call fused its loop"));
            chunk_signal_kernel<<<1, 1, 0, s>>>(rx->d_chunk_flags + k, ticket);
            if (cudaGetLastError() != cudaSuccess) return bail(fail(QPSK_B200_ERR_CUDA, "chunk signal launch failed"));
            rx->launches += 1;
            rx_end_chunk(rx, j.F);
            continue;
        }
        cudaStream_t sl = s;
        if (chunked && !fused) {
            sl = rx->s_loop;
            CU(cudaEventRecord(rx->ev_front, s));
            CU(cudaStreamWaitEvent(sl, rx->ev_front, 0));
            loop_stream_used = true;
        }
        if (!fused) rx->timed_loop = true;
        rc = rx_launch_loop_and_decode(rx, j, fused, sl, k);
        if (rc) return rc;
        rx_end_chunk(rx, j.F);
    }
    rx->timed_chunks = k;
    if (chase || follow_fb) {
        rx->timed_loop = true;
        rx->timed_loop_chunks = 1;
        if (rx->d_frames_t) {                                             // K4 over the whole call, behind the loop
            int rc = launch_frame_decode(rx->nsym / 4, rx->d_dibits_t, rx->d_frames_t, rx->d_crc_ok_t, rx->d_rotation_t, rx->d_counters, 0, chase_args.c1, rx->Cpad, F, rx->s_loop);
            if (rc) return bail(rc);
            rx->launches += 1;
        }
        loop_stream_used = true;
    } else {
        rx->timed_loop_chunks = k;
    }
    if (loop_stream_used) {
        CU(cudaEventRecord(rx->ev_front, rx->s_loop));
        CU(cudaStreamWaitEvent(s, rx->ev_front, 0));
    }
    if (seeding) { rx->no_fuse = saved_no_fuse; rx->loop_seeded = true; }
    if (rx->est_on && !seeding) {
        int rc = rx_launch_estimator(rx, 0, rx->C, F, first_slot, s);
        if (rc) return rc;
    }
    rx->lastF = F;
    rx->timed = true;
    rx->plan_chunks = k;
    if (chase) rx->plan_loop = QPSK_B200_LOOP_CHASING;
    return 0;
}

// The launch policy alone, for a device it is told about: what qpsk_b200_rx_process_device would do with nchan channels x
// nframes frames at symbol rate rs on a GPU of nsm SMs (default configuration, exact arithmetic).  Touches no device: the
// host-side tests pin the policy's decisions for the BASELINE shapes with it.
extern "C" int qpsk_b200_debug_plan(int nsm, int sm_clock_khz, int nchan, int nframes, double rs, int* frame_chunks, int* frame_blocks, int* loop_mode) {
    if (nsm < 1 || nchan < 1 || nframes < 1) return fail(QPSK_B200_ERR_ARG, "nsm, nchan and nframes must be positive");
    const int sps = (int)(9600.0 / rs);
    if (sps != 4 && sps != 8) return fail(QPSK_B200_ERR_ARG, "samples/symbol %d unsupported (4 = 2400 baud, 8 = 1200 baud)", sps);
    qpsk_b200_rx* rx = new (std::nothrow) qpsk_b200_rx();
    if (!rx) return fail(QPSK_B200_ERR_ARG, "out of host memory");
    memset(rx, 0, sizeof *rx);
    rx->cfg.mode = QPSK_B200_MODE_EXACT;
    rx->C = nchan; rx->Cpad = (nchan + QPSK_GROUP - 1) / QPSK_GROUP * QPSK_GROUP; rx->maxF = nframes; rx->N = 512; rx->sps = sps; rx->nsym = 512 / sps;
    rx->nsm = nsm; rx->sm_clock_khz = sm_clock_khz > 0 ? sm_clock_khz : 1965000;
    rx->front_v1 = true; rx->relay_mode = 1; rx->chunk_div = QPSK_MAX_CHUNKS; rx->chase_smem = 120 * 1024;
    rx->block_progress_n = (size_t)(rx->Cpad / QPSK_GROUP) * nframes + 1;
    rx->d_block_progress = reinterpret_cast<unsigned long long*>(rx);      // only ever tested against null here
    const int ngroups = rx->Cpad / QPSK_GROUP, F = nframes;
    const int fc = rx_plan_chunks(rx, nchan, F);
    int chunks = 1, blocks = 1, mode = QPSK_B200_LOOP_FUSED;
    if (fc < F) {
        chunks = (F + fc - 1) / fc;
        blocks = rx_frame_blocks(rx, ngroups, nchan, fc, true);
        mode = rx_chase_plan(rx, F, fc, false) ? QPSK_B200_LOOP_CHASING : blocks == 1 ? QPSK_B200_LOOP_FUSED : QPSK_B200_LOOP_STANDALONE;
    } else if (rx_relay_blocks(rx, ngroups, F) > 1) {
        blocks = rx_relay_blocks(rx, ngroups, F);
        const int fpb = (F + blocks - 1) / blocks;
        blocks = (F + fpb - 1) / fpb;
        mode = QPSK_B200_LOOP_RELAYED;
    } else {
        blocks = rx_frame_blocks(rx, ngroups, nchan, F, false);
        const int fpb = (F + blocks - 1) / blocks;
        blocks = (F + fpb - 1) / fpb;
        mode = blocks == 1 ? QPSK_B200_LOOP_FUSED : QPSK_B200_LOOP_STANDALONE;
    }
    delete rx;
    if (frame_chunks) *frame_chunks = chunks;
    if (frame_blocks) *frame_blocks = blocks;
    if (loop_mode) *loop_mode = mode;
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_rx_last_plan(const qpsk_b200_rx* rx, int* frame_chunks, int* frame_blocks, int* loop_mode) {
    if (!rx) return fail(QPSK_B200_ERR_ARG, "null receiver");
    if (frame_chunks) *frame_chunks = rx->plan_chunks;
    if (frame_blocks) *frame_blocks = rx->plan_fblocks;
    if (loop_mode) *loop_mode = rx->plan_loop;
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_rx_process_device(qpsk_b200_rx* rx, const int16_t* d_pcm, int nframes, void* cuda_stream) {
    if (!rx || !d_pcm) return fail(QPSK_B200_ERR_ARG, "null argument");
    if (nframes < 1 || nframes > rx->maxF) return fail(QPSK_B200_ERR_ARG, "nframes %d outside 1..%d", nframes, rx->maxF);
    if ((reinterpret_cast<uintptr_t>(d_pcm) & 15) != 0) return fail(QPSK_B200_ERR_ARG, "d_pcm must be 16-byte aligned");
    if (rx->inflight > 0) return fail(QPSK_B200_ERR_STATE, "%d submitted host call(s) not waited for (qpsk_b200_rx_wait)", rx->inflight);
    CU(cudaSetDevice(rx->cfg.device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : rx->stream;
    return rx_run_call(rx, d_pcm, (size_t)nframes * rx->N, nframes, s);
}

extern "C" int qpsk_b200_rx_set_loop(qpsk_b200_rx* rx, float alpha, float beta, float min_freq, float max_freq) {
    if (!rx) return fail(QPSK_B200_ERR_ARG, "null receiver");
    rx->loop.alpha = alpha; rx->loop.beta = beta; rx->loop.min_freq = min_freq; rx->loop.max_freq = max_freq;
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_rx_get_loop_state(qpsk_b200_rx* rx, float* h_phase_freq) {
    if (!rx || !h_phase_freq) return fail(QPSK_B200_ERR_ARG, "null argument");
    CU(cudaSetDevice(rx->cfg.device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(h_phase_freq, rx->d_loop_state, (size_t)rx->C * sizeof(float2), cudaMemcpyDeviceToHost));
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_rx_set_loop_state(qpsk_b200_rx* rx, const float* h_phase_freq) {
    if (!rx || !h_phase_freq) return fail(QPSK_B200_ERR_ARG, "null argument");
    CU(cudaSetDevice(rx->cfg.device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(rx->d_loop_state, h_phase_freq, (size_t)rx->C * sizeof(float2), cudaMemcpyHostToDevice));
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_rx_sync(qpsk_b200_rx* rx) {
    if (!rx) return fail(QPSK_B200_ERR_ARG, "null receiver");
    CU(cudaSetDevice(rx->cfg.device));
    CU(cudaDeviceSynchronize());
    int dog = 0;
    CU(cudaMemcpy(&dog, rx->d_chunk_flags + QPSK_MAX_CHUNKS, sizeof dog, cudaMemcpyDeviceToHost));
    if (dog != 0) return fail(QPSK_B200_ERR_STATE, "the chasing loop kernel of call ticket %d gave up waiting for a frame chunk: results of that call are incomplete", dog);
    return QPSK_B200_OK;
}

extern "C" long long qpsk_b200_rx_launch_count(const qpsk_b200_rx* rx) { return rx ? rx->launches : 0; }

extern "C" int qpsk_b200_rx_last_kernel_ms(qpsk_b200_rx* rx, float* front_ms, float* costas_ms) {
    if (!rx || !rx->timed) return fail(QPSK_B200_ERR_STATE, "no device-resident process call yet");
    CU(cudaSetDevice(rx->cfg.device));
    CU(cudaDeviceSynchronize());
    float fr = 0.0f, lp = 0.0f;
    for (int k = 0; k < rx->timed_chunks; k++) {
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, rx->ev_fr[2 * k], rx->ev_fr[2 * k + 1]));
        fr += ms;
        if (rx->timed_loop && k < rx->timed_loop_chunks) { CU(cudaEventElapsedTime(&ms, rx->ev_lp[2 * k], rx->ev_lp[2 * k + 1])); lp += ms; }
    }
    if (front_ms) *front_ms = fr;
    if (costas_ms) *costas_ms = lp;
    return QPSK_B200_OK;
}

extern "C" size_t qpsk_b200_rx_output_bytes(const qpsk_b200_rx* rx, int what) {
    if (!rx) return 0;
    const size_t C = rx->C, F = rx->lastF, S = rx->nsym, N = rx->N;
    switch (what) {
        case QPSK_B200_OUT_DIBITS: return C * F * (S / 4);
        case QPSK_B200_OUT_INDEX: return C * F * sizeof(int);
        case QPSK_B200_OUT_TRACK: return C * F * sizeof(float2);
        case QPSK_B200_OUT_DEC: return C * F * S * sizeof(float2);
        case QPSK_B200_OUT_SYMBOLS: return C * F * S * sizeof(float2);
        case QPSK_B200_OUT_FIR: return C * F * N * sizeof(float2);
        case QPSK_B200_OUT_TAPS: return (size_t)rx->cfg.ntaps * sizeof(float);
        case QPSK_B200_OUT_FRAMES: return C * F * (S / 4);
        case QPSK_B200_OUT_CRC_OK: return C * F;
        case QPSK_B200_OUT_ROTATION: return C * F;
        case QPSK_B200_OUT_TIMING_SUM: return C * F * sizeof(float2);
        case QPSK_B200_OUT_TIMING_TAU: return C * F * sizeof(float);
        case QPSK_B200_OUT_OFFSET_BIN: return C * sizeof(int);
        case QPSK_B200_OUT_OFFSET_HZ: return C * sizeof(float);
        default: return 0;
    }
}

static int ensure_scratch(qpsk_b200_rx* rx, size_t bytes) {
    if (rx->scratch_bytes >= bytes) return 0;
    if (rx->d_scratch) { cudaFree(rx->d_scratch); rx->d_scratch = nullptr; rx->scratch_bytes = 0; }
    CU(cudaMalloc(&rx->d_scratch, bytes));
    rx->scratch_bytes = bytes;
    return 0;
}

template <typename T>
static int download_transposed(qpsk_b200_rx* rx, const T* d_src, int rows, void* h_dst, cudaStream_t s) {
    const size_t bytes = (size_t)rx->C * rows * sizeof(T);
    int rc = ensure_scratch(rx, bytes);
    if (rc) return rc;
    dim3 grid((rx->C + 31) / 32, (rows + 31) / 32), block(32, 8);
    transpose_to_channel_major<T><<<grid, block, 0, s>>>(d_src, (T*)rx->d_scratch, rows, rx->C, rx->Cpad);
    CU(cudaGetLastError());
    rx->launches += 1;
    CU(cudaMemcpyAsync(h_dst, rx->d_scratch, bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return 0;
}

extern "C" int qpsk_b200_rx_read(qpsk_b200_rx* rx, int what, void* h_dst, size_t bytes) {
    if (!rx || !h_dst) return fail(QPSK_B200_ERR_ARG, "null argument");
    if (what != QPSK_B200_OUT_TAPS && rx->lastF == 0) return fail(QPSK_B200_ERR_STATE, "no process call yet");
    const size_t need = qpsk_b200_rx_output_bytes(rx, what);
    if (need == 0 || bytes != need) return fail(QPSK_B200_ERR_ARG, "output %d needs %zu bytes, got %zu", what, need, bytes);
    CU(cudaSetDevice(rx->cfg.device));
    CU(cudaDeviceSynchronize());
    cudaStream_t s = rx->stream;
    const int F = rx->lastF, S = rx->nsym;
    switch (what) {
        case QPSK_B200_OUT_TAPS: memcpy(h_dst, rx->taps, need); return 0;
        case QPSK_B200_OUT_DIBITS: return download_transposed<unsigned>(rx, rx->d_dibits_t, F * (S / 16), h_dst, s);
        case QPSK_B200_OUT_INDEX: return download_transposed<int>(rx, rx->d_index_t, F, h_dst, s);
        case QPSK_B200_OUT_TRACK: return download_transposed<float2>(rx, rx->d_track_t, F, h_dst, s);
        case QPSK_B200_OUT_FRAMES:
            if (!rx->d_frames_t) return fail(QPSK_B200_ERR_STATE, "frames were not decoded (QPSK_B200_DECODE_FRAMES)");
            return download_transposed<unsigned>(rx, rx->d_frames_t, F * (S / 16), h_dst, s);
        case QPSK_B200_OUT_CRC_OK:
            if (!rx->d_crc_ok_t) return fail(QPSK_B200_ERR_STATE, "frames were not decoded (QPSK_B200_DECODE_FRAMES)");
            return download_transposed<uint8_t>(rx, rx->d_crc_ok_t, F, h_dst, s);
        case QPSK_B200_OUT_ROTATION:
            if (!rx->d_rotation_t) return fail(QPSK_B200_ERR_STATE, "rotations were not resolved (QPSK_B200_DECODE_FRAMES | QPSK_B200_RESOLVE_ROTATION)");
            return download_transposed<uint8_t>(rx, rx->d_rotation_t, F, h_dst, s);
        case QPSK_B200_OUT_TIMING_SUM:
        case QPSK_B200_OUT_TIMING_TAU: {
            if (!rx->d_timing_t) return fail(QPSK_B200_ERR_STATE, "the timing statistic was not accumulated (QPSK_B200_ESTIMATE_TIMING)");
            if (what == QPSK_B200_OUT_TIMING_SUM) return download_transposed<float2>(rx, rx->d_timing_t, F, h_dst, s);
            float2* tmp = new (std::nothrow) float2[(size_t)rx->C * F];
            if (!tmp) return fail(QPSK_B200_ERR_ARG, "out of host memory");
            int rc = download_transposed<float2>(rx, rx->d_timing_t, F, tmp, s);
            if (rc == 0) {
                float* tau = static_cast<float*>(h_dst);
                const double sps = (double)rx->sps;
                for (size_t i = 0; i < (size_t)rx->C * F; i++) {
                    double t = -atan2((double)tmp[i].y, (double)tmp[i].x) * sps / kTau;      // tau = -arg(S) sps / (2 pi)
                    if (t < 0.0) t += sps;
                    tau[i] = (float)t;
                }
            }
            delete[] tmp;
            return rc;
        }
        case QPSK_B200_OUT_OFFSET_BIN:
        case QPSK_B200_OUT_OFFSET_HZ: {
            if (!rx->est_on) return fail(QPSK_B200_ERR_STATE, "the estimator did not run (QPSK_B200_ESTIMATE_OFFSET)");
            CU(cudaMemcpy(h_dst, rx->d_est_bins, (size_t)rx->C * sizeof(int), cudaMemcpyDeviceToHost));
            if (what == QPSK_B200_OUT_OFFSET_HZ) {
                const int n = rx->est_call_n;
                int32_t* b = static_cast<int32_t*>(h_dst);
                float* hz = static_cast<float*>(h_dst);
                for (int c = 0; c < rx->C; c++) {
                    const int k = b[c] < n / 2 ? b[c] : b[c] - n;             // signed bin of the 4x tone
                    hz[c] = (float)((double)k * (double)rx->cfg.rs / (4.0 * (double)n));
                }
            }
            return 0;
        }
        case QPSK_B200_OUT_SYMBOLS:
            if (!rx->d_costas_dbg) return fail(QPSK_B200_ERR_STATE, "symbols were not kept (QPSK_B200_KEEP_SYMBOLS)");
            return download_transposed<float2>(rx, rx->d_costas_dbg, F * S, h_dst, s);
        case QPSK_B200_OUT_FIR:
            if (!rx->d_fir_dbg) return fail(QPSK_B200_ERR_STATE, "filter output was not kept (QPSK_B200_KEEP_FIR)");
            CU(cudaMemcpy(h_dst, rx->d_fir_dbg, need, cudaMemcpyDeviceToHost));
            return 0;
        case QPSK_B200_OUT_DEC: {
            if (rx->cfg.flags & QPSK_B200_TRANSIENT_SYMBOLS) return fail(QPSK_B200_ERR_STATE, "the decimated symbols were not kept (QPSK_B200_TRANSIENT_SYMBOLS)");
            // frames of the last call live in ring slots (slot_base_before + 1 + f); slot_base already advanced by F
            const int base_before = ((rx->slot_base - F) % rx->nslots + rx->nslots) % rx->nslots;
            int rc = ensure_scratch(rx, need + (size_t)F * S * rx->Cpad * sizeof(float2));
            if (rc) return rc;
            float2* lin = reinterpret_cast<float2*>(static_cast<char*>(rx->d_scratch) + need);
            for (int f = 0; f < F; f++) {
                const int slot = (base_before + 1 + f) % rx->nslots;
                CU(cudaMemcpyAsync(lin + (size_t)f * S * rx->Cpad, rx->d_dec_ring + (size_t)slot * S * rx->Cpad,
                                   (size_t)S * rx->Cpad * sizeof(float2), cudaMemcpyDeviceToDevice, s));
            }
            dim3 grid((rx->C + 31) / 32, (F * S + 31) / 32), block(32, 8);
            transpose_to_channel_major<float2><<<grid, block, 0, s>>>(lin, (float2*)rx->d_scratch, F * S, rx->C, rx->Cpad);
            CU(cudaGetLastError());
            rx->launches += 1;
            CU(cudaMemcpyAsync(h_dst, rx->d_scratch, need, cudaMemcpyDeviceToHost, s));
            CU(cudaStreamSynchronize(s));
            return 0;
        }
        default: return fail(QPSK_B200_ERR_ARG, "unknown output %d", what);
    }
}

extern "C" int qpsk_b200_rx_device_dibits(qpsk_b200_rx* rx, const uint32_t** d_ptr, int* cpad) {
    if (!rx || !d_ptr) return fail(QPSK_B200_ERR_ARG, "null argument");
    *d_ptr = rx->d_dibits_t;
    if (cpad) *cpad = rx->Cpad;
    return QPSK_B200_OK;
}

// Host-buffer entry points.  A call is cut into jobs -- channel slices when channels are plentiful (each slice keeps
// the fused kernel busy for several waves), frame chunks when they are few (rx_plan_chunks) -- that flow through the
// copy-in stream, the compute stream(s) and the copy-out stream with double-buffered staging, so the PCIe transfers of
// one job overlap the kernels of its neighbours.  qpsk_b200_rx_submit_host only enqueues (up to two calls may be in
// flight: reading the next batch of PCM overlaps the GPU working on this one); qpsk_b200_rx_wait completes the oldest.
static int rx_host_streams(qpsk_b200_rx* rx) {
    if (rx->s_in) return 0;
    CU(cudaStreamCreateWithFlags(&rx->s_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&rx->s_out, cudaStreamNonBlocking));
    for (int b = 0; b < 2; b++) {
        CU(cudaEventCreateWithFlags(&rx->ev_in[b], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&rx->ev_cmp[b], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&rx->ev_res[b], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&rx->ev_out[b], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&rx->ev_done[b], cudaEventDisableTiming));
    }
    return 0;
}

// after an error in the middle of a host call: nothing may still read or write the caller's buffers on return
static int rx_host_abort(qpsk_b200_rx* rx, int rc) {
    cudaStreamSynchronize(rx->s_in);
    cudaStreamSynchronize(rx->stream);
    cudaStreamSynchronize(rx->s_loop);
    cudaStreamSynchronize(rx->s_out);
    rx->inflight = 0;
    rx->no_fuse = (rx->cfg.flags & QPSK_B200_NO_FUSE) != 0;      // a seeding call (PREROTATE_OFFSET) may have been cut short
    rx->needs_reset = true;      // some jobs of the call ran, others did not: channel state is inconsistent
    return rc;
}

static int rx_submit_host(qpsk_b200_rx* rx, const int16_t* h_pcm, int nframes, uint8_t* h_dibits, bool copy_only) {
    if (!rx || !h_pcm) return fail(QPSK_B200_ERR_ARG, "null argument");
    if (nframes < 1 || nframes > rx->maxF) return fail(QPSK_B200_ERR_ARG, "nframes %d outside 1..%d", nframes, rx->maxF);
    if (rx->needs_reset) return fail(QPSK_B200_ERR_STATE, "an earlier call failed half way: qpsk_b200_rx_reset() first");
    if (rx->inflight >= 2) return fail(QPSK_B200_ERR_STATE, "two calls already in flight: qpsk_b200_rx_wait() first");
    CU(cudaSetDevice(rx->cfg.device));
    int rc = rx_host_streams(rx);
    if (rc) return rc;
    const int F = nframes, N = rx->N, C = rx->C, W = rx->nsym / 16;
    const size_t row_bytes = (size_t)F * N * sizeof(int16_t);
    // channel slices: ~256 MiB of PCM, whole 32-channel groups, at least 4 waves of fused CTAs (2 per SM) when possible
    // (4 half waves = 2 waves of whole-call CTAs; when the call will also be cut into frame chunks, below, HALF a wave per job:
    // one CTA per SM runs a frame in ~50 us instead of 83, a job's compute always fits under the next job's copy, and nothing
    // queues up behind the last copy but the last, small job itself -- 80.26 / 79.44 / 79.14 ms per step for 4 / 2 / 1 half waves)
    const bool chunk_ok = rx->host_tail_chunks > 1 && F >= 32 && !rx->d_fir_dbg && !rx->d_costas_dbg && !rx->no_chunk && !(rx->prerotate && !rx->loop_seeded);
    int slice = (int)((256ull << 20) / row_bytes) / QPSK_GROUP * QPSK_GROUP;
    const int slice_floor = (chunk_ok ? rx->host_slice_halfwaves : 4) * 2 * rx->nsm * QPSK_GROUP / 2;
    if (slice < slice_floor) slice = slice_floor;
    if (slice > C) slice = C;
    const int nslices = (C + slice - 1) / slice;
    // Frame chunks.  One slice (few channels): the device-resident plan, the loop of chunk k apart from and under the front end
    // of chunk k + 1.  Several slices: the call is bound by the PCM arriving, and what it adds to the last byte's arrival is the
    // compute of the LAST job -- a quarter of the frames per job makes that job a quarter as long (65,536 channels x 64 frames:
    // 5.5 ms of whole-stream CTAs behind the last copy become 1.4 ms); the loop stays fused, state carries as between calls.
    int fc = F;
    // (a host call's chunks are copies too: 16 at most -- with 32, configs[1]'s 8 KB / 128 B rows in and out cost 0.3 ms of 6.0)
    if (nslices == 1) fc = rx_plan_chunks(rx, C, F, rx->chunk_div < 16 ? rx->chunk_div : 16);
    else if (chunk_ok)
        fc = (F + rx->host_tail_chunks - 1) / rx->host_tail_chunks;
    const bool chunked = fc < F;
    const bool loop_apart = chunked && nslices == 1;
    // staging for one job: slice x fc frames of PCM in, slice x fc frames of packed dibits out
    const size_t pcm_job = (size_t)slice * fc * N * sizeof(int16_t), out_job = (size_t)slice * fc * W * sizeof(unsigned);
    if (rx->stage_pcm_bytes < pcm_job || rx->stage_out_bytes < out_job) {
        CU(cudaDeviceSynchronize());
        for (int b = 0; b < 2; b++) {
            if (rx->d_pcm_stage2[b]) { cudaFree(rx->d_pcm_stage2[b]); rx->d_pcm_stage2[b] = nullptr; }
            if (rx->d_out_stage2[b]) { cudaFree(rx->d_out_stage2[b]); rx->d_out_stage2[b] = nullptr; }
        }
        rx->stage_pcm_bytes = rx->stage_out_bytes = 0;
        for (int b = 0; b < 2; b++) {
            CU(cudaMalloc((void**)&rx->d_pcm_stage2[b], pcm_job));
            CU(cudaMalloc((void**)&rx->d_out_stage2[b], out_job));
        }
        rx->stage_pcm_bytes = pcm_job; rx->stage_out_bytes = out_job;
    }
    cudaStream_t sc = rx->stream;
    const int first_slot = (rx->slot_base + 1) % rx->nslots;
    const bool seeding = rx->prerotate && !rx->loop_seeded && !copy_only;      // see rx_run_call
    if (!copy_only) rx_plan_est_emit(rx, F, seeding);
    const bool saved_no_fuse = rx->no_fuse;
    if (seeding) rx->no_fuse = true;
    bool loop_stream_used = false, est_done = false;
    for (int f0 = 0; f0 < F; f0 += fc) {
        const int Fj = (F - f0 < fc) ? F - f0 : fc;
        if (!copy_only) {
            rc = rx_begin_chunk(rx, Fj, sc);
            if (rc) return rx_host_abort(rx, rc);
        }
        for (int i = 0; i < nslices; i++) {
            const unsigned long long seq = rx->job_seq++;
            const int b = (int)(seq & 1), c0 = i * slice, nc = (c0 + slice <= C) ? slice : C - c0;
            // PCM staging buffer b was last read by the front end of job seq-2
            if (seq >= 2) CU(cudaStreamWaitEvent(rx->s_in, rx->ev_cmp[b], 0));
            cudaError_t e;
            if (!chunked) e = cudaMemcpyAsync(rx->d_pcm_stage2[b], h_pcm + (size_t)c0 * F * N, (size_t)nc * row_bytes, cudaMemcpyHostToDevice, rx->s_in);
            else e = cudaMemcpy2DAsync(rx->d_pcm_stage2[b], (size_t)Fj * N * 2, h_pcm + (size_t)c0 * F * N + (size_t)f0 * N, row_bytes,
                                       (size_t)Fj * N * 2, nc, cudaMemcpyHostToDevice, rx->s_in);
            if (e != cudaSuccess) return rx_host_abort(rx, fail(QPSK_B200_ERR_CUDA, "PCM upload failed: %s", cudaGetErrorString(e)));
            CU(cudaEventRecord(rx->ev_in[b], rx->s_in));
            CU(cudaStreamWaitEvent(sc, rx->ev_in[b], 0));
            RxJob j;
            j.d_pcm = rx->d_pcm_stage2[b]; j.pcm_row = (size_t)Fj * N; j.c0 = c0; j.nc = nc; j.F = Fj; j.f_off = f0;
            cudaStream_t sr = sc;                               // the stream the job's results appear on
            if (!copy_only) {
                bool fused = false;
                rc = rx_launch_front(rx, j, loop_apart, sc, &fused);
                if (!rc && seeding) rc = rx_seed_loop(rx, c0, (c0 + nc < C) ? c0 + nc : C, F, first_slot, sc);
                if (rc) { rx->no_fuse = saved_no_fuse; return rx_host_abort(rx, rc); }
                CU(cudaEventRecord(rx->ev_cmp[b], sc));         // the PCM staging buffer is free again
                if (loop_apart && !fused) {
                    sr = rx->s_loop;
                    CU(cudaStreamWaitEvent(sr, rx->ev_cmp[b], 0));
                    loop_stream_used = true;
                }
                rc = rx_launch_loop_and_decode(rx, j, fused, sr);
                if (rc) return rx_host_abort(rx, rc);
            } else {
                CU(cudaEventRecord(rx->ev_cmp[b], sc));
            }
            if (h_dibits) {
                // output staging buffer b was last drained by the D2H copy of job seq-2
                if (seq >= 2) CU(cudaStreamWaitEvent(sr, rx->ev_out[b], 0));
                const int words = Fj * W;
                if (!copy_only) {
                    dim3 grid((nc + 31) / 32, (words + 31) / 32), block(32, 8);
                    transpose_to_channel_major<unsigned><<<grid, block, 0, sr>>>(rx->d_dibits_t + (size_t)f0 * W * rx->Cpad + c0, rx->d_out_stage2[b], words, nc, rx->Cpad);
                    CU(cudaGetLastError());
                    rx->launches += 1;
                }
                CU(cudaEventRecord(rx->ev_res[b], sr));
                CU(cudaStreamWaitEvent(rx->s_out, rx->ev_res[b], 0));
                uint8_t* dst = h_dibits + ((size_t)c0 * F * W + (size_t)f0 * W) * 4;
                if (!chunked) e = cudaMemcpyAsync(dst, rx->d_out_stage2[b], (size_t)nc * words * 4, cudaMemcpyDeviceToHost, rx->s_out);
                else e = cudaMemcpy2DAsync(dst, (size_t)F * W * 4, rx->d_out_stage2[b], (size_t)words * 4, (size_t)words * 4, nc, cudaMemcpyDeviceToHost, rx->s_out);
                if (e != cudaSuccess) return rx_host_abort(rx, fail(QPSK_B200_ERR_CUDA, "dibit download failed: %s", cudaGetErrorString(e)));
                CU(cudaEventRecord(rx->ev_out[b], rx->s_out));
            }
        }
        if (!copy_only) rx_end_chunk(rx, Fj);
        // the estimator reads the call's first symbols only: in a multi-slice call it runs as soon as the chunk that holds them
        // is through (every slice's loop included: fused, same stream), not behind the last copy where the caller waits for it
        if (rx->est_on && !copy_only && !seeding && !est_done && chunked && !loop_apart
            && (long long)(f0 + Fj) * rx->nsym >= est_burst_length(rx, F) + rx->nsym) {
            rc = rx_launch_estimator(rx, 0, C, F, first_slot, sc);
            if (rc) return rx_host_abort(rx, rc);
            est_done = true;
        }
    }
    if (loop_stream_used) {                                     // the compute stream joins the loop stream at the end of the call
        CU(cudaEventRecord(rx->ev_front, rx->s_loop));
        CU(cudaStreamWaitEvent(sc, rx->ev_front, 0));
    }
    if (seeding) { rx->no_fuse = saved_no_fuse; rx->loop_seeded = true; }
    if (rx->est_on && !copy_only && !seeding && !est_done) {
        rc = rx_launch_estimator(rx, 0, C, F, first_slot, sc);
        if (rc) return rx_host_abort(rx, rc);
    }
    // completion of this call = its last compute work and its last copy out
    const int slot = (int)(rx->call_seq++ & 1);
    CU(cudaEventRecord(rx->ev_front, sc));
    CU(cudaStreamWaitEvent(rx->s_out, rx->ev_front, 0));
    CU(cudaEventRecord(rx->ev_done[slot], rx->s_out));
    rx->inflight += 1;
    if (!copy_only) rx->lastF = F;
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_rx_submit_host(qpsk_b200_rx* rx, const int16_t* h_pcm, int nframes, uint8_t* h_dibits) {
    return rx_submit_host(rx, h_pcm, nframes, h_dibits, false);
}

extern "C" int qpsk_b200_rx_wait(qpsk_b200_rx* rx) {
    if (!rx) return fail(QPSK_B200_ERR_ARG, "null receiver");
    if (rx->inflight <= 0) return QPSK_B200_OK;
    CU(cudaSetDevice(rx->cfg.device));
    const int slot = (int)((rx->call_seq - (unsigned long long)rx->inflight) & 1);      // the oldest call in flight
    cudaError_t e = cudaEventSynchronize(rx->ev_done[slot]);
    if (e != cudaSuccess) return rx_host_abort(rx, fail(QPSK_B200_ERR_CUDA, "waiting for a submitted call failed: %s", cudaGetErrorString(e)));
    rx->inflight -= 1;
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_rx_process_host(qpsk_b200_rx* rx, const int16_t* h_pcm, int nframes, uint8_t* h_dibits) {
    if (rx && rx->inflight > 0) return fail(QPSK_B200_ERR_STATE, "%d submitted call(s) not waited for (qpsk_b200_rx_wait)", rx->inflight);
    int rc = rx_submit_host(rx, h_pcm, nframes, h_dibits, false);
    if (rc) return rc;
    return qpsk_b200_rx_wait(rx);
}

// Measurement aid: exactly the host<->device traffic of qpsk_b200_rx_process_host -- same slices, same streams, same
// events -- with no kernel launched, so the end-to-end rate can be quoted against the copy ceiling of the same run.
extern "C" int qpsk_b200_rx_probe_copy_host(qpsk_b200_rx* rx, const int16_t* h_pcm, int nframes, uint8_t* h_scratch_out) {
    if (rx && rx->inflight > 0) return fail(QPSK_B200_ERR_STATE, "%d submitted call(s) not waited for (qpsk_b200_rx_wait)", rx->inflight);
    int rc = rx_submit_host(rx, h_pcm, nframes, h_scratch_out, true);
    if (rc) return rc;
    return qpsk_b200_rx_wait(rx);
}

// =============================================================================================
// channel-batched rrc_fir
// =============================================================================================
struct qpsk_b200_fir {
    long long id;
    int ntaps, C, mode, device;
    float taps[QPSK_MAX_TAPS];
    float2* d_state;      // [C][ntaps]
    float2* d_stage;      // host-path staging (lazy)
    size_t stage_elems;
    float2* d_stage2[2];  // sliced host path: two slices in flight (lazy)
    size_t stage2_elems;
    cudaStream_t s_in, s_out;
    cudaEvent_t ev_in[2], ev_k[2], ev_out[2];
    float2* d_halo;       // inputs in front of time blocks 1.. (lazy), see fir_save_halo_kernel
    size_t halo_elems;
    int sm_count;
    cudaStream_t stream;
    cudaEvent_t ev[2];
    bool timed;
};

extern "C" int qpsk_b200_rrc_make(float* taps, int ntaps, float fs, float rs, float alpha) {
    if (!taps || ntaps < 1 || ntaps > QPSK_MAX_TAPS) return fail(QPSK_B200_ERR_ARG, "ntaps must be 1..%d", QPSK_MAX_TAPS);
    qpsk_host_rrc_make(taps, ntaps, fs, rs, alpha);
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_fir_destroy(qpsk_b200_fir* f) {
    if (!f) return 0;
    cudaSetDevice(f->device);
    if (f->d_state) cudaFree(f->d_state);
    if (f->d_stage) cudaFree(f->d_stage);
    for (auto& p : f->d_stage2) if (p) cudaFree(p);
    if (f->s_in) cudaStreamDestroy(f->s_in);
    if (f->s_out) cudaStreamDestroy(f->s_out);
    for (int b = 0; b < 2; b++) {
        if (f->ev_in[b]) cudaEventDestroy(f->ev_in[b]);
        if (f->ev_k[b]) cudaEventDestroy(f->ev_k[b]);
        if (f->ev_out[b]) cudaEventDestroy(f->ev_out[b]);
    }
    if (f->d_halo) cudaFree(f->d_halo);
    for (auto& e : f->ev) if (e) cudaEventDestroy(e);
    if (f->stream) cudaStreamDestroy(f->stream);
    delete f;
    return 0;
}

static int check_device(int device) {
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(QPSK_B200_ERR_CUDA, "CUDA device %d not present (%d devices)", device, ndev);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(QPSK_B200_ERR_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
    return 0;
}

extern "C" int qpsk_b200_fir_create(const float* taps, int ntaps, int nchan, int mode, int device, qpsk_b200_fir** out) {
    if (!taps || !out) return fail(QPSK_B200_ERR_ARG, "null argument");
    *out = nullptr;
    if (ntaps != 127 && ntaps != 256) return fail(QPSK_B200_ERR_ARG, "ntaps %d unsupported: kernels are instantiated for 127 (rrc_fir.h:13) and 256 (long-tap profile)", ntaps);
    if (nchan < 1) return fail(QPSK_B200_ERR_ARG, "nchan must be positive");
    if (mode != QPSK_B200_MODE_EXACT && mode != QPSK_B200_MODE_FAST) return fail(QPSK_B200_ERR_ARG, "bad mode");
    int rc = check_device(device);
    if (rc) return rc;
    qpsk_b200_fir* f = new (std::nothrow) qpsk_b200_fir();
    if (!f) return fail(QPSK_B200_ERR_ARG, "out of host memory");
    memset(f, 0, sizeof *f);
    f->id = g_next_id++;
    f->ntaps = ntaps; f->C = nchan; f->mode = mode; f->device = device;
    memcpy(f->taps, taps, sizeof(float) * ntaps);
    cudaError_t e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
    for (auto& ev : f->ev) if (e == cudaSuccess) e = cudaEventCreate(&ev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&f->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_state, (size_t)nchan * ntaps * sizeof(float2));
    if (e == cudaSuccess) e = cudaMemset(f->d_state, 0, (size_t)nchan * ntaps * sizeof(float2));
    if (e != cudaSuccess) { qpsk_b200_fir_destroy(f); return fail(QPSK_B200_ERR_CUDA, "allocating FIR state failed: %s", cudaGetErrorString(e)); }
    *out = f;
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_fir_reset(qpsk_b200_fir* f) {
    if (!f) return fail(QPSK_B200_ERR_ARG, "null filter");
    CU(cudaSetDevice(f->device));
    CU(cudaMemsetAsync(f->d_state, 0, (size_t)f->C * f->ntaps * sizeof(float2), f->stream));
    CU(cudaStreamSynchronize(f->stream));
    return QPSK_B200_OK;
}

// Cut the call into time blocks so that the grid is (nearly) a whole number of waves: 16,384 channels are 512 CTAs,
// 1.73 waves of 2 x 148 -- 13 % of the machine idle in the tail -- while 4 time blocks make it 6.92 waves.
static void fir_time_blocks(int groups, int ntiles, int halo_tiles, int slots, int* nblocks, int* tiles_per_block) {
    int best_nb = 1, best_tpb = ntiles;
    double best_eff = 0.0;
    for (int nb = 1; nb <= 64; nb++) {
        const int tpb = (ntiles + nb - 1) / nb;
        if (nb > 1 && tpb < halo_tiles + 2) break;
        const int nbe = (ntiles + tpb - 1) / tpb;
        const long long total = (long long)groups * nbe;
        const long long waves = (total + slots - 1) / slots;
        const double eff = (double)total / (double)(waves * slots);
        if (eff > best_eff + 0.005) { best_eff = eff; best_nb = nbe; best_tpb = tpb; }
        if (best_eff >= 0.97) break;
    }
    *nblocks = best_nb; *tiles_per_block = best_tpb;
}

template <int NTAPS, int MODE>
static cudaError_t launch_fir(qpsk_b200_fir* f, FirArgs a, cudaStream_t s) {
    constexpr int NW = NTAPS > 128 ? 16 : 8;        // filter warps per CTA, see FirSmem
    using Smem = FirSmem<NTAPS, NW>;
    const size_t smem = sizeof(Smem);
    cudaError_t e = cudaFuncSetAttribute(fir_kernel<NTAPS, MODE, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fir_kernel<NTAPS, MODE, NW>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    const int groups = (a.C + QPSK_GROUP - 1) / QPSK_GROUP, ntiles = (a.T + Smem::TILE - 1) / Smem::TILE;
    const int per_sm = smem <= 112 * 1024 ? 2 : 1;
    fir_time_blocks(groups, ntiles, Smem::HT, f->sm_count * per_sm, &a.nblocks, &a.tiles_per_block);
    if (a.nblocks > 1) {
        // one halo slot per time block: slot 0 is a copy of the delay line, so that no CTA of the main kernel reads
        // `state` while the CTA of the last time block rewrites it (block dispatch order is not a guarantee)
        const size_t need = (size_t)a.C * a.nblocks * (NTAPS - 1);
        if (f->halo_elems < need) {
            if (f->d_halo) { cudaFree(f->d_halo); f->d_halo = nullptr; f->halo_elems = 0; }
            e = cudaMalloc((void**)&f->d_halo, need * sizeof(float2));
            if (e != cudaSuccess) return e;
            f->halo_elems = need;
        }
        fir_save_halo_kernel<<<dim3(a.C, a.nblocks), 128, 0, s>>>(a.data, a.state, f->d_halo, a.T, NTAPS - 1, a.nblocks, a.tiles_per_block * Smem::TILE);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    a.halo = f->d_halo;
    fir_kernel<NTAPS, MODE, NW><<<dim3(groups, a.nblocks), 32 * NW, smem, s>>>(a, tap_bank<NTAPS>(f->taps));
    return cudaGetLastError();
}

// channels [c0, c0 + nc) of the filter bank over d_samples (exactly those rows), on stream s
static int fir_run(qpsk_b200_fir* f, float* d_samples, int c0, int nc, int nsamples, cudaStream_t s, bool timed);

extern "C" int qpsk_b200_fir_process_device(qpsk_b200_fir* f, float* d_samples, int nsamples, void* cuda_stream) {
    if (!f || !d_samples) return fail(QPSK_B200_ERR_ARG, "null argument");
    if (nsamples < 1) return fail(QPSK_B200_ERR_ARG, "nsamples must be positive");
    CU(cudaSetDevice(f->device));
    return fir_run(f, d_samples, 0, f->C, nsamples, cuda_stream ? (cudaStream_t)cuda_stream : f->stream, true);
}

static int fir_run(qpsk_b200_fir* f, float* d_samples, int c0, int nc, int nsamples, cudaStream_t s, bool timed) {
    FirArgs a;
    a.data = reinterpret_cast<float2*>(d_samples); a.state = f->d_state + (size_t)c0 * f->ntaps; a.C = nc; a.T = nsamples;
    if (timed) CU(cudaEventRecord(f->ev[0], s));
    cudaError_t e;
    const bool fast = f->mode == QPSK_B200_MODE_FAST;
    // exact mode of the general filter: the IEEE form (no flush to zero: rrc_fir takes any float input, rrc_fir.c:24-26)
    if (f->ntaps == 127) e = fast ? launch_fir<127, QPSK_MODE_FAST>(f, a, s) : launch_fir<127, QPSK_MODE_IEEE>(f, a, s);
    else                 e = fast ? launch_fir<256, QPSK_MODE_FAST>(f, a, s) : launch_fir<256, QPSK_MODE_IEEE>(f, a, s);
    if (e != cudaSuccess) return fail(QPSK_B200_ERR_CUDA, "FIR kernel launch failed: %s", cudaGetErrorString(e));
    if (timed) { CU(cudaEventRecord(f->ev[1], s)); f->timed = true; }
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_fir_process_host(qpsk_b200_fir* f, float* h_samples, int nsamples) {
    if (!f || !h_samples) return fail(QPSK_B200_ERR_ARG, "null argument");
    if (nsamples < 1) return fail(QPSK_B200_ERR_ARG, "nsamples must be positive");
    CU(cudaSetDevice(f->device));
    // Large banks go through in channel slices over three streams (copy in / filter / copy out, two slices in flight), like
    // the receiver's host path: with page-locked buffers the two PCIe directions and the kernel overlap.
    const size_t slice_target = (size_t)32 << 20;                       // complex samples per slice (256 MiB)
    if ((size_t)f->C * nsamples > 2 * slice_target && f->C >= 64) {
        int nc_slice = (int)(slice_target / (size_t)nsamples) / 32 * 32;
        if (nc_slice < 32) nc_slice = 32;
        const size_t se = (size_t)nc_slice * nsamples;
        if (!f->s_in) {
            CU(cudaStreamCreateWithFlags(&f->s_in, cudaStreamNonBlocking));
            CU(cudaStreamCreateWithFlags(&f->s_out, cudaStreamNonBlocking));
            for (int b = 0; b < 2; b++) {
                CU(cudaEventCreateWithFlags(&f->ev_in[b], cudaEventDisableTiming));
                CU(cudaEventCreateWithFlags(&f->ev_k[b], cudaEventDisableTiming));
                CU(cudaEventCreateWithFlags(&f->ev_out[b], cudaEventDisableTiming));
            }
        }
        if (f->stage2_elems < se) {
            for (auto& p : f->d_stage2) if (p) { cudaFree(p); p = nullptr; }
            f->stage2_elems = 0;
            for (auto& p : f->d_stage2) CU(cudaMalloc((void**)&p, se * sizeof(float2)));
            f->stage2_elems = se;
        }
        float2* h = reinterpret_cast<float2*>(h_samples);
        int i = 0;
        for (int c0 = 0; c0 < f->C; c0 += nc_slice, i++) {
            const int b = i & 1, nc = (f->C - c0 < nc_slice) ? f->C - c0 : nc_slice;
            const size_t bytes = (size_t)nc * nsamples * sizeof(float2);
            if (i >= 2) CU(cudaStreamWaitEvent(f->s_in, f->ev_out[b], 0));          // buffer b has been copied out
            CU(cudaMemcpyAsync(f->d_stage2[b], h + (size_t)c0 * nsamples, bytes, cudaMemcpyHostToDevice, f->s_in));
            CU(cudaEventRecord(f->ev_in[b], f->s_in));
            CU(cudaStreamWaitEvent(f->stream, f->ev_in[b], 0));
            int rc = fir_run(f, reinterpret_cast<float*>(f->d_stage2[b]), c0, nc, nsamples, f->stream, false);
            if (rc) return rc;
            CU(cudaEventRecord(f->ev_k[b], f->stream));
            CU(cudaStreamWaitEvent(f->s_out, f->ev_k[b], 0));
            CU(cudaMemcpyAsync(h + (size_t)c0 * nsamples, f->d_stage2[b], bytes, cudaMemcpyDeviceToHost, f->s_out));
            CU(cudaEventRecord(f->ev_out[b], f->s_out));
        }
        CU(cudaStreamSynchronize(f->s_out));
        CU(cudaStreamSynchronize(f->stream));
        return QPSK_B200_OK;
    }
    const size_t elems = (size_t)f->C * nsamples;
    if (f->stage_elems < elems) {
        if (f->d_stage) { cudaFree(f->d_stage); f->d_stage = nullptr; f->stage_elems = 0; }
        CU(cudaMalloc((void**)&f->d_stage, elems * sizeof(float2)));
        f->stage_elems = elems;
    }
    CU(cudaMemcpyAsync(f->d_stage, h_samples, elems * sizeof(float2), cudaMemcpyHostToDevice, f->stream));
    int rc = qpsk_b200_fir_process_device(f, reinterpret_cast<float*>(f->d_stage), nsamples, f->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h_samples, f->d_stage, elems * sizeof(float2), cudaMemcpyDeviceToHost, f->stream));
    CU(cudaStreamSynchronize(f->stream));
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_fir_get_memory(qpsk_b200_fir* f, float* h_memory) {
    if (!f || !h_memory) return fail(QPSK_B200_ERR_ARG, "null argument");
    CU(cudaSetDevice(f->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(h_memory, f->d_state, (size_t)f->C * f->ntaps * sizeof(float2), cudaMemcpyDeviceToHost));
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_fir_set_memory(qpsk_b200_fir* f, const float* h_memory) {
    if (!f || !h_memory) return fail(QPSK_B200_ERR_ARG, "null argument");
    CU(cudaSetDevice(f->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(f->d_state, h_memory, (size_t)f->C * f->ntaps * sizeof(float2), cudaMemcpyHostToDevice));
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_fir_last_kernel_ms(qpsk_b200_fir* f, float* ms) {
    if (!f || !ms || !f->timed) return fail(QPSK_B200_ERR_STATE, "no process call yet");
    CU(cudaEventSynchronize(f->ev[1]));
    CU(cudaEventElapsedTime(ms, f->ev[0], f->ev[1]));
    return QPSK_B200_OK;
}

// =============================================================================================
// batched FFT + argmax
// =============================================================================================
struct qpsk_b200_fft {
    int n, log2n, device, nsm;
    float2* d_tw;
    void* d_stage;  size_t stage_bytes;
    int* d_bin;     float* d_mag2;  size_t out_cap;
    cudaStream_t stream;
    cudaEvent_t ev[2];
    bool timed;
};

extern "C" int qpsk_b200_fft_destroy(qpsk_b200_fft* f) {
    if (!f) return 0;
    cudaSetDevice(f->device);
    if (f->d_tw) cudaFree(f->d_tw);
    if (f->d_stage) cudaFree(f->d_stage);
    if (f->d_bin) cudaFree(f->d_bin);
    if (f->d_mag2) cudaFree(f->d_mag2);
    for (auto& e : f->ev) if (e) cudaEventDestroy(e);
    if (f->stream) cudaStreamDestroy(f->stream);
    delete f;
    return 0;
}

extern "C" int qpsk_b200_fft_create(int n, int device, qpsk_b200_fft** out) {
    if (!out) return fail(QPSK_B200_ERR_ARG, "null argument");
    *out = nullptr;
    int lg = 0;
    while ((1 << lg) < n) lg++;
    if (n < 2 || n > 8192 || (1 << lg) != n) return fail(QPSK_B200_ERR_ARG, "fft length %d unsupported: powers of two 2..8192", n);
    int rc = check_device(device);
    if (rc) return rc;
    qpsk_b200_fft* f = new (std::nothrow) qpsk_b200_fft();
    if (!f) return fail(QPSK_B200_ERR_ARG, "out of host memory");
    memset(f, 0, sizeof *f);
    f->n = n; f->log2n = lg; f->device = device;
    cudaDeviceGetAttribute(&f->nsm, cudaDevAttrMultiProcessorCount, device);
    const int ntw = qpsk_fft_tw_count(n);
    float2* tw = new float2[ntw];
    qpsk_fft_make_twiddles(n, tw);
    cudaError_t e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
    for (auto& ev : f->ev) if (e == cudaSuccess) e = cudaEventCreate(&ev);
    if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_tw, sizeof(float2) * ntw);
    if (e == cudaSuccess) e = cudaMemcpy(f->d_tw, tw, sizeof(float2) * ntw, cudaMemcpyHostToDevice);
    delete[] tw;
    if (e != cudaSuccess) { qpsk_b200_fft_destroy(f); return fail(QPSK_B200_ERR_CUDA, "allocating FFT state failed: %s", cudaGetErrorString(e)); }
    *out = f;
    return QPSK_B200_OK;
}

template <int LOG2N, bool GEN>
static cudaError_t launch_fft_g(const FftArgs& a, int nsm, cudaStream_t s) {
    using Cfg = FftCfg<LOG2N>;
    cudaError_t e = cudaFuncSetAttribute(fft_kernel<LOG2N, GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fft_kernel<LOG2N, GEN>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fft_kernel<LOG2N, GEN>, Cfg::THREADS, Cfg::SMEM);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const int passes = (a.nbursts + Cfg::FPB - 1) / Cfg::FPB;
    int grid = nsm * per_sm;                      // persistent: a multiple of the SM count
    if (grid > passes) grid = passes;
    fft_kernel<LOG2N, GEN><<<grid, Cfg::THREADS, Cfg::SMEM, s>>>(a);
    return cudaGetLastError();
}

// the estimator call (forward, argmax only) has its own lean instantiation
template <int LOG2N>
static cudaError_t launch_fft_n(const FftArgs& a, int nsm, cudaStream_t s) {
    if (a.spectrum == nullptr && a.bin != nullptr && a.im_sign > 0.0f) return launch_fft_g<LOG2N, false>(a, nsm, s);
    return launch_fft_g<LOG2N, true>(a, nsm, s);
}

static cudaError_t launch_fft(int log2n, const FftArgs& a, int nsm, cudaStream_t s) {
    switch (log2n) {
        case 1: return launch_fft_n<1>(a, nsm, s);   case 2: return launch_fft_n<2>(a, nsm, s);
        case 3: return launch_fft_n<3>(a, nsm, s);   case 4: return launch_fft_n<4>(a, nsm, s);
        case 5: return launch_fft_n<5>(a, nsm, s);   case 6: return launch_fft_n<6>(a, nsm, s);
        case 7: return launch_fft_n<7>(a, nsm, s);   case 8: return launch_fft_n<8>(a, nsm, s);
        case 9: return launch_fft_n<9>(a, nsm, s);   case 10: return launch_fft_n<10>(a, nsm, s);
        case 11: return launch_fft_n<11>(a, nsm, s); case 12: return launch_fft_n<12>(a, nsm, s);
        case 13: return launch_fft_n<13>(a, nsm, s);
        default: return cudaErrorInvalidValue;
    }
}

static int fft_run(qpsk_b200_fft* f, const float* d_in, float* d_out, int nbursts, int inverse, int32_t* d_bin, float* d_mag2, cudaStream_t s) {
    // n >= 2048: the bursts travel by bulk copy (cp.async.bulk), which wants 16-byte aligned sources
    if (f->n >= 2048 && (reinterpret_cast<uintptr_t>(d_in) & 15) != 0) return fail(QPSK_B200_ERR_ARG, "fft input must be 16-byte aligned for n >= 2048");
    FftArgs a;
    a.in = reinterpret_cast<const float2*>(d_in); a.spectrum = reinterpret_cast<float2*>(d_out);
    a.bin = d_bin; a.mag2 = d_mag2; a.tw = f->d_tw; a.nbursts = nbursts;
    a.im_sign = inverse ? -1.0f : 1.0f;
    a.scale = inverse ? 1.0f : 1.0f / (float)f->n;           // fft.c:105-107 vs fft.c:130-136
    fft_consts_host(a.kbase);
    fft_wsplit_host(a.wsplit);
    fft_two_host(a.two);
    CU(cudaEventRecord(f->ev[0], s));
    cudaError_t e = launch_fft(f->log2n, a, f->nsm, s);
    if (e != cudaSuccess) return fail(QPSK_B200_ERR_CUDA, "FFT kernel launch failed: %s", cudaGetErrorString(e));
    CU(cudaEventRecord(f->ev[1], s));
    f->timed = true;
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_fft_argmax_device(qpsk_b200_fft* f, const float* d_in, int nbursts, int32_t* d_bin, float* d_mag2, void* cuda_stream) {
    if (!f || !d_in || !d_bin) return fail(QPSK_B200_ERR_ARG, "null argument");
    if (nbursts < 1) return fail(QPSK_B200_ERR_ARG, "nbursts must be positive");
    CU(cudaSetDevice(f->device));
    return fft_run(f, d_in, nullptr, nbursts, 0, d_bin, d_mag2, cuda_stream ? (cudaStream_t)cuda_stream : f->stream);
}

extern "C" int qpsk_b200_fft_transform_device(qpsk_b200_fft* f, const float* d_in, float* d_out, int nbursts, int inverse, void* cuda_stream) {
    if (!f || !d_in || !d_out) return fail(QPSK_B200_ERR_ARG, "null argument");
    if (nbursts < 1) return fail(QPSK_B200_ERR_ARG, "nbursts must be positive");
    CU(cudaSetDevice(f->device));
    return fft_run(f, d_in, d_out, nbursts, inverse, nullptr, nullptr, cuda_stream ? (cudaStream_t)cuda_stream : f->stream);
}

static int fft_stage_in(qpsk_b200_fft* f, const float* h_in, int nbursts) {
    const size_t bytes = (size_t)nbursts * f->n * sizeof(float2);
    if (f->stage_bytes < bytes) {
        if (f->d_stage) { cudaFree(f->d_stage); f->d_stage = nullptr; f->stage_bytes = 0; }
        CU(cudaMalloc(&f->d_stage, bytes));
        f->stage_bytes = bytes;
    }
    CU(cudaMemcpyAsync(f->d_stage, h_in, bytes, cudaMemcpyHostToDevice, f->stream));
    return 0;
}

extern "C" int qpsk_b200_fft_argmax_host(qpsk_b200_fft* f, const float* h_in, int nbursts, int32_t* h_bin, float* h_mag2) {
    if (!f || !h_in || !h_bin) return fail(QPSK_B200_ERR_ARG, "null argument");
    if (nbursts < 1) return fail(QPSK_B200_ERR_ARG, "nbursts must be positive");
    CU(cudaSetDevice(f->device));
    int rc = fft_stage_in(f, h_in, nbursts);
    if (rc) return rc;
    if (f->out_cap < (size_t)nbursts) {
        if (f->d_bin) cudaFree(f->d_bin);
        if (f->d_mag2) cudaFree(f->d_mag2);
        f->d_bin = nullptr; f->d_mag2 = nullptr; f->out_cap = 0;
        CU(cudaMalloc((void**)&f->d_bin, sizeof(int) * nbursts));
        CU(cudaMalloc((void**)&f->d_mag2, sizeof(float) * nbursts));
        f->out_cap = nbursts;
    }
    rc = fft_run(f, reinterpret_cast<const float*>(f->d_stage), nullptr, nbursts, 0, f->d_bin, f->d_mag2, f->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h_bin, f->d_bin, sizeof(int) * nbursts, cudaMemcpyDeviceToHost, f->stream));
    if (h_mag2) CU(cudaMemcpyAsync(h_mag2, f->d_mag2, sizeof(float) * nbursts, cudaMemcpyDeviceToHost, f->stream));
    CU(cudaStreamSynchronize(f->stream));
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_fft_transform_host(qpsk_b200_fft* f, const float* h_in, float* h_out, int nbursts, int inverse) {
    if (!f || !h_in || !h_out) return fail(QPSK_B200_ERR_ARG, "null argument");
    if (nbursts < 1) return fail(QPSK_B200_ERR_ARG, "nbursts must be positive");
    CU(cudaSetDevice(f->device));
    int rc = fft_stage_in(f, h_in, nbursts);
    if (rc) return rc;
    rc = fft_run(f, reinterpret_cast<const float*>(f->d_stage), reinterpret_cast<float*>(f->d_stage), nbursts, inverse, nullptr, nullptr, f->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h_out, f->d_stage, (size_t)nbursts * f->n * sizeof(float2), cudaMemcpyDeviceToHost, f->stream));
    CU(cudaStreamSynchronize(f->stream));
    return QPSK_B200_OK;
}

// ---- transforms longer than one CTA can hold (n > 8192): the four-step decomposition n = N1 N2 over the batched kernel --
// N2 transforms of length N1 down the columns (after a transpose), the twiddles w_n^(i2 k1) (evaluated in double) applied
// while transposing back, N1 transforms of length N2, a last transpose into natural order.  Forward is scaled by
// 1/N1 * 1/N2 = 1/n and the inverse is unscaled, as in the reference (fft.c:105-107, 122-128), which takes any power of two.
__global__ void fft_transpose_twiddle_kernel(const float2* __restrict__ src, float2* __restrict__ dst, int rows, int cols, long long n, int tw_sign) {
    __shared__ float2 tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        if (r < rows && c < cols) {
            float2 v = src[(size_t)r * cols + c];
            if (tw_sign != 0) {                                   // times exp(-+ 2 pi i r c / n)
                const long long e = ((long long)r * c) % n;
                double sn, cs;
                sincospi(2.0 * (double)e / (double)n, &sn, &cs);
                const double wr = cs, wi = tw_sign > 0 ? -sn : sn;
                v = make_float2((float)((double)v.x * wr - (double)v.y * wi), (float)((double)v.x * wi + (double)v.y * wr));
            }
            tile[j][threadIdx.x] = v;
        }
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[(size_t)c * rows + r] = tile[threadIdx.x][j];
    }
}

extern "C" int qpsk_b200_fft_big_host(const float* h_in, float* h_out, int n, int inverse, int device) {
    if (!h_in || !h_out) return fail(QPSK_B200_ERR_ARG, "null argument");
    int lg = 0;
    while ((1LL << lg) < n) lg++;
    if (n <= 8192 || (1LL << lg) != n || lg > 26) return fail(QPSK_B200_ERR_ARG, "long fft length %d unsupported: powers of two 16384..2^26", n);
    int rc = check_device(device);
    if (rc) return rc;
    const int n1 = 1 << ((lg + 1) / 2), n2 = n / n1;
    qpsk_b200_fft *f1 = nullptr, *f2 = nullptr;
    DevBuf a, b;
    cudaStream_t s = nullptr;
    auto done = [&](int code) { if (f1) qpsk_b200_fft_destroy(f1); if (f2) qpsk_b200_fft_destroy(f2); if (s) cudaStreamDestroy(s); return code; };
    if ((rc = qpsk_b200_fft_create(n1, device, &f1)) != 0) return done(rc);
    if ((rc = qpsk_b200_fft_create(n2, device, &f2)) != 0) return done(rc);
    if (cudaMalloc(&a.p, (size_t)n * sizeof(float2)) != cudaSuccess || cudaMalloc(&b.p, (size_t)n * sizeof(float2)) != cudaSuccess ||
        cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess)
        return done(fail(QPSK_B200_ERR_CUDA, "allocating the long transform failed: %s", cudaGetErrorString(cudaGetLastError())));
    float2 *A = (float2*)a.p, *B = (float2*)b.p;
    const dim3 blk(32, 8);
    cudaError_t e = cudaMemcpyAsync(A, h_in, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, s);
    // x[i1 N2 + i2] as [N1][N2] -> [N2][N1]
    if (e == cudaSuccess) { fft_transpose_twiddle_kernel<<<dim3((n2 + 31) / 32, (n1 + 31) / 32), blk, 0, s>>>(A, B, n1, n2, n, 0); e = cudaGetLastError(); }
    if (e != cudaSuccess) return done(fail(QPSK_B200_ERR_CUDA, "long transform failed: %s", cudaGetErrorString(e)));
    if ((rc = qpsk_b200_fft_transform_device(f1, (const float*)B, (float*)B, n2, inverse, s)) != 0) return done(rc);
    // Y[i2][k1] w^(i2 k1) -> [N1][N2]
    fft_transpose_twiddle_kernel<<<dim3((n1 + 31) / 32, (n2 + 31) / 32), blk, 0, s>>>(B, A, n2, n1, n, inverse ? -1 : 1);
    if ((e = cudaGetLastError()) != cudaSuccess) return done(fail(QPSK_B200_ERR_CUDA, "long transform failed: %s", cudaGetErrorString(e)));
    if ((rc = qpsk_b200_fft_transform_device(f2, (const float*)A, (float*)A, n1, inverse, s)) != 0) return done(rc);
    // Z[k1][k2] -> X[k1 + N1 k2] = [N2][N1]
    fft_transpose_twiddle_kernel<<<dim3((n2 + 31) / 32, (n1 + 31) / 32), blk, 0, s>>>(A, B, n1, n2, n, 0);
    if ((e = cudaGetLastError()) == cudaSuccess) e = cudaMemcpyAsync(h_out, B, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return done(fail(QPSK_B200_ERR_CUDA, "long transform failed: %s", cudaGetErrorString(e)));
    return done(QPSK_B200_OK);
}

extern "C" int qpsk_b200_fft_last_kernel_ms(qpsk_b200_fft* f, float* ms) {
    if (!f || !ms || !f->timed) return fail(QPSK_B200_ERR_STATE, "no transform yet");
    CU(cudaEventSynchronize(f->ev[1]));
    CU(cudaEventElapsedTime(ms, f->ev[0], f->ev[1]));
    return QPSK_B200_OK;
}

// =============================================================================================
// bit stages: standalone entry points on host buffers
// =============================================================================================
extern "C" int qpsk_b200_rx_crc_counters(qpsk_b200_rx* rx, unsigned long long* frames, unsigned long long* passes) {
    if (!rx) return fail(QPSK_B200_ERR_ARG, "null receiver");
    if (!rx->d_counters) return fail(QPSK_B200_ERR_STATE, "frames are not decoded (QPSK_B200_DECODE_FRAMES)");
    CU(cudaSetDevice(rx->cfg.device));
    CU(cudaDeviceSynchronize());
    unsigned long long h[2];
    CU(cudaMemcpy(h, rx->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    if (frames) *frames = h[0];
    if (passes) *passes = h[1];
    return QPSK_B200_OK;
}


extern "C" int qpsk_b200_bits_crc16(const uint8_t* h_data, int nbytes, int nframes, uint16_t* h_crc, int device) {
    if (!h_crc || nframes < 1 || nbytes < 0 || (nbytes > 0 && !h_data)) return fail(QPSK_B200_ERR_ARG, "bad argument");
    int rc = check_device(device);
    if (rc) return rc;
    DevBuf d, c;
    const size_t bytes = (size_t)nbytes * nframes;
    CU(cudaMalloc(&d.p, bytes ? bytes : 1));
    CU(cudaMalloc(&c.p, sizeof(uint16_t) * nframes));
    if (bytes) CU(cudaMemcpy(d.p, h_data, bytes, cudaMemcpyHostToDevice));
    crc16_rows_kernel<<<(nframes + 127) / 128, 128>>>((const uint8_t*)d.p, nbytes, nframes, (uint16_t*)c.p);
    CU(cudaGetLastError());
    CU(cudaMemcpy(h_crc, c.p, sizeof(uint16_t) * nframes, cudaMemcpyDeviceToHost));
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_bits_interleave(uint8_t* h_data, int nbytes, int nframes, int dir, int device) {
    if (!h_data || nframes < 1 || nbytes < 1) return fail(QPSK_B200_ERR_ARG, "bad argument");
    if (nbytes >= 8192) return fail(QPSK_B200_ERR_ARG, "nbytes %d: the reference's uint16_t bit count wraps at 8192 bytes (interleave.c:49)", nbytes);
    if (dir != 0 && dir != 1) return fail(QPSK_B200_ERR_ARG, "dir must be 0 (INTERLEAVE) or 1 (DEINTERLEAVE)");
    int rc = check_device(device);
    if (rc) return rc;
    DevBuf d;
    const size_t bytes = (size_t)nbytes * nframes;
    CU(cudaMalloc(&d.p, bytes));
    CU(cudaMemcpy(d.p, h_data, bytes, cudaMemcpyHostToDevice));
    const int b = interleave_prime(nbytes * 8);
    const int warps = 8;
    const size_t smem = (size_t)warps * ((nbytes + 3) / 4) * sizeof(unsigned);
    CU(cudaFuncSetAttribute(interleave_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    interleave_rows_kernel<<<(nframes + warps - 1) / warps, warps * 32, smem>>>((uint8_t*)d.p, nbytes, nframes, b, dir);
    CU(cudaGetLastError());
    CU(cudaMemcpy(h_data, d.p, bytes, cudaMemcpyDeviceToHost));
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_bits_scramble(uint8_t* h_dibits, int ndibits, int nframes, int device) {
    if (!h_dibits || nframes < 1 || ndibits < 1) return fail(QPSK_B200_ERR_ARG, "bad argument");
    int rc = check_device(device);
    if (rc) return rc;
    // the keystream does not depend on the data: evaluate the LFSR once per call, XOR on the device
    uint8_t* ks = new (std::nothrow) uint8_t[ndibits];
    if (!ks) return fail(QPSK_B200_ERR_ARG, "out of host memory");
    uint16_t reg = (uint16_t)QPSK_SCRAMBLE_SEED;
    for (int k = 0; k < ndibits; k++) { const unsigned b0 = lfsr_step(reg); const unsigned b1 = lfsr_step(reg); ks[k] = (uint8_t)(b0 | (b1 << 1)); }
    DevBuf d, k;
    const size_t total = (size_t)ndibits * nframes;
    cudaError_t e = cudaMalloc(&d.p, total);
    if (e == cudaSuccess) e = cudaMalloc(&k.p, ndibits);
    if (e == cudaSuccess) e = cudaMemcpy(d.p, h_dibits, total, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(k.p, ks, ndibits, cudaMemcpyHostToDevice);
    delete[] ks;
    if (e != cudaSuccess) return fail(QPSK_B200_ERR_CUDA, "scramble staging failed: %s", cudaGetErrorString(e));
    scramble_rows_kernel<<<(unsigned)((total + 255) / 256), 256>>>((uint8_t*)d.p, (const uint8_t*)k.p, ndibits, total);
    CU(cudaGetLastError());
    CU(cudaMemcpy(h_dibits, d.p, total, cudaMemcpyDeviceToHost));
    return QPSK_B200_OK;
}

// [C][rows] channel-major words -> [rows][Cpad] channel-fastest (the upload twin of transpose_to_channel_major)
__global__ void transpose_from_channel_major(const unsigned* __restrict__ src, unsigned* __restrict__ dst, int rows, int C, int Cpad) {
    __shared__ unsigned tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < C) tile[j][threadIdx.x] = src[(size_t)c * rows + r];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        if (r < rows && c < C) dst[(size_t)r * Cpad + c] = tile[threadIdx.x][j];
    }
}

static int frames_codec(const uint8_t* h_in, int nbytes, int nchan, int nframes, uint8_t* h_out, uint8_t* h_crc_ok, uint8_t* h_rotation,
                        int device, bool encode) {
    if (!h_in || !h_out || nchan < 1 || nframes < 1) return fail(QPSK_B200_ERR_ARG, "bad argument");
    if (nbytes != 16 && nbytes != 32) return fail(QPSK_B200_ERR_ARG, "frame codec supports 16- and 32-byte frames, got %d", nbytes);
    int rc = check_device(device);
    if (rc) return rc;
    const int W = nbytes / 4, rows = nframes * W, Cpad = (nchan + 31) / 32 * 32;
    const size_t words = (size_t)rows * Cpad, bytes_cm = (size_t)nchan * rows * 4;
    DevBuf cm, a, b, ok, rot, cnt, okcm;
    CU(cudaMalloc(&cm.p, bytes_cm));
    CU(cudaMalloc(&a.p, words * 4));
    CU(cudaMalloc(&b.p, words * 4));
    CU(cudaMemcpy(cm.p, h_in, bytes_cm, cudaMemcpyHostToDevice));
    dim3 tgrid((nchan + 31) / 32, (rows + 31) / 32), tblock(32, 8);
    transpose_from_channel_major<<<tgrid, tblock>>>((const unsigned*)cm.p, (unsigned*)a.p, rows, nchan, Cpad);
    CU(cudaGetLastError());
    if (encode) {
        FrameEncodeArgs ea;
        ea.ks = make_keystream(nbytes);
        ea.payload_t = (const unsigned*)a.p; ea.dibits_t = (unsigned*)b.p; ea.C = nchan; ea.Cpad = Cpad; ea.F = nframes;
        dim3 grid((nchan + 127) / 128, nframes);
        if (nbytes == 32) frame_encode_kernel<32><<<grid, 128>>>(ea); else frame_encode_kernel<16><<<grid, 128>>>(ea);
        CU(cudaGetLastError());
    } else {
        CU(cudaMalloc(&ok.p, (size_t)nframes * Cpad));
        CU(cudaMalloc(&cnt.p, 2 * sizeof(unsigned long long)));
        CU(cudaMemset(cnt.p, 0, 2 * sizeof(unsigned long long)));
        if (h_rotation) CU(cudaMalloc(&rot.p, (size_t)nframes * Cpad));
        rc = launch_frame_decode(nbytes, (const unsigned*)a.p, (unsigned*)b.p, (uint8_t*)ok.p, (uint8_t*)rot.p, (unsigned long long*)cnt.p, 0, nchan,
                                 Cpad, nframes, 0);
        if (rc) return rc;
    }
    transpose_to_channel_major<unsigned><<<tgrid, tblock>>>((const unsigned*)b.p, (unsigned*)cm.p, rows, nchan, Cpad);
    CU(cudaGetLastError());
    CU(cudaMemcpy(h_out, cm.p, bytes_cm, cudaMemcpyDeviceToHost));
    if (!encode && h_crc_ok) {
        CU(cudaMalloc(&okcm.p, (size_t)nchan * nframes));
        dim3 g2((nchan + 31) / 32, (nframes + 31) / 32);
        transpose_to_channel_major<uint8_t><<<g2, tblock>>>((const uint8_t*)ok.p, (uint8_t*)okcm.p, nframes, nchan, Cpad);
        CU(cudaGetLastError());
        CU(cudaMemcpy(h_crc_ok, okcm.p, (size_t)nchan * nframes, cudaMemcpyDeviceToHost));
    }
    if (!encode && h_rotation) {
        if (!okcm.p) CU(cudaMalloc(&okcm.p, (size_t)nchan * nframes));
        dim3 g2((nchan + 31) / 32, (nframes + 31) / 32);
        transpose_to_channel_major<uint8_t><<<g2, tblock>>>((const uint8_t*)rot.p, (uint8_t*)okcm.p, nframes, nchan, Cpad);
        CU(cudaGetLastError());
        CU(cudaMemcpy(h_rotation, okcm.p, (size_t)nchan * nframes, cudaMemcpyDeviceToHost));
    }
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_frames_encode(const uint8_t* h_payload, int nbytes, int nchan, int nframes, uint8_t* h_dibits, int device) {
    return frames_codec(h_payload, nbytes, nchan, nframes, h_dibits, nullptr, nullptr, device, true);
}

extern "C" int qpsk_b200_frames_decode(const uint8_t* h_dibits, int nbytes, int nchan, int nframes, uint8_t* h_frames, uint8_t* h_crc_ok, int device) {
    return frames_codec(h_dibits, nbytes, nchan, nframes, h_frames, h_crc_ok, nullptr, device, false);
}

extern "C" int qpsk_b200_frames_decode_rotated(const uint8_t* h_dibits, int nbytes, int nchan, int nframes, uint8_t* h_frames,
                                               uint8_t* h_crc_ok, uint8_t* h_rotation, int device) {
    if (!h_rotation) return fail(QPSK_B200_ERR_ARG, "bad argument");
    return frames_codec(h_dibits, nbytes, nchan, nframes, h_frames, h_crc_ok, h_rotation, device, false);
}

// =============================================================================================
// batched transmit path
// =============================================================================================
struct qpsk_b200_tx {
    long long id;
    int C, Cpad, sps, ntaps, packet_samples, sample_pos, device;
    float fs;
    float taps[QPSK_MAX_TAPS];
    float2* d_phase;     // [Cpad]
    float2* d_rect;      // [Cpad]
    float2* d_hist;      // [Cpad][128/sps]
    uint8_t* d_sym_stage;  int16_t* d_pcm_stage;  size_t stage_syms;
    cudaStream_t stream;
    cudaStream_t last_stream;   // the stream the most recent modulate call ran on (end_packet / reset order themselves after it)
};

extern "C" int qpsk_b200_tx_destroy(qpsk_b200_tx* tx) {
    if (!tx) return 0;
    cudaSetDevice(tx->device);
    void* ptrs[] = { tx->d_phase, tx->d_rect, tx->d_hist, tx->d_sym_stage, tx->d_pcm_stage };
    for (void* p : ptrs) if (p) cudaFree(p);
    if (tx->stream) cudaStreamDestroy(tx->stream);
    delete tx;
    return 0;
}

extern "C" int qpsk_b200_tx_reset(qpsk_b200_tx* tx) {
    if (!tx) return fail(QPSK_B200_ERR_ARG, "null transmitter");
    CU(cudaSetDevice(tx->device));
    CU(cudaDeviceSynchronize());          // the legacy-stream copies below do not order themselves after non-blocking streams
    float c0[2];
    qpsk_host_cis(0.0, 0, c0);                                                  // qpsk.c:316 fbb_tx_phase = cmplx(0)
    float2* ph = new float2[tx->Cpad];
    for (int i = 0; i < tx->Cpad; i++) ph[i] = make_float2(c0[0], c0[1]);
    cudaError_t e = cudaMemcpy(tx->d_phase, ph, sizeof(float2) * tx->Cpad, cudaMemcpyHostToDevice);
    delete[] ph;
    if (e != cudaSuccess) return fail(QPSK_B200_ERR_CUDA, "tx reset failed: %s", cudaGetErrorString(e));
    CU(cudaMemset(tx->d_hist, 0, sizeof(float2) * tx->Cpad * (QPSK_CHUNK / tx->sps)));
    tx->sample_pos = 0;
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_tx_create(float fs, float rs, float rrc_alpha, const float* carrier_hz, int nchan, int packet_symbols,
                                   int device, qpsk_b200_tx** out) {
    if (!carrier_hz || !out) return fail(QPSK_B200_ERR_ARG, "null argument");
    *out = nullptr;
    const int sps = (int)((double)fs / (double)rs);
    if (sps != 4 && sps != 8) return fail(QPSK_B200_ERR_ARG, "samples/symbol %d unsupported (4 or 8)", sps);
    if (nchan < 1 || packet_symbols < 1 || (packet_symbols * sps) % QPSK_CHUNK != 0)
        return fail(QPSK_B200_ERR_ARG, "packet_symbols*sps must be a positive multiple of %d", QPSK_CHUNK);
    int rc = check_device(device);
    if (rc) return rc;
    qpsk_b200_tx* tx = new (std::nothrow) qpsk_b200_tx();
    if (!tx) return fail(QPSK_B200_ERR_ARG, "out of host memory");
    memset(tx, 0, sizeof *tx);
    tx->id = g_next_id++;
    tx->C = nchan; tx->Cpad = (nchan + 31) / 32 * 32; tx->sps = sps; tx->ntaps = 127; tx->device = device;
    tx->packet_samples = packet_symbols * sps;
    tx->fs = fs;
    qpsk_host_rrc_make(tx->taps, tx->ntaps, fs, rs, rrc_alpha);                 // qpsk.c:308
    float2* rect = new float2[tx->Cpad];
    for (int i = 0; i < tx->Cpad; i++) {
        float t2[2];
        qpsk_host_cis(kTau * (double)carrier_hz[i < nchan ? i : nchan - 1] / (double)fs, 0, t2);   // qpsk.c:320
        rect[i] = make_float2(t2[0], t2[1]);
    }
    cudaError_t e = cudaStreamCreateWithFlags(&tx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void**)&tx->d_phase, sizeof(float2) * tx->Cpad);
    if (e == cudaSuccess) e = cudaMalloc((void**)&tx->d_rect, sizeof(float2) * tx->Cpad);
    if (e == cudaSuccess) e = cudaMalloc((void**)&tx->d_hist, sizeof(float2) * tx->Cpad * (QPSK_CHUNK / sps));
    if (e == cudaSuccess) e = cudaMemcpy(tx->d_rect, rect, sizeof(float2) * tx->Cpad, cudaMemcpyHostToDevice);
    delete[] rect;
    if (e != cudaSuccess) { qpsk_b200_tx_destroy(tx); return fail(QPSK_B200_ERR_CUDA, "allocating transmitter state failed: %s", cudaGetErrorString(e)); }
    rc = qpsk_b200_tx_reset(tx);
    if (rc) { qpsk_b200_tx_destroy(tx); return rc; }
    *out = tx;
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_tx_set_carrier(qpsk_b200_tx* tx, const float* carrier_hz) {
    if (!tx || !carrier_hz) return fail(QPSK_B200_ERR_ARG, "null argument");
    CU(cudaSetDevice(tx->device));
    float2* rect = new (std::nothrow) float2[tx->Cpad];
    if (!rect) return fail(QPSK_B200_ERR_ARG, "out of host memory");
    for (int i = 0; i < tx->Cpad; i++) {
        float t2[2];
        qpsk_host_cis(kTau * (double)carrier_hz[i < tx->C ? i : tx->C - 1] / (double)tx->fs, 0, t2);   // qpsk.c:320
        rect[i] = make_float2(t2[0], t2[1]);
    }
    cudaError_t e = cudaDeviceSynchronize();                      // earlier calls still read the old table
    if (e == cudaSuccess) e = cudaMemcpy(tx->d_rect, rect, sizeof(float2) * tx->Cpad, cudaMemcpyHostToDevice);
    delete[] rect;
    if (e != cudaSuccess) return fail(QPSK_B200_ERR_CUDA, "carrier upload failed: %s", cudaGetErrorString(e));
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_channel_awgn_device(int16_t* d_pcm, int nchan, long long nsamples, const float* h_sigma, unsigned long long seed,
                                             long long first_sample, int first_channel, int device, void* cuda_stream) {
    if (!d_pcm || !h_sigma) return fail(QPSK_B200_ERR_ARG, "null argument");
    if (nchan < 1 || nsamples < 1 || first_sample < 0 || (first_sample & 1)) return fail(QPSK_B200_ERR_ARG, "bad argument");
    if ((unsigned long long)nchan * (unsigned long long)((nsamples + 511) / 512) > 0x7fffffffull) return fail(QPSK_B200_ERR_ARG, "too many samples for one call");
    int rc = check_device(device);
    if (rc) return rc;
    CU(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)cuda_stream;
    DevBuf sg;
    CU(cudaMalloc(&sg.p, sizeof(float) * nchan));
    CU(cudaMemcpyAsync(sg.p, h_sigma, sizeof(float) * nchan, cudaMemcpyHostToDevice, s));
    AwgnArgs a;
    a.pcm = d_pcm; a.sigma = (const float*)sg.p; a.C = nchan; a.T = nsamples; a.first_sample = first_sample;
    a.seed_lo = (unsigned)seed; a.seed_hi = (unsigned)(seed >> 32); a.chan_base = first_channel;
    const long long pairs = (nsamples + 1) / 2;
    awgn_kernel<<<(unsigned)((pairs + 255) / 256) * (unsigned)nchan, 256, 0, s>>>(a);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s));                                 // the sigma table is freed on return
    return QPSK_B200_OK;
}

template <int SPS>
static cudaError_t launch_tx(const TxArgs& a, const float* taps, cudaStream_t s) {
    const size_t smem = sizeof(TxSmem<SPS>);
    cudaError_t e = cudaFuncSetAttribute(tx_kernel<127, SPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tx_kernel<127, SPS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    tx_kernel<127, SPS><<<a.Cpad / QPSK_GROUP, 256, smem, s>>>(a, tap_bank<127>(taps));
    return cudaGetLastError();
}

static int tx_run(qpsk_b200_tx* tx, const uint8_t* d_symbols, const float2* d_symbols_cf, int nsym, int16_t* d_pcm, void* cuda_stream) {
    if (!tx || (!d_symbols && !d_symbols_cf) || !d_pcm) return fail(QPSK_B200_ERR_ARG, "null argument");
    if (nsym < 1) return fail(QPSK_B200_ERR_ARG, "nsym %d must be positive", nsym);
    if ((reinterpret_cast<uintptr_t>(d_pcm) & 7) != 0) return fail(QPSK_B200_ERR_ARG, "d_pcm must be 8-byte aligned");
    CU(cudaSetDevice(tx->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : tx->stream;
    TxArgs a;
    a.symbols = d_symbols; a.symbols_cf = d_symbols_cf; a.pcm = d_pcm; a.phase_state = tx->d_phase; a.rect = tx->d_rect; a.sym_hist = tx->d_hist;
    a.C = tx->C; a.Cpad = tx->Cpad; a.nsym = nsym; a.packet_samples = tx->packet_samples; a.sample_pos = tx->sample_pos;
    cudaError_t e = tx->sps == 4 ? launch_tx<4>(a, tx->taps, s) : launch_tx<8>(a, tx->taps, s);
    if (e != cudaSuccess) return fail(QPSK_B200_ERR_CUDA, "tx kernel launch failed: %s", cudaGetErrorString(e));
    tx->sample_pos = (int)(((long long)tx->sample_pos + (long long)nsym * tx->sps) % tx->packet_samples);
    tx->last_stream = s;
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_tx_process_device(qpsk_b200_tx* tx, const uint8_t* d_symbols, int nsym, int16_t* d_pcm, void* cuda_stream) {
    return tx_run(tx, d_symbols, nullptr, nsym, d_pcm, cuda_stream);
}

extern "C" int qpsk_b200_tx_symbols_host(qpsk_b200_tx* tx, const float* h_symbols, int nsym, int16_t* h_pcm) {
    if (!tx || !h_symbols || !h_pcm) return fail(QPSK_B200_ERR_ARG, "null argument");
    if (nsym < 1) return fail(QPSK_B200_ERR_ARG, "nsym must be positive");
    CU(cudaSetDevice(tx->device));
    const size_t syms = (size_t)tx->C * nsym;
    DevBuf ds, dp;
    CU(cudaMalloc(&ds.p, syms * sizeof(float2)));
    CU(cudaMalloc(&dp.p, syms * tx->sps * sizeof(int16_t)));
    CU(cudaMemcpyAsync(ds.p, h_symbols, syms * sizeof(float2), cudaMemcpyHostToDevice, tx->stream));
    int rc = tx_run(tx, nullptr, (const float2*)ds.p, nsym, (int16_t*)dp.p, tx->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h_pcm, dp.p, syms * tx->sps * sizeof(int16_t), cudaMemcpyDeviceToHost, tx->stream));
    CU(cudaStreamSynchronize(tx->stream));
    return QPSK_B200_OK;
}

extern "C" int qpsk_b200_tx_process_host(qpsk_b200_tx* tx, const uint8_t* h_symbols, int nsym, int16_t* h_pcm) {
    if (!tx || !h_symbols || !h_pcm) return fail(QPSK_B200_ERR_ARG, "null argument");
    if (nsym < 1) return fail(QPSK_B200_ERR_ARG, "nsym must be positive");
    CU(cudaSetDevice(tx->device));
    const size_t syms = (size_t)tx->C * nsym;
    if (tx->stage_syms < syms) {
        if (tx->d_sym_stage) cudaFree(tx->d_sym_stage);
        if (tx->d_pcm_stage) cudaFree(tx->d_pcm_stage);
        tx->d_sym_stage = nullptr; tx->d_pcm_stage = nullptr; tx->stage_syms = 0;
        CU(cudaMalloc((void**)&tx->d_sym_stage, syms));
        CU(cudaMalloc((void**)&tx->d_pcm_stage, syms * tx->sps * sizeof(int16_t)));
        tx->stage_syms = syms;
    }
    CU(cudaMemcpyAsync(tx->d_sym_stage, h_symbols, syms, cudaMemcpyHostToDevice, tx->stream));
    int rc = qpsk_b200_tx_process_device(tx, tx->d_sym_stage, nsym, tx->d_pcm_stage, tx->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h_pcm, tx->d_pcm_stage, syms * tx->sps * sizeof(int16_t), cudaMemcpyDeviceToHost, tx->stream));
    CU(cudaStreamSynchronize(tx->stream));
    return QPSK_B200_OK;
}

__global__ void tx_normalise_kernel(float2* __restrict__ phase, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float2 ph = phase[c];
    const double dr = (double)ph.x, di = (double)ph.y;                 // qpsk.c:253 with glibc hypotf semantics
    const float mag = __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(di, di))));
    phase[c] = make_float2(__fdiv_rn(ph.x, mag), __fdiv_rn(ph.y, mag));
}

extern "C" int qpsk_b200_tx_end_packet(qpsk_b200_tx* tx) {
    if (!tx) return fail(QPSK_B200_ERR_ARG, "null transmitter");
    CU(cudaSetDevice(tx->device));
    if (tx->sample_pos != 0) {     // not already normalised by a packet boundary inside the last call
        // on the stream the last modulate call used: the kernel there stores the phasor this one normalises
        cudaStream_t s = tx->last_stream ? tx->last_stream : tx->stream;
        tx_normalise_kernel<<<(tx->Cpad + 127) / 128, 128, 0, s>>>(tx->d_phase, tx->Cpad);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(s));
        tx->sample_pos = 0;
    }
    return QPSK_B200_OK;
}

// =============================================================================================
// test hook: the device NCO (sincosf_glibc) over arbitrary arguments
// =============================================================================================
__global__ void nco_probe_kernel(const float* __restrict__ in, float* __restrict__ s_out, float* __restrict__ c_out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s, c;
    sincosf_glibc(in[i], s, c);
    s_out[i] = s;
    c_out[i] = c;
}

extern "C" int qpsk_b200_debug_nco(const float* h_in, float* h_sin, float* h_cos, int n, int device) {
    if (!h_in || !h_sin || !h_cos || n < 1) return fail(QPSK_B200_ERR_ARG, "bad argument");
    int rc = check_device(device);
    if (rc) return rc;
    DevBuf a, b, c;
    CU(cudaMalloc(&a.p, sizeof(float) * n));
    CU(cudaMalloc(&b.p, sizeof(float) * n));
    CU(cudaMalloc(&c.p, sizeof(float) * n));
    CU(cudaMemcpy(a.p, h_in, sizeof(float) * n, cudaMemcpyHostToDevice));
    nco_probe_kernel<<<(n + 255) / 256, 256>>>((const float*)a.p, (float*)b.p, (float*)c.p, n);
    CU(cudaGetLastError());
    CU(cudaMemcpy(h_sin, b.p, sizeof(float) * n, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_cos, c.p, sizeof(float) * n, cudaMemcpyDeviceToHost));
    return QPSK_B200_OK;
}

// =============================================================================================
// Extension (never part of the parity path): carrier-offset estimate of every channel
// from the 4th power of its decimated symbols, through the batched FFT + argmax kernel.  QPSK on the axes
// raised to the 4th power is a pure tone at 4 x offset.
// =============================================================================================
__global__ void symbol_power4_kernel(const float2* __restrict__ ring, float2* __restrict__ bursts, int c_begin, int C /* end */, int Cpad, int nsym,
                                     int nslots, int first_slot, int n /* burst length */) {
    // bursts[c][k] = ring[frame k / nsym][k % nsym][c] ^ 4 ; 32 x 32 tile transpose (channel-fastest -> channel-major)
    __shared__ float2 tile[32][33];
    const int c0 = c_begin + blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int k = k0 + j, c = c0 + threadIdx.x;
        if (k < n && c < C) {
            const int slot = (first_slot + k / nsym) % nslots;
            tile[j][threadIdx.x] = power4_exact(ring[((size_t)slot * nsym + (k % nsym)) * Cpad + c]);
        }
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, k = k0 + threadIdx.x;
        if (k < n && c < C) bursts[(size_t)c * n + k] = tile[threadIdx.x][j];
    }
}

extern "C" int qpsk_b200_rx_estimate_offset(qpsk_b200_rx* rx, int log2n, float* h_offset_hz, int32_t* h_bin) {
    if (!rx || !h_offset_hz) return fail(QPSK_B200_ERR_ARG, "null argument");
    if (rx->lastF == 0) return fail(QPSK_B200_ERR_STATE, "no process call yet");
    const int n = 1 << log2n;
    if (log2n < 5 || log2n > 13 || n > rx->lastF * rx->nsym)
        return fail(QPSK_B200_ERR_ARG, "burst length 2^%d must be 32..8192 and at most the %d symbols of the last call", log2n, rx->lastF * rx->nsym);
    CU(cudaSetDevice(rx->cfg.device));
    CU(cudaDeviceSynchronize());
    if (!rx->est_fft || rx->est_fft_n != n) {
        if (rx->est_fft) { qpsk_b200_fft_destroy(rx->est_fft); rx->est_fft = nullptr; }
        int rc = qpsk_b200_fft_create(n, rx->cfg.device, &rx->est_fft);
        if (rc) return rc;
        rx->est_fft_n = n;
    }
    DevBuf bursts, bins, mags;
    CU(cudaMalloc(&bursts.p, (size_t)rx->C * n * sizeof(float2)));
    CU(cudaMalloc(&bins.p, (size_t)rx->C * sizeof(int)));
    CU(cudaMalloc(&mags.p, (size_t)rx->C * sizeof(float)));
    // the frames of the last call sit in ring slots (slot_base_before + 1 + f); use its first n symbols
    const int base_before = ((rx->slot_base - rx->lastF) % rx->nslots + rx->nslots) % rx->nslots;
    dim3 grid((rx->C + 31) / 32, (n + 31) / 32), block(32, 8);
    symbol_power4_kernel<<<grid, block, 0, rx->stream>>>(rx->d_dec_ring, (float2*)bursts.p, 0, rx->C, rx->Cpad, rx->nsym, rx->nslots,
                                                        (base_before + 1) % rx->nslots, n);
    CU(cudaGetLastError());
    int rc = qpsk_b200_fft_argmax_device(rx->est_fft, (const float*)bursts.p, rx->C, (int32_t*)bins.p, (float*)mags.p, rx->stream);
    if (rc) return rc;
    int32_t* hb = new (std::nothrow) int32_t[rx->C];
    if (!hb) return fail(QPSK_B200_ERR_ARG, "out of host memory");
    cudaError_t e = cudaMemcpyAsync(hb, bins.p, (size_t)rx->C * sizeof(int), cudaMemcpyDeviceToHost, rx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(rx->stream);
    if (e != cudaSuccess) { delete[] hb; return fail(QPSK_B200_ERR_CUDA, "estimate download failed: %s", cudaGetErrorString(e)); }
    for (int c = 0; c < rx->C; c++) {
        const int k = hb[c] < n / 2 ? hb[c] : hb[c] - n;              // signed bin of the 4x tone
        h_offset_hz[c] = (float)((double)k * (double)rx->cfg.rs / (4.0 * (double)n));
        if (h_bin) h_bin[c] = hb[c];
    }
    delete[] hb;
    rx->launches += 2;
    return QPSK_B200_OK;
}

// =============================================================================================
// measurement aid: the FP32 pipe's own ceiling on this device (csrc/probe.cuh)
// =============================================================================================
extern "C" int qpsk_b200_probe_fp32(int device, int fused, double* tap_updates_per_s, float* kernel_ms) {
    if (!tap_updates_per_s) return fail(QPSK_B200_ERR_ARG, "null argument");
    int rc = check_device(device);
    if (rc) return rc;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int threads = 256, grid = prop.multiProcessorCount * 4, iters = 4000;      // 8 warps per scheduler, ~3.5 ms per timed launch
    float2 hx[1024];
    ProbeTaps ht;
    for (int i = 0; i < 1024; i++) hx[i] = make_float2(0.001f * (i % 97) - 0.04f, 0.002f * (i % 89) - 0.08f);
    for (int i = 0; i < 128 + 2 * QPSK_PROBE_R; i++) ht.t[i] = make_float2(0.01f * (i % 13) - 0.05f, 0.01f * (i % 13) - 0.05f);
    DevBuf x, o;
    CU(cudaMalloc(&x.p, sizeof hx));
    CU(cudaMalloc(&o.p, (size_t)grid * threads * QPSK_PROBE_R * sizeof(float2)));
    CU(cudaMemcpy(x.p, hx, sizeof hx, cudaMemcpyHostToDevice));
    cudaStream_t s;
    CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    cudaError_t e = cudaSuccess;
    for (int rep = 0; rep < 6 && e == cudaSuccess; rep++) {          // rep 0 is the warm-up
        cudaEventRecord(e0, s);
        if (fused) fp32_pipe_probe_kernel<1><<<grid, threads, 0, s>>>((const float2*)x.p, ht, (float2*)o.p, rep ? iters : 10);
        else       fp32_pipe_probe_kernel<0><<<grid, threads, 0, s>>>((const float2*)x.p, ht, (float2*)o.p, rep ? iters : 10);
        e = cudaGetLastError();
        cudaEventRecord(e1, s);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        float ms = 0.0f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(s);
    if (e != cudaSuccess) return fail(QPSK_B200_ERR_CUDA, "FP32 pipe probe failed: %s", cudaGetErrorString(e));
    *tap_updates_per_s = (double)grid * threads * (double)iters * QPSK_PROBE_TAPS * QPSK_PROBE_R / (best * 1e-3);
    if (kernel_ms) *kernel_ms = best;
    return QPSK_B200_OK;
}
