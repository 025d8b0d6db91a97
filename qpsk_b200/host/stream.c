/* stream.c -- continuous receiver over raw s16le PCM files, one file per channel: the reference's on-disk format
 * (TX_FILENAME, qpsk.h:14, written by qpsk.c:331) and its read loop (qpsk.c:339-354: fread 512 samples, rx_frame, stop
 * at the first short read), for many channels at once.
 *
 * The reference reads a frame only after rx_frame returned.  Here a batch of frames_per_batch frames of every channel is
 * read into page-locked memory by a few reader threads while the GPU works on the previous batch (two batches in flight
 * through qpsk_b200_rx_submit_host / qpsk_b200_rx_wait), and every completed batch is handed to the caller's sink.  All
 * channel state (filter history, mixer phasor, decimation delay, loop phase and frequency) stays in HBM between
 * batches, so the result is identical to one pass over the whole files.  C11 + pthreads over the C-ABI only. */
#define _POSIX_C_SOURCE 200809L
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "../../include/qpsk_b200.h"

#define NBUF 3          /* two batches in flight + the one being read */
#define MAX_READERS 16

struct qpsk_b200_stream {
    qpsk_b200_rx *rx;
    int nchan, frames_per_batch, frame_size, nsym, nreaders;
    int *fd;
    long long nframes_total;                /* whole frames in the shortest file */
    int16_t *pcm[NBUF];                     /* [C][frames_per_batch * frame_size], page-locked */
    uint8_t *dibits[NBUF];                  /* [C][frames_per_batch * nsym / 4], page-locked */
    char err[256];
};

int qpsk_b200_stream_set_error(const char *text);   /* qpsk_b200.cu: makes qpsk_b200_last_error() return `text` */

static int stream_fail(qpsk_b200_stream *st, int code, const char *what, const char *arg) {
    snprintf(st->err, sizeof st->err, "%s%s%s", what, arg ? ": " : "", arg ? arg : "");
    qpsk_b200_stream_set_error(st->err);
    return code;
}

int qpsk_b200_stream_close(qpsk_b200_stream *st) {
    if (!st) return QPSK_B200_OK;
    if (st->fd) {
        for (int c = 0; c < st->nchan; c++)
            if (st->fd[c] >= 0) close(st->fd[c]);
        free(st->fd);
    }
    for (int b = 0; b < NBUF; b++) {
        if (st->pcm[b]) qpsk_b200_host_free(st->pcm[b]);
        if (st->dibits[b]) qpsk_b200_host_free(st->dibits[b]);
    }
    free(st);
    return QPSK_B200_OK;
}

int qpsk_b200_stream_open(qpsk_b200_rx *rx, const char *const *paths, int nchan, int frames_per_batch, int frame_size, int nsym,
                          qpsk_b200_stream **out) {
    if (!rx || !paths || !out || nchan < 1 || frames_per_batch < 1 || frame_size < 1 || nsym < 4) {
        qpsk_b200_stream_set_error("qpsk_b200_stream_open: bad argument");
        return QPSK_B200_ERR_ARG;
    }
    *out = NULL;
    qpsk_b200_stream *st = calloc(1, sizeof *st);
    if (!st) { qpsk_b200_stream_set_error("out of host memory"); return QPSK_B200_ERR_ARG; }
    st->rx = rx; st->nchan = nchan; st->frames_per_batch = frames_per_batch; st->frame_size = frame_size; st->nsym = nsym;
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    st->nreaders = ncpu > 1 ? (int)(ncpu / 2) : 1;
    if (st->nreaders > MAX_READERS) st->nreaders = MAX_READERS;
    if (st->nreaders > nchan) st->nreaders = nchan;
    st->fd = malloc(sizeof(int) * (size_t)nchan);
    if (!st->fd) { qpsk_b200_stream_close(st); qpsk_b200_stream_set_error("out of host memory"); return QPSK_B200_ERR_ARG; }
    for (int c = 0; c < nchan; c++) st->fd[c] = -1;
    long long shortest = -1;
    for (int c = 0; c < nchan; c++) {
        st->fd[c] = open(paths[c], O_RDONLY);
        struct stat sb;
        if (st->fd[c] < 0 || fstat(st->fd[c], &sb) != 0) {
            int rc = stream_fail(st, QPSK_B200_ERR_ARG, "cannot open PCM file", paths[c]);
            qpsk_b200_stream_close(st);
            return rc;
        }
        const long long frames = (long long)sb.st_size / ((long long)frame_size * 2);   /* a ragged tail is ignored, qpsk.c:350-351 */
        if (shortest < 0 || frames < shortest) shortest = frames;
    }
    st->nframes_total = shortest;
    const size_t pcm_bytes = (size_t)nchan * frames_per_batch * frame_size * sizeof(int16_t);
    const size_t out_bytes = (size_t)nchan * frames_per_batch * (nsym / 4);
    for (int b = 0; b < NBUF; b++) {
        if (qpsk_b200_host_alloc(pcm_bytes, (void **)&st->pcm[b]) != 0 || qpsk_b200_host_alloc(out_bytes, (void **)&st->dibits[b]) != 0) {
            qpsk_b200_stream_close(st);
            return QPSK_B200_ERR_CUDA;      /* the library set the text */
        }
    }
    *out = st;
    return QPSK_B200_OK;
}

long long qpsk_b200_stream_frames(const qpsk_b200_stream *st) { return st ? st->nframes_total : 0; }

struct reader_job {
    qpsk_b200_stream *st;
    int16_t *dst;
    long long first_frame;
    int nframes, c_begin, c_end, failed;
};

static void *reader_main(void *arg) {
    struct reader_job *j = arg;
    const qpsk_b200_stream *st = j->st;
    const size_t row = (size_t)j->nframes * st->frame_size;              /* samples per channel in this batch */
    for (int c = j->c_begin; c < j->c_end; c++) {
        char *p = (char *)(j->dst + (size_t)c * row);
        size_t left = row * 2;
        off_t off = (off_t)j->first_frame * st->frame_size * 2;
        while (left > 0) {
            ssize_t n = pread(st->fd[c], p, left, off);
            if (n <= 0) { if (n < 0 && errno == EINTR) continue; j->failed = 1; return NULL; }
            p += n; off += n; left -= (size_t)n;
        }
    }
    return NULL;
}

static int read_batch(qpsk_b200_stream *st, int16_t *dst, long long first_frame, int nframes) {
    pthread_t th[MAX_READERS];
    struct reader_job job[MAX_READERS];
    const int T = st->nreaders;
    for (int t = 0; t < T; t++) {
        job[t].st = st; job[t].dst = dst; job[t].first_frame = first_frame; job[t].nframes = nframes; job[t].failed = 0;
        job[t].c_begin = (int)((long long)st->nchan * t / T);
        job[t].c_end = (int)((long long)st->nchan * (t + 1) / T);
    }
    int started = 0, bad = 0;
    for (int t = 1; t < T; t++) {
        if (pthread_create(&th[t], NULL, reader_main, &job[t]) != 0) break;
        started = t;
    }
    for (int t = started + 1; t < T; t++) reader_main(&job[t]);        /* threads that could not start: read here */
    reader_main(&job[0]);
    for (int t = 1; t <= started; t++) pthread_join(th[t], NULL);
    for (int t = 0; t < T; t++) bad |= job[t].failed;
    return bad ? stream_fail(st, QPSK_B200_ERR_ARG, "short read from a PCM file", NULL) : 0;
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int qpsk_b200_stream_run(qpsk_b200_stream *st, qpsk_b200_stream_sink sink, void *user, qpsk_b200_stream_stats *stats) {
    if (!st) { qpsk_b200_stream_set_error("null stream"); return QPSK_B200_ERR_ARG; }
    const int Fb = st->frames_per_batch;
    const long long total = st->nframes_total;
    const long long nbatches = (total + Fb - 1) / Fb;
    double t_read = 0.0, t_wait = 0.0;
    const double t0 = now_s();
    int rc = 0;
    /* step k: read batch k while the GPU works on batch k-1, submit it (two in flight), then wait for batch k-1 and
     * hand it to the sink */
    for (long long k = 0; k < nbatches + 1 && rc == 0; k++) {
        if (k < nbatches) {
            const long long f0 = k * Fb;
            const int nf = (int)((total - f0 < Fb) ? total - f0 : Fb);
            const int b = (int)(k % NBUF);
            double t = now_s();
            rc = read_batch(st, st->pcm[b], f0, nf);
            t_read += now_s() - t;
            if (rc == 0) rc = qpsk_b200_rx_submit_host(st->rx, st->pcm[b], nf, st->dibits[b]);
        }
        const long long done = k - 1;
        if (rc == 0 && done >= 0 && done < nbatches) {
            double t = now_s();
            rc = qpsk_b200_rx_wait(st->rx);
            t_wait += now_s() - t;
            if (rc == 0 && sink) {
                const long long f0 = done * Fb;
                const int nf = (int)((total - f0 < Fb) ? total - f0 : Fb);
                rc = sink(user, f0, nf, st->dibits[done % NBUF]);
                if (rc != 0) stream_fail(st, rc, "the sink asked to stop", NULL);
            }
        }
    }
    if (rc != 0) {          /* nothing of ours may still be in flight when the caller gets the error */
        qpsk_b200_rx_wait(st->rx);
        qpsk_b200_rx_wait(st->rx);
    }
    if (stats) {
        stats->frames = total;
        stats->seconds = now_s() - t0;
        stats->read_seconds = t_read;
        stats->wait_seconds = t_wait;
        stats->readers = st->nreaders;
    }
    return rc;
}
