"""Streaming ingest of the reference's on-disk format: raw s16le PCM (qpsk.h:14 TX_FILENAME, written by
qpsk.c:331 and read back 512 samples at a time by qpsk.c:348).

`receive_files` is a thin caller of the library's C stream reader (qpsk_b200/host/stream.c): one file per channel,
batches of `frames_per_call` frames read by host threads into page-locked buffers while the GPU works on the previous
batch (qpsk_b200_rx_submit_host / qpsk_b200_rx_wait), all channel state (filter history, mixer phasor, decimation
delay, loop phase/frequency) carried in HBM between batches -- the explicit state block is the resume mechanism
(SURVEY.md section 5)."""
import ctypes as C

import numpy as np

from . import _capi as capi
from .receiver import Receiver, unpack_dibits


def receive_files(paths, rs=2400.0, frames_per_call=64, device=0, on_chunk=None, keep=True, stats=None, **rx_kwargs):
    """Demodulate one raw s16le file per channel.  Returns the decided dibits, uint8 [C, nsym_total]
    (dibit = bits[0] | bits[1] << 1 of qpsk_demod), or None with keep=False (then `on_chunk` is the consumer).
    Trailing samples that do not fill a 512-sample frame are ignored, as in the reference's read loop
    (qpsk.c:350-351).  All files are cut to the shortest one.  `on_chunk(first_frame, dibits [C, nf*nsym], rx)` is called
    for every completed batch; `stats`, a dict, receives the run's timing (frames, seconds, read_seconds, wait_seconds)."""
    L = capi.lib()
    rx = Receiver(len(paths), frames_per_call, rs=rs, device=device, **rx_kwargs)
    st = C.c_void_p()
    arr = (C.c_char_p * len(paths))(*[str(p).encode() for p in paths])
    out, nchan, bytes_per_frame = [], len(paths), rx.nsym // 4
    try:
        capi.check(L.qpsk_b200_stream_open(rx.h, arr, nchan, frames_per_call, rx.frame_size, rx.nsym, C.byref(st)))

        def sink(_user, first_frame, nframes, dibits):
            packed = np.ctypeslib.as_array(dibits, shape=(nchan, nframes * bytes_per_frame))
            if keep or on_chunk is not None:
                chunk = unpack_dibits(packed)
                if on_chunk is not None:
                    on_chunk(int(first_frame), chunk, rx)
                if keep:
                    out.append(chunk)
            return 0

        ss = capi.StreamStats()
        capi.check(L.qpsk_b200_stream_run(st, capi.STREAM_SINK(sink), None, C.byref(ss)))
        if stats is not None:
            stats.update(frames=int(ss.frames), seconds=ss.seconds, read_seconds=ss.read_seconds, wait_seconds=ss.wait_seconds,
                         readers=int(ss.readers), samples_per_s=ss.frames * rx.frame_size * nchan / max(ss.seconds, 1e-12))
    finally:
        if st:
            L.qpsk_b200_stream_close(st)
        rx.close()
    if not keep:
        return None
    return np.concatenate(out, axis=1) if out else np.zeros((nchan, 0), np.uint8)
