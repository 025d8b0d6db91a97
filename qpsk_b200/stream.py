"""Streaming ingest of the reference's on-disk format: raw s16le PCM (qpsk.h:14 TX_FILENAME, written by
qpsk.c:331 and read back 512 samples at a time by qpsk.c:348).

`receive_files` turns the batch receiver into a continuous one: each file is one channel, every call to the
library consumes up to `frames_per_call` frames per channel from page-locked staging buffers, and all channel
state (filter history, mixer phasor, decimation delay, loop phase/frequency) stays in HBM between calls -- the
explicit state block is the resume mechanism."""
import numpy as np

from .receiver import Receiver, unpack_dibits


def receive_files(paths, rs=2400.0, frames_per_call=64, device=0, on_chunk=None, **rx_kwargs):
    """Demodulate one raw s16le file per channel.  Returns the decided dibits, uint8 [C, nsym_total]
    (dibit = bits[0] | bits[1] << 1 of qpsk_demod).  Trailing samples that do not fill a 512-sample frame are
    ignored, as in the reference's read loop (qpsk.c:350-351).  All files are cut to the shortest one."""
    files = [np.memmap(p, dtype="<i2", mode="r") for p in paths]
    nframes_total = min(len(f) for f in files) // 512
    rx = Receiver(len(files), frames_per_call, rs=rs, device=device, **rx_kwargs)
    try:
        import torch
        staging = torch.empty((len(files), frames_per_call * 512), dtype=torch.int16).pin_memory().numpy()
    except Exception:       # pinned memory is an optimisation, not a requirement
        staging = np.empty((len(files), frames_per_call * 512), np.int16)
    out = []
    done = 0
    while done < nframes_total:
        nf = min(frames_per_call, nframes_total - done)
        view = staging[:, :nf * 512]
        for c, f in enumerate(files):
            view[c] = f[done * 512:(done + nf) * 512]
        chunk = unpack_dibits(rx.rx_frames(np.ascontiguousarray(view)))
        if on_chunk is not None:
            on_chunk(done, chunk, rx)
        out.append(chunk)
        done += nf
    rx.close()
    return np.concatenate(out, axis=1) if out else np.zeros((len(files), 0), np.uint8)
