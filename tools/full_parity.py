#!/usr/bin/env python
"""Config-2 parity over EVERY channel ("all channels once"): the bench's own workload -- 65,536 distinct
2400-baud channels x 64 frames (random carriers, AWGN 20 dB), 2.1 G samples -- through the receiver on the GPU and
through the oracle on the host cores, one process per core; every dibit, timing index and loop (phase, freq) pair
must be bit-identical.  Test infrastructure: this is one of the places allowed to execute oracle/.

usage: python tools/full_parity.py [nchan] [nframes]     (about a minute on a 32-core box)
"""
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

def _scratch_path(nbytes):
    """A file the worker processes can map: /dev/shm when it has room, else the temp directory."""
    import shutil
    import tempfile
    for d in ("/dev/shm", tempfile.gettempdir()):
        try:
            if os.path.isdir(d) and shutil.disk_usage(d).free > nbytes + (256 << 20):
                return os.path.join(d, "qpsk_full_parity_%d.i16" % os.getpid())
        except OSError:
            pass
    raise SystemExit("no scratch space for %d bytes of PCM" % nbytes)


def _worker(args):
    path, c0, c1, nsamp = args
    from oracle import Oracle
    o = Oracle()
    pcm = np.memmap(path, dtype=np.int16, mode="r").reshape(-1, nsamp)
    out = o.rx_run(np.ascontiguousarray(pcm[c0:c1]), want=("index", "dibit", "phase", "freq"))
    d = out["dibit"]
    packed = (d[:, 0::4] | (d[:, 1::4] << 2) | (d[:, 2::4] << 4) | (d[:, 3::4] << 6)).astype(np.uint8)
    return c0, packed, out["index"], out["phase"], out["freq"]


def main():
    import torch
    import qpsk_b200
    from qpsk_b200 import capi
    import bench
    nchan = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    nsamp = nframes * 512
    dev = torch.device("cuda", 0)
    pcm = bench.synth_pcm_gpu(torch, qpsk_b200, nchan, nsamp, dev, 0, seed=97)
    rx = qpsk_b200.Receiver(nchan, nframes, decode_frames=True)
    rx.process_device(pcm.data_ptr(), nframes)
    rx.sync()
    got = rx.read(capi.OUT_DIBITS)
    idx = rx.read(capi.OUT_INDEX)
    track = rx.read(capi.OUT_TRACK)
    # the host-buffer entry point on the same PCM (channel slices x frame chunks, copies overlapped): the same bytes
    host_pcm = pcm.cpu().numpy()
    del pcm
    rx.reset()
    host_got = rx.rx_frames(host_pcm)
    host_bad = int((host_got != got).sum())
    del host_got
    rx.close()
    SHM = _scratch_path(nchan * nsamp * 2)
    host = np.memmap(SHM, dtype=np.int16, mode="w+", shape=(nchan, nsamp))
    host[:] = host_pcm
    host.flush()
    del host_pcm
    cores = os.cpu_count() or 1
    step = max(32, nchan // (cores * 8))
    jobs = [(SHM, c, min(nchan, c + step), nsamp) for c in range(0, nchan, step)]
    t0 = time.time()
    bad = {"dibit_bytes": 0, "index": 0, "phase": 0, "freq": 0, "host_path_dibit_bytes": host_bad}
    try:
        with mp.get_context("spawn").Pool(cores) as pool:
            for c0, packed, windex, wphase, wfreq in pool.imap_unordered(_worker, jobs):
                c1 = c0 + packed.shape[0]
                bad["dibit_bytes"] += int((got[c0:c1] != packed).sum())
                bad["index"] += int((idx[c0:c1] != windex).sum())
                bad["phase"] += int((np.ascontiguousarray(track[c0:c1, :, 0]).view(np.uint32) != wphase.view(np.uint32)).sum())
                bad["freq"] += int((np.ascontiguousarray(track[c0:c1, :, 1]).view(np.uint32) != wfreq.view(np.uint32)).sum())
    finally:
        os.unlink(SHM)
    dt = time.time() - t0
    print({"channels": nchan, "frames": nframes, "symbols": nchan * nframes * 128, "oracle_cores": cores,
           "oracle_seconds": round(dt, 1), "mismatches": bad})
    if any(bad.values()):
        raise SystemExit("PARITY FAILURE")
    print("bit-exact on every channel")


if __name__ == "__main__":
    main()
