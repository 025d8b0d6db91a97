"""The headline shape end to end through qpsk_b200_rx_process_host (65,536 channels x 64 frames from pinned host PCM, packed
dibits back to pinned host memory) for QPSK_B200_HOST_CHUNKS = 1, 2, 4, 8 frame chunks per call, next to the copy-only probe.
usage: python tools/e2e_time.py [chunks[:slice half waves] ...]"""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, qpsk_b200
from qpsk_b200 import capi
dev = torch.device("cuda", 0)
C, F = 65536, 64
pcm = bench.synth_pcm_gpu(torch, qpsk_b200, C, F * 512, dev, 0, seed=5)
h_pcm = torch.empty((C, F * 512), dtype=torch.int16).pin_memory()
h_pcm.copy_(pcm); del pcm
h_out = torch.empty((C, F * 128 // 4), dtype=torch.uint8).pin_memory()
L = capi.lib()
ref = None
for chunks, hw in [(a.split(":") + ["1"])[:2] for a in (sys.argv[1:] or ["1", "2", "4", "8"])]:
    os.environ["QPSK_B200_HOST_CHUNKS"] = chunks
    os.environ["QPSK_B200_HOST_SLICE_HALFWAVES"] = hw
    rx = qpsk_b200.Receiver(C, F, device=0, decode_frames=True, estimate_offset=True, transient_symbols=True)
    res = {}
    for name, fn in (("e2e", L.qpsk_b200_rx_process_host), ("copy", L.qpsk_b200_rx_probe_copy_host)):
        for _ in range(2):
            capi.check(fn(rx.h, ctypes.c_void_p(h_pcm.data_ptr()), F, ctypes.c_void_p(h_out.data_ptr())))
        if name == "e2e":
            chk = int(h_out.numpy().astype(np.int64).sum())
        ts = []
        for _ in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            capi.check(fn(rx.h, ctypes.c_void_p(h_pcm.data_ptr()), F, ctypes.c_void_p(h_out.data_ptr())))
            ts.append(time.perf_counter() - t0)
        res[name] = float(np.median(ts)) * 1e3
    if ref is None: ref = chk
    print("host chunks %s, slices of %s half waves: e2e %.2f ms = %.2f Gsamples/s, copy only %.2f ms, ratio %.3f, dibit checksum %d %s" % (
        chunks, hw, res["e2e"], C * F * 512 / res["e2e"] / 1e6, res["copy"], res["copy"] / res["e2e"], chk, "OK" if chk == ref else "MISMATCH"), flush=True)
    rx.close()
