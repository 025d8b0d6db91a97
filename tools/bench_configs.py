#!/usr/bin/env python
"""Timing of the BASELINE.json configurations other than the headline (which is bench.py):

    config 1  1,024 x 1200-baud channels (10 m profile), AWGN + carrier offsets, one B200
    config 3  long-tap RRC stress: 256 taps, 8 samples/symbol, 16,384 channels (FMA-bound regime)
    config 4  FFT timing/frequency estimator sweep: batched 256..8192-point FFT + argmax
    plus the transmit path and the bit stages

Each line printed is one JSON record (CUDA-event timings, inputs resident in HBM, >= 3 warm-ups,
working sets larger than L2 or an explicit L2 flush between iterations).  Output of a run is kept
under profiles/.  Usage: python tools/bench_configs.py [--quick]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import qpsk_b200  # noqa: E402
from qpsk_b200 import capi  # noqa: E402

HBM_PEAK = 6550.7
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    HBM_PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", HBM_PEAK)
SM, CLK = 148, 1.965e9


def flush_l2(buf):
    buf.add_(1)


def timed(fn, iters, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        if flush is not None:
            flush_l2(flush)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sh = stream.cuda_stream
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)      # 256 MiB > 126 MB L2
    out = []

    # ---- config 1: 1,024 x 1200-baud channels, F = 256 frames ------------------------------------
    for mode_name, mode in (("exact", capi.MODE_EXACT), ("fast", capi.MODE_FAST)):
        Cn, F = 1024, 256
        sps, nsym = 8, F * 512 // 8
        g = torch.Generator(device=dev); g.manual_seed(1)
        carriers = (1500.0 + (torch.rand(Cn, generator=g, device=dev) * 150 - 75)).cpu().numpy()
        tx = qpsk_b200.Transmitter(carriers, rs=1200.0)
        sym = torch.randint(0, 4, (Cn, nsym), generator=g, device=dev, dtype=torch.uint8)
        pcm = torch.empty((Cn, F * 512), dtype=torch.int16, device=dev)
        tx.modulate_device(sym.data_ptr(), nsym, pcm.data_ptr(), sh)
        torch.cuda.synchronize()
        tx.close()
        rx = qpsk_b200.Receiver(Cn, F, rs=1200.0, mode=mode, decode_frames=True)
        ms = timed(lambda: rx.process_device(pcm.data_ptr(), F, sh), 5, flush)
        k1, k3 = rx.kernel_ms()
        samples = Cn * F * 512
        out.append({"config": "1: 1,024 x 1200-baud channels x 256 frames, full pipeline, %s" % mode_name, "ms": ms,
                    "msamples_s": samples / ms / 1e3, "decoded_mbit_s": samples / ms / 1e3 * 2 / sps,
                    "kernels_ms": {"rx_front": k1, "costas": k3}, "l2": "flushed between iterations",
                    "fp32_frac": samples * 127 / (k1 * 1e-3) / ((32 if mode_name == "exact" else 64) * SM * CLK)})
        rx.close()
        del pcm, sym

    # ---- the 1200-baud profile at the headline scale: 65,536 channels x 64 frames (sps 8, 64 symbols per frame) ----
    if not args.quick:
        Cn, F = 65536, 64
        nsym = F * 512 // 8
        g = torch.Generator(device=dev); g.manual_seed(2)
        carriers = (1500.0 + (torch.rand(Cn, generator=g, device=dev) * 150 - 75)).cpu().numpy()
        tx = qpsk_b200.Transmitter(carriers, rs=1200.0)
        sym = torch.randint(0, 4, (Cn, nsym), generator=g, device=dev, dtype=torch.uint8)
        pcm = torch.empty((Cn, F * 512), dtype=torch.int16, device=dev)
        tx.modulate_device(sym.data_ptr(), nsym, pcm.data_ptr(), sh)
        torch.cuda.synchronize()
        tx.close()
        del sym
        rx = qpsk_b200.Receiver(Cn, F, rs=1200.0, decode_frames=True)
        ms = timed(lambda: rx.process_device(pcm.data_ptr(), F, sh), 5)
        k1, k3 = rx.kernel_ms()
        samples = Cn * F * 512
        out.append({"config": "1 at scale: 65,536 x 1200-baud channels x 64 frames, full pipeline, exact", "ms": ms,
                    "msamples_s": samples / ms / 1e3, "decoded_mbit_s": samples / ms / 1e3 * 2 / 8,
                    "kernels_ms": {"rx_front": k1, "costas": k3}, "l2": "4 GiB of PCM per step exceeds L2",
                    "fp32_frac": samples * 127 / (k1 * 1e-3) / (32 * SM * CLK)})
        rx.close()
        del pcm

    # ---- config 3: 256-tap FIR, 16,384 channels x T samples, in place -------------------------------
    for ntaps, rs in ((256, 1200.0), (127, 2400.0)):
        for mode_name, mode in (("exact", capi.MODE_EXACT), ("fast", capi.MODE_FAST)):
            Cn, T = 16384, (8192 if args.quick else 65536)      # 8 GiB filtered in place
            taps = qpsk_b200.rrc_make(ntaps, 9600.0, rs, 0.35)
            x = torch.randn((Cn, T, 2), device=dev, dtype=torch.float32)
            f = qpsk_b200.Fir(taps, Cn, mode=mode)
            ms = timed(lambda: f.filter_device(x.data_ptr(), T, sh), 3)
            km = f.kernel_ms()
            samples = Cn * T
            out.append({"config": "3: rrc_fir %d taps, 16,384 channels x %d complex samples in place, %s" % (ntaps, T, mode_name),
                        "ms": ms, "kernel_ms": km, "msamples_s": samples / ms / 1e3,
                        "hbm_gbs": samples * 16 / (km * 1e-3) / 1e9, "hbm_frac": samples * 16 / (km * 1e-3) / 1e9 / HBM_PEAK,
                        "fp32_frac": samples * ntaps / (km * 1e-3) / ((32 if mode_name == "exact" else 64) * SM * CLK),
                        "l2": "working set %.1f GiB > L2" % (samples * 8 / 2 ** 30)})
            f.close()
            del x

    # ---- config 3 end to end: host-buffer entry point, page-locked samples in place (sliced + pipelined, PCIe-bound) ----
    if not args.quick:
        import time
        Cn, T = 16384, 16384                                          # 2 GiB in, 2 GiB out
        taps = qpsk_b200.rrc_make(127, 9600.0, 2400.0, 0.35)
        h = torch.randn((Cn, T, 2), dtype=torch.float32).pin_memory()
        f = qpsk_b200.Fir(taps, Cn)
        L = capi.lib()
        for _ in range(2):
            capi.check(L.qpsk_b200_fir_process_host(f.h, C.c_void_p(h.data_ptr()), T))
        t0 = time.perf_counter()
        for _ in range(3):
            capi.check(L.qpsk_b200_fir_process_host(f.h, C.c_void_p(h.data_ptr()), T))
        ms = (time.perf_counter() - t0) / 3 * 1e3
        out.append({"config": "3 end to end: rrc_fir 127 taps, 16,384 x 16,384 complex samples in page-locked host memory, in place", "ms": ms,
                    "msamples_s": Cn * T / ms / 1e3, "pcie_gbs_each_way": Cn * T * 8 / (ms * 1e-3) / 1e9})
        f.close()
        del h

    # ---- config 4: FFT + argmax sweep, 131,072 bursts per GPU (= 1 M bursts over 8 GPUs) -------------
    for n in (256, 512, 1024, 2048, 4096, 8192):
        nb = 131072 if not args.quick else 16384
        x = torch.randn((nb, n, 2), device=dev, dtype=torch.float32)
        bins = torch.empty(nb, dtype=torch.int32, device=dev)
        mag = torch.empty(nb, dtype=torch.float32, device=dev)
        f = qpsk_b200.Fft(n)
        ms = timed(lambda: f.argmax_device(x.data_ptr(), nb, bins.data_ptr(), mag.data_ptr(), sh), 5, flush if nb * n * 8 < (256 << 20) else None)
        km = f.kernel_ms()
        by = nb * (8 * n + 8)
        out.append({"config": "4: FFT+argmax n=%d, %d bursts" % (n, nb), "ms": ms, "kernel_ms": km, "mpoints_s": nb * n / km / 1e3,
                    "hbm_gbs": by / (km * 1e-3) / 1e9, "hbm_frac": by / (km * 1e-3) / 1e9 / HBM_PEAK,
                    "gflops": 5.0 * n * np.log2(n) * nb / (km * 1e-3) / 1e9})
        f.close()
        del x

    # ---- transmit path: 65,536 channels x 8,192 symbols (the bench's synthetic-signal generator) -----
    Cn, nsym = (65536, 8192) if not args.quick else (8192, 4096)
    carriers = np.full(Cn, 1550.0, np.float32)
    tx = qpsk_b200.Transmitter(carriers)
    sym = torch.randint(0, 4, (Cn, nsym), device=dev, dtype=torch.uint8)
    pcm = torch.empty((Cn, nsym * 4), dtype=torch.int16, device=dev)
    ms = timed(lambda: tx.modulate_device(sym.data_ptr(), nsym, pcm.data_ptr(), sh), 3)
    out.append({"config": "tx: %d channels x %d symbols (2400 baud)" % (Cn, nsym), "ms": ms, "msamples_s": Cn * nsym * 4 / ms / 1e3,
                "hbm_gbs": Cn * nsym * 9 / (ms * 1e-3) / 1e9})
    tx.close()
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
