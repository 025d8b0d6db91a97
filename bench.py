#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched QPSK receiver (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the receive hot path (mixer -> RRC matched filter -> timing -> decimate
-> Costas -> slicer) over one batch of synthetic PCM: 65,536 concurrent 2400-baud channels per
GPU x 64 frames of 512 samples (BASELINE.json configs[2]; configs[1] and the others are parity
cases in tests/).  Channels are sharded across ranks with no data-path collective (weak
scaling: every GPU owns its own 65,536 channels); NCCL only sums the statistics counters.

`value` is measured with the PCM already resident in HBM; `e2e` goes through the host-buffer
entry point (qpsk_b200_rx_process_host) with pinned host PCM in and packed dibits out.
`--impl reference` times the unmodified reference C code (oracle/_ref, built from
/root/reference by oracle/Makefile) on the host cores instead.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NCHAN = int(os.environ.get("QPSK_BENCH_NCHAN", 65536))
NFRAMES = int(os.environ.get("QPSK_BENCH_NFRAMES", 64))
FRAME = 512
SPS = 4
NTAPS = 127
METRIC = "complex Msamples/s (decoded Mbit/s = value/2), fused FIR->timing->Costas->slicer, 65,536 x 2400-baud channels per GPU"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


# measured on this pool's B200 by tools/fp32_pipe_bench.cu (profiles/r01_fp32_pipe.md): complex
# tap-updates per clock per SM for the exact (FMUL2.FTZ+FADD2) and fused (FFMA2) formulations
FP32_TAPS_PER_CLK_SM = {"exact": 32.0, "fast": 64.0}


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs (NVML every 20 ms; nvidia-smi as a fallback)."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.sm, self.reasons, self.sm_max = [], set(), None
        self.stop_flag = threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = gpu_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[gpu_index])
                except Exception:
                    idx = gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        r = [x.strip() for x in out.split(",")]
        if len(r) >= 6 and r[0].isdigit():
            self.sm.append(int(r[0]))
            self.sm_max = int(r[1]) if r[1].isdigit() else self.sm_max
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    self.sm.append(self.nvml.nvmlDeviceGetClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
                    mask = self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(self.nvml, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for name, bit in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                else:
                    self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.02 if self.nvml is not None else 0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["clock sampling unavailable"]}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def synth_pcm_gpu(torch, qpsk_b200, nchan, nsamp, device, local, seed, rs=2400.0, esn0_db=20.0):
    """Synthetic QPSK PCM [nchan, nsamp] int16, generated on the GPU: random dibits -> the library's own
    batched transmit path (qpsk_packet_mod/tx_frame semantics, packets of 256 symbols) at a per-channel
    carrier CENTER + U(-75, 75) Hz -> the library's counter-based AWGN at Es/N0 = 20 dB (qpsk_b200_channel_awgn_device)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    sps = int(9600.0 / rs)
    nsym = nsamp // sps
    carriers = (1500.0 + (torch.rand(nchan, generator=g, device=device) * 150.0 - 75.0)).float().cpu().numpy()
    tx = qpsk_b200.Transmitter(carriers, rs=rs, device=local)
    sym = torch.randint(0, 4, (nchan, nsym), generator=g, device=device, dtype=torch.uint8)
    pcm = torch.empty((nchan, nsamp), dtype=torch.int16, device=device)
    tx.modulate_device(sym.data_ptr(), nsym, pcm.data_ptr())
    torch.cuda.synchronize()
    tx.close()
    del sym
    power = pcm[:4096].float().pow(2).mean()
    sigma = float((power * sps / (2.0 * 10.0 ** (esn0_db / 10.0))).sqrt())
    # the library's counter-based test noise (Philox-keyed Irwin-Hall, reproducible sample for sample by the oracle)
    qpsk_b200.awgn_device(pcm.data_ptr(), nchan, nsamp, sigma, seed=seed, device=local)
    torch.cuda.synchronize()
    return pcm


# ------------------------------------------------------------------------------------------------
# CPU side: the unmodified reference (oracle/_ref) on the host cores
# ------------------------------------------------------------------------------------------------
def _ref_worker(args):
    flavour, seed, nframes, budget_s = args
    import numpy as np
    from oracle import Ref
    r = Ref(flavour)
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, size=(nframes * FRAME // (256 * SPS), 512), dtype=np.int32)
    r.tx_reset(1500.0 + float(rng.uniform(-75, 75)))
    pcm = np.concatenate([r.packet_mod(b) for b in bits])[: nframes * FRAME]
    pcm = np.clip(pcm + rng.normal(0, 400, pcm.shape), -32768, 32767).astype(np.int16)
    # calibrate, then run for about budget_s seconds
    t = r.rx_time(pcm, 1)
    reps = max(1, int(budget_s / max(t, 1e-6)))
    t = r.rx_time(pcm, reps)
    return reps * nframes * FRAME, t


def cpu_reference_rate(budget_s=10.0, flavour="2400_O2", nframes=256):
    """Aggregate Msamples/s of the reference rx_frame over all host cores, one process per core."""
    import multiprocessing as mp
    from oracle import Ref
    if not Ref.available(flavour):
        return None
    cores = os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_ref_worker, [(flavour, 1000 + i, nframes, budget_s) for i in range(cores)])
    rate = sum(n / t for n, t in res) / 1e6
    return {"value": rate, "unit": "Msamples/s", "cores": cores, "kind": "reference",
            "sample": "unmodified reference rx_frame (qpsk.c:88-218, gcc -O2 build oracle/_ref/libref_%s.so), one process per "
                      "core, each ~%.0f s over a %d-frame 2400-baud channel" % (flavour, budget_s, nframes)}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    t0 = time.time()
    per_step = []
    base = None
    for _ in range(args.warmup + args.steps):
        base = cpu_reference_rate(budget_s=2.0)
        if base is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (reference tree absent at build time)"}))
            return 0
        per_step.append(base["value"])
    vals = per_step[args.warmup:]
    value = sum(vals) / len(vals)
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * (time.time() - t0) / (args.warmup + args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[2]: 65,536 concurrent 2400-baud channels per GPU x 64 frames x 512 samples (reference arm: the unmodified "
                                   "reference rx_frame, qpsk.c:88-218, on every host core, each step a bounded ~2 s sample of such channels)",
                       "channels_per_gpu": NCHAN, "frames_per_step": NFRAMES},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="exact", choices=["exact", "fast"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: ONE set of 65,536 channels split over the ranks (default: 65,536 channels per GPU, weak)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import qpsk_b200
    from qpsk_b200 import capi

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    W = max(3, args.warmup)
    global NCHAN
    if args.strong and world > 1:
        start, count = qpsk_b200.shard.partition(NCHAN, world, rank)      # contiguous block of the one channel set
        assert count * world == NCHAN, "strong scaling wants the channel count divisible by the ranks"
        NCHAN = count
    # host side of the end-to-end leg: keep this rank's pinned buffers on the GPU's own NUMA node
    numa_cores = qpsk_b200.shard.bind_host_to_gpu(local) if world > 1 else None

    mode = capi.MODE_EXACT if args.mode == "exact" else capi.MODE_FAST
    nsamp = NFRAMES * FRAME
    pcm = synth_pcm_gpu(torch, qpsk_b200, NCHAN, nsamp, dev, local, seed=97 + rank)
    # the full pipeline of configs[2]: ... -> slicer -> descramble/de-interleave/CRC16 per frame
    # plus the FFT frequency estimator as a stage of every call (configs[2] names it; the reference itself never
    # calls its fftn): 65,536 bursts of 1,024 symbols -> 4th power -> FFT -> argmax per step
    rx = qpsk_b200.Receiver(NCHAN, NFRAMES, rs=2400.0, mode=mode, device=local, decode_frames=True,
                            estimate_offset=not bool(int(os.environ.get("QPSK_BENCH_NO_ESTIMATOR", "0"))),
                            no_fuse=bool(int(os.environ.get("QPSK_BENCH_NO_FUSE", "0"))))
    torch.cuda.synchronize()
    # a real (non-default) stream: the C-ABI treats a NULL stream as "the context's own stream", and
    # torch.cuda.Event only sees work on the stream it is recorded on
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    for _ in range(W):
        rx.process_device(pcm.data_ptr(), NFRAMES, stream)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = rx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    front_ms = []
    ev0.record()
    for _ in range(args.steps):
        rx.process_device(pcm.data_ptr(), NFRAMES, stream)
    ev1.record()
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = rx.launch_count() - launches0
    # device time of the dominant kernel: CUDA events recorded by the library around the kernel on the launching
    # stream, read back after each of a second set of identical steps (reading them inside the timed loop would
    # put a host synchronisation between the steps)
    for _ in range(max(3, args.steps)):
        rx.process_device(pcm.data_ptr(), NFRAMES, stream)
        torch.cuda.synchronize()
        front_ms.append(rx.kernel_ms())
    clocks = sampler.summary()
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    samples_per_step = NCHAN * nsamp
    value = world * samples_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- statistics gather (the only collective): symbols decided + mean |freq| per GPU
    track = rx.read(capi.OUT_TRACK)
    nfr, npass = rx.crc_counters()
    locked = int(np.sum(np.abs(np.abs(track[:, -1, 1]) * 2400.0 / (2 * np.pi)) < 80.0))     # |offset| within the +-75 Hz spread
    stats = qpsk_b200.shard.reduce_stats([NCHAN * NFRAMES * (FRAME // SPS), nfr, npass, locked], device=dev)

    # ---- end to end through the host-buffer entry point (pinned PCM in, packed dibits out)
    e2e = None
    if not args.no_e2e:
        h_pcm = torch.empty((NCHAN, nsamp), dtype=torch.int16, pin_memory=True)
        h_pcm.copy_(pcm)
        h_out = torch.empty((NCHAN, NFRAMES * (FRAME // SPS) // 4), dtype=torch.uint8, pin_memory=True)
        L = capi.lib()
        for _ in range(2):
            capi.check(L.qpsk_b200_rx_process_host(rx.h, ctypes.c_void_p(h_pcm.data_ptr()), NFRAMES, ctypes.c_void_p(h_out.data_ptr())))
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            capi.check(L.qpsk_b200_rx_process_host(rx.h, ctypes.c_void_p(h_pcm.data_ptr()), NFRAMES, ctypes.c_void_p(h_out.data_ptr())))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * samples_per_step * args.steps / float(tt.item()) / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(h_pcm.numel() * 2), "d2h_bytes_per_step": int(h_out.numel())}
        del h_pcm, h_out

    if rank == 0:
        hbm_peak, peak_kind = load_peaks()
        k_front = sum(m[0] for m in front_ms) / len(front_ms)
        k_costas = sum(m[1] for m in front_ms) / len(front_ms)
        # algorithmic bytes of the fused pipeline: 2 B PCM in + 2 bits per symbol out per sample
        alg_bytes = samples_per_step * (2.0 + 2.0 / SPS / 8.0)
        achieved = alg_bytes / (k_front * 1e-3) / 1e9
        sm_mhz = clocks.get("sm_mhz") or 1965
        fp_peak = FP32_TAPS_PER_CLK_SM[args.mode] * 148 * sm_mhz * 1e6      # complex tap-updates/s at the sampled clock
        fp_ach = samples_per_step * NTAPS / (k_front * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if (args.strong and world > 1) else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[2]: %d concurrent 2400-baud channels per GPU x %d frames x 512 samples, full mixer->FIR(127 taps)"
                                   "->timing->Costas->slicer->descramble/deinterleave/CRC16 + FFT(1024)/argmax frequency estimator, %s arithmetic" % (NCHAN, NFRAMES, args.mode),
                       "channels_per_gpu": NCHAN, "frames_per_step": NFRAMES, "l2": "inputs (4 GiB PCM per step) exceed L2; no flush needed",
                       "decoded_mbit_s": value / 2.0, "parallelism": "channels sharded over %d GPU(s), no data-path collective" % world},
            "roofline": {"bound": "hbm", "kernel": "rx_front_kernel", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": 10.86e9 * (NCHAN * NFRAMES / (65536.0 * 64.0)), "peak_kind": peak_kind, "kernel_ms": k_front,
                         "note": "kernel is FP32-issue-bound, not HBM-bound (508 flop per 2.06 B): see fp32; traffic = dram read+write of one launch from profiles/r01_rx_front_v6.summary.csv (PCM 4.3 GB in, symbol ring 4.3 GB out, frame scratch spill ~2.4 GB)",
                         "fp32": {"achieved_tap_updates_per_s": fp_ach, "peak_tap_updates_per_s": fp_peak, "frac": fp_ach / fp_peak,
                                  "peak_kind": "2 packed FP32 instr per tap at 64 lanes/clk/SM x 148 SM x sampled SM clock"}},
            "kernels_ms": {"rx_front": k_front, "costas": k_costas},
            "clocks": clocks, "gpu_launches": int(launches),
            "stats": {"symbols": stats[0], "frames_crc_checked": stats[1], "frames_crc_ok": stats[2], "channels_locked": stats[3],
                      "note": "random payload: CRC passes are chance (2^-16); counters show K4 ran over every frame"},
        }
        if e2e is not None:
            if numa_cores is not None:
                e2e["host_cores_rank0"] = "%d cores local to GPU %d (NVML affinity)" % (len(numa_cores), local)
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:      # reported at N = 1 only (rank 0), per the measurement contract
            try:
                cb = cpu_reference_rate(budget_s=8.0)
                if cb is not None:
                    line["cpu_baseline"] = cb
            except Exception as ex:  # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "error": str(ex)}
        print(json.dumps(line))
    rx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
