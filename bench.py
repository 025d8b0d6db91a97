#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched QPSK receiver (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the receive hot path (mixer -> RRC matched filter -> timing -> decimate
-> Costas -> slicer -> descramble/de-interleave/CRC16, + the FFT frequency estimator) over one batch of
synthetic PCM: 65,536 concurrent 2400-baud channels per GPU x 64 frames of 512 samples
(BASELINE.json configs[2]).  Channels are sharded across ranks with no data-path collective (weak
scaling: every GPU owns its own 65,536 channels); NCCL only sums the statistics counters.

`value` is measured with the PCM already resident in HBM; `e2e` goes through the host-buffer
entry point (qpsk_b200_rx_process_host) with pinned host PCM in and packed dibits out, next to the
same copies with no kernels (`e2e.copy_only`, the ingest ceiling of the same run).

The line also carries `configs`: one clock-stamped record per other BASELINE.json configuration
(configs[0] the reference's own binary, configs[1] 1,024 x 1200 baud, configs[3] 256-tap rrc_fir,
configs[4] the FFT + argmax sweep, sharded over the ranks), each with its own roofline and CPU baseline.
At N > 1 only the sharded configurations (the headline and configs[4]) run.

`--impl reference` times the unmodified reference C code (oracle/_ref, built from
/root/reference by oracle/Makefile) on the host cores instead.
"""
import argparse
import csv
import ctypes
import glob
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
WORKLOAD = ("configs[2]: 65,536 concurrent 2400-baud channels per GPU x 64 frames x 512 samples, full mixer->FIR(127 taps)->timing->Costas->slicer"
            "->descramble/deinterleave/CRC16 pipeline")
# the decimated symbols are a hand-off between the timing stage and the loop, not an output of the headline pipeline: their ring slots
# are dropped from L2 once consumed (QPSK_B200_TRANSIENT_SYMBOLS; every decision is unchanged, tests/test_rx_parity_gpu.py)
TRANSIENT = bool(int(os.environ.get("QPSK_BENCH_TRANSIENT", "1")))
FFT_SIZES = (256, 512, 1024, 2048, 4096, 8192)
FFT_BURSTS_PER_GPU = (1 << 20) // 8          # configs[4]: 1 M bursts over 8 GPUs


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json, burst copy figure)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(kernel_prefix, pattern):
    """DRAM read+write bytes per launch of the newest committed ncu summary matching profiles/<pattern>
    (one `ncu --set full` capture, condensed by tools/ncu_summary.py).  Returns (bytes, file) or (None, None)."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for path in reversed(files):
        tot, seen = 0.0, 0
        try:
            with open(path) as f:
                for row in csv.DictReader(f):
                    if kernel_prefix in row["kernel"] and row["metric"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                        tot += float(row["value"]) * unit.get(row["unit"], 1.0)
                        seen += 1
        except Exception:
            continue
        if seen >= 2:
            return tot, os.path.relpath(path, ROOT)
    return None, None


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs (NVML every 10 ms; nvidia-smi as a fallback)."""

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

    def _sample(self):
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

    def run(self):
        while not self.stop_flag.is_set():
            self._sample()
            self.stop_flag.wait(0.01 if self.nvml is not None else 0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["clock sampling unavailable"]}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def synth_pcm_gpu(torch, qpsk_b200, nchan, nsamp, device, local, seed, rs=2400.0, esn0_db=20.0, doppler_hz_per_s=0.0):
    """Synthetic QPSK PCM [nchan, nsamp] int16, generated on the GPU: random dibits -> the library's own
    batched transmit path (qpsk_packet_mod/tx_frame semantics, packets of 256 symbols) at a per-channel
    carrier CENTER + U(-75, 75) Hz (optionally stepped along a linear Doppler ramp, one step per 2,048 symbols)
    -> the library's counter-based AWGN at Es/N0 = 20 dB (qpsk_b200_channel_awgn_device)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    sps = int(9600.0 / rs)
    nsym = nsamp // sps
    carriers = (1500.0 + (torch.rand(nchan, generator=g, device=device) * 150.0 - 75.0)).float().cpu().numpy()
    tx = qpsk_b200.Transmitter(carriers, rs=rs, device=local)
    sym = torch.randint(0, 4, (nchan, nsym), generator=g, device=device, dtype=torch.uint8)
    pcm = torch.empty((nchan, nsamp), dtype=torch.int16, device=device)
    if doppler_hz_per_s == 0.0:
        tx.modulate_device(sym.data_ptr(), nsym, pcm.data_ptr())
    else:
        import numpy as np
        rate = (torch.rand(nchan, generator=g, device=device) * 2.0 - 1.0).float().cpu().numpy() * doppler_hz_per_s
        step = 2048
        sym_t = sym.view(nchan, nsym // step, step).transpose(0, 1).contiguous()      # [steps][C][step]
        tmp = torch.empty((nchan, step * sps), dtype=torch.int16, device=device)
        for k in range(nsym // step):
            tx.set_carrier((carriers + rate * (k * step / rs)).astype(np.float32))
            tx.modulate_device(sym_t[k].data_ptr(), step, tmp.data_ptr())
            torch.cuda.synchronize()
            pcm[:, k * step * sps:(k + 1) * step * sps] = tmp
        del sym_t, tmp
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
# CPU side: the unmodified reference (oracle/_ref) on the host cores, one pinned process per core
# ------------------------------------------------------------------------------------------------
def _pin(core):
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass


def _cpu_worker(task):
    """One host core's share of a CPU baseline.  task = (kind, core, payload, budget_s) -> (units, seconds)."""
    kind, core, payload, budget_s = task
    _pin(core)
    import numpy as np
    from oracle import Ref, RefAlg
    if kind == "rx":                       # unmodified rx_frame (qpsk.c:88-218) over one channel's PCM
        flavour, pcm = payload
        r = Ref(flavour)
        t = r.rx_time(pcm, 1)
        reps = max(1, int(budget_s / max(t, 1e-6)))
        t = r.rx_time(pcm, reps)
        return reps * len(pcm), t
    if kind == "fir":                      # unmodified rrc_fir (rrc_fir.c:17-30), NTAPS = 256 flavour
        flavour, nsamp, seed = payload
        r = Ref(flavour)
        rng = np.random.default_rng(seed)
        x = (rng.standard_normal(nsamp) + 1j * rng.standard_normal(nsamp)).astype(np.complex64)
        mem = np.zeros(r.ntaps, np.complex64)
        done, t0 = 0, time.perf_counter()
        while True:
            buf = x.copy()
            r.fir(mem, buf)
            done += nsamp
            if time.perf_counter() - t0 >= budget_s:
                break
        return done, time.perf_counter() - t0
    if kind == "fft":                      # unmodified fftn (fft.c:110-120), complex double, + the argmax in numpy
        n, seed = payload
        a = RefAlg("alg_O2")
        rng = np.random.default_rng(seed)
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex128)
        out = np.zeros_like(x)
        fn, xin, xout = a.L.fftn, x.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p)
        done, t0 = 0, time.perf_counter()
        while True:
            for _ in range(8):
                fn(xin, xout, n)
                int(np.argmax(out.real * out.real + out.imag * out.imag))
            done += 8
            if time.perf_counter() - t0 >= budget_s:
                break
        return done, time.perf_counter() - t0
    raise ValueError(kind)


class CpuPool:
    """All host cores this process may use, one spawned worker pinned to each (sched_setaffinity)."""

    def __init__(self):
        import multiprocessing as mp
        try:
            self.cores = sorted(os.sched_getaffinity(0))
        except Exception:
            self.cores = list(range(os.cpu_count() or 1))
        self.pool = mp.get_context("spawn").Pool(len(self.cores))

    def rate(self, kind, payloads, budget_s):
        """payloads: one per core (cycled).  Returns aggregate units per second."""
        tasks = [(kind, c, payloads[i % len(payloads)], budget_s) for i, c in enumerate(self.cores)]
        res = self.pool.map(_cpu_worker, tasks, chunksize=1)
        return sum(n / t for n, t in res)

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_synth_channels(nchan, nframes, flavour, seed):
    """Reference-side synthetic PCM (no GPU): the reference's own qpsk_packet_mod + numpy noise, one row per channel."""
    import numpy as np
    from oracle import Ref
    r = Ref(flavour)
    rows = []
    for c in range(nchan):
        rng = np.random.default_rng(seed + c)
        bits = rng.integers(0, 2, size=(nframes * FRAME // (256 * r.sps), 512), dtype=np.int32)
        r.tx_reset(1500.0 + float(rng.uniform(-75, 75)))
        pcm = np.concatenate([r.packet_mod(b) for b in bits])[: nframes * FRAME]
        rows.append(np.clip(pcm + rng.normal(0, 400, pcm.shape), -32768, 32767).astype(np.int16))
    return np.stack(rows)


def cpu_rx_baseline(pool, pcm_rows, flavour, budget_s, what):
    """Aggregate Msamples/s of the unmodified rx_frame over all host cores; pcm_rows int16 [k][T], one row per worker (cycled)."""
    from oracle import Ref
    if not Ref.available(flavour):
        return None
    rate = pool.rate("rx", [(flavour, row) for row in pcm_rows], budget_s) / 1e6
    return {"value": rate, "unit": "Msamples/s", "cores": len(pool.cores), "kind": "reference",
            "sample": "unmodified reference rx_frame (qpsk.c:88-218, oracle/_ref/libref_%s.so), one process pinned to each of %d cores, each ~%.0f s "
                      "over %s" % (flavour, len(pool.cores), budget_s, what)}


def headline_cpu_baseline(pool, pcm_rows, budget_s, what):
    cb = cpu_rx_baseline(pool, pcm_rows, "2400_O2", budget_s, what)
    if cb is None:
        return None
    cb["build"] = "gcc -O2"
    shipped = cpu_rx_baseline(pool, pcm_rows, "2400", max(2.0, budget_s / 2), what)
    if shipped is not None:      # the reference's own Makefile flags (Makefile:7: -std=c11, no -O)
        cb["as_shipped"] = {"value": shipped["value"], "unit": "Msamples/s", "build": "reference Makefile:7 flags (-std=c11, no -O)"}
    return cb


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    import numpy as np
    from oracle import Ref
    if not Ref.available("2400_O2"):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (reference tree absent at build time)"}))
        return 0
    t0 = time.time()
    pool = CpuPool()
    ncores = len(pool.cores)
    # the GPU arm's own PCM (same generator, same seed): its first `cores` channels, one per worker
    rows, what = None, None
    try:
        import torch
        if torch.cuda.is_available():
            import qpsk_b200
            dev = torch.device("cuda", 0)
            pcm = synth_pcm_gpu(torch, qpsk_b200, NCHAN, NFRAMES * FRAME, dev, 0, seed=97)
            rows = pcm[:ncores].cpu().numpy()
            del pcm
            torch.cuda.empty_cache()
            what = "the first %d of the GPU arm's 65,536 channels (same generator and seed), 64 frames each, repeated" % ncores
    except Exception:
        rows = None
    if rows is None:
        rows = cpu_synth_channels(min(ncores, 8), 256, "2400_O2", 1000)
        what = "256-frame 2400-baud channels synthesised on the CPU with the reference's own transmit path"
    per_step, base = [], None
    for _ in range(args.warmup + args.steps):
        base = cpu_rx_baseline(pool, rows, "2400_O2", 2.0, what)
        per_step.append(base["value"])
    vals = per_step[args.warmup:]
    value = sum(vals) / len(vals)
    base["value"] = value
    base["build"] = "gcc -O2"
    shipped = cpu_rx_baseline(pool, rows, "2400", 2.0, what)
    if shipped is not None:
        base["as_shipped"] = {"value": shipped["value"], "unit": "Msamples/s", "build": "reference Makefile:7 flags (-std=c11, no -O)"}
    pool.close()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * (time.time() - t0) / (args.warmup + args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "channels_per_gpu": NCHAN, "frames_per_step": NFRAMES,
                       "arm": "the unmodified reference rx_frame on every host core; each step is a bounded ~2 s sample of the workload's channels"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU side helpers
# ------------------------------------------------------------------------------------------------
class Ctx:
    pass


def timed_loop(ctx, fn, min_seconds=0.25, min_iters=5, warm=3):
    """W >= 3 warm-ups, then >= min_iters back-to-back calls for at least min_seconds, bracketed by CUDA events on the
    launching stream, with the SM clock sampled meanwhile.  Returns (ms per call, calls, clocks)."""
    torch = ctx.torch
    for _ in range(max(3, warm)):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    one = max(e0.elapsed_time(e1), 1e-3)
    iters = int(max(min_iters, min(20000, min_seconds * 1e3 / one)))
    sampler = ClockSampler(ctx.local)
    sampler.start()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    sampler._sample()
    clocks = sampler.summary()
    return e0.elapsed_time(e1) / iters, iters, clocks


def fp32_roofline(ctx, kernel, tap_updates, kernel_ms, mode):
    """FP32-pipe roofline: achieved complex tap-updates/s of `kernel` against the probe's measured ceiling (4 flop each)."""
    peak = ctx.fp32_peak[mode]
    ach = tap_updates / (kernel_ms * 1e-3)
    return {"bound": "fp32", "kernel": kernel, "achieved": ach * 4 / 1e12, "peak": peak * 4 / 1e12, "unit": "TFLOP/s", "frac": ach / peak,
            "kernel_ms": kernel_ms,
            "peak_kind": "measured in this run: qpsk_b200_probe_fp32 (%s), a kernel of nothing but the filter's multiply+add pairs in the "
                         "steady loop's shape (16 accumulators, one shared-memory load per 32 packed instructions, taps from the constant "
                         "bank, 8 warps per scheduler): 0.95 of the nominal 128 lanes/clk/SM, event-timed; 1 complex tap-update = 4 flop"
                         % ("FFMA2" if mode == "fast" else "FMUL2+FADD2, the reference's unfused arithmetic")}


def bench_config1(ctx, pool, want_cpu):
    """configs[1]: 1,024 independent 1200-baud channels (10 m profile), AWGN + Doppler, one B200."""
    torch, qpsk_b200, capi = ctx.torch, ctx.qpsk_b200, ctx.capi
    Cn, F, sps = 1024, 256, 8
    nsamp = F * FRAME
    pcm = synth_pcm_gpu(torch, qpsk_b200, Cn, nsamp, ctx.dev, ctx.local, seed=11, rs=1200.0, doppler_hz_per_s=5.0)
    rx = qpsk_b200.Receiver(Cn, F, rs=1200.0, device=ctx.local, decode_frames=True)
    flush = ctx.flush

    def step():
        flush.add_(1)                                       # 256 MiB of PCM is twice the L2: flush between iterations all the same
        rx.process_device(pcm.data_ptr(), F, ctx.stream)

    def flush_only():
        flush.add_(1)
    ms_all, iters, clocks = timed_loop(ctx, step)
    ms_flush, _, _ = timed_loop(ctx, flush_only)
    ms = ms_all - ms_flush
    rx.process_device(pcm.data_ptr(), F, ctx.stream)
    torch.cuda.synchronize()
    k_front, k_loop = rx.kernel_ms()
    samples = Cn * nsamp
    rec = {"config": "configs[1]", "workload": "1,024 x 1200-baud channels (sps 8, 64 symbols/frame) x 256 frames, Es/N0 20 dB + carrier offsets + linear Doppler "
                                              "(<= 5 Hz/s), full pipeline, exact arithmetic",
           "ms": ms, "iters": iters, "value": samples / ms / 1e3, "unit": "Msamples/s", "decoded_mbit_s": samples / ms / 1e3 * 2 / sps,
           "kernels_ms": {"rx_front_sum_over_chunks": k_front, "costas_sum_over_chunks": k_loop},
           "l2": "an L2 flush (256 MiB write) between iterations, its own time (%.3f ms) subtracted" % ms_flush,
           "roofline": dict(fp32_roofline(ctx, "rx_front_kernel<127,8,exact> (all frame chunks)", samples * NTAPS, k_front, "exact"),
                            note="the step is paced by the Costas loop's per-symbol dependency chain (1,024 streams x 16,384 symbols, strictly sequential "
                                 "per stream, qpsk.c:196-207), run per frame chunk on a second stream under the next chunk's front end; "
                                 "step_frac_of_fp32 = the whole step against the same ceiling",
                            step_frac_of_fp32=samples * NTAPS / (ms * 1e-3) / ctx.fp32_peak["exact"]),
           "clocks": clocks}
    # end to end through the host entry point
    h_pcm = torch.empty((Cn, nsamp), dtype=torch.int16, pin_memory=True)
    h_pcm.copy_(pcm)
    h_out = torch.empty((Cn, F * (FRAME // sps) // 4), dtype=torch.uint8, pin_memory=True)
    L = capi.lib()

    def host_step():
        capi.check(L.qpsk_b200_rx_process_host(rx.h, ctypes.c_void_p(h_pcm.data_ptr()), F, ctypes.c_void_p(h_out.data_ptr())))
    for _ in range(3):
        host_step()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        host_step()
    dt = (time.perf_counter() - t0) / n
    rec["e2e"] = {"value": samples / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(h_pcm.numel() * 2), "d2h_bytes_per_step": int(h_out.numel()),
                  "ms": dt * 1e3}
    if want_cpu:
        rows = pcm[: len(pool.cores)].cpu().numpy()
        cb = cpu_rx_baseline(pool, rows, "1200_O2", 3.0, "the first %d of this config's channels, 256 frames each, repeated" % len(rows))
        if cb is not None:
            rec["cpu_baseline"] = cb
    rx.close()
    return rec


def bench_config3(ctx, pool, want_cpu):
    """configs[3]: long-tap RRC stress, 256 taps at 8 samples/symbol over 16,384 channels (rrc_fir API only)."""
    torch, qpsk_b200, capi = ctx.torch, ctx.qpsk_b200, ctx.capi
    Cn, T, ntaps = 16384, 65536, 256
    taps = qpsk_b200.rrc_make(ntaps, 9600.0, 1200.0, 0.35)
    x = torch.randn((Cn, T, 2), device=ctx.dev, dtype=torch.float32)
    recs = []
    for mode_name, mode in (("exact", capi.MODE_EXACT), ("fast", capi.MODE_FAST)):
        f = qpsk_b200.Fir(taps, Cn, mode=mode, device=ctx.local)
        ms, iters, clocks = timed_loop(ctx, lambda: f.filter_device(x.data_ptr(), T, ctx.stream), min_seconds=0.2, min_iters=4)
        f.filter_device(x.data_ptr(), T, ctx.stream)
        torch.cuda.synchronize()
        km = f.kernel_ms()
        samples = Cn * T
        rl = fp32_roofline(ctx, "fir_kernel<256,%s>" % mode_name, samples * ntaps, km, mode_name)
        rl["hbm"] = {"achieved": samples * 16 / (km * 1e-3) / 1e9, "peak": ctx.hbm_peak, "unit": "GB/s", "frac": samples * 16 / (km * 1e-3) / 1e9 / ctx.hbm_peak,
                     "bytes_per_sample": 16}
        recs.append({"config": "configs[3]", "workload": "rrc_fir, 256 taps (rs 1200, sps 8, alpha .35), 16,384 channels x 65,536 complex samples filtered in place "
                                                        "(8 GiB, > L2), %s arithmetic" % mode_name,
                     "ms": ms, "iters": iters, "value": samples / ms / 1e3, "unit": "Msamples/s", "roofline": rl, "clocks": clocks,
                     "l2": "8 GiB working set exceeds L2"})
        f.close()
        x.normal_()                                          # in-place filtering 20 times over grows the data: draw fresh samples
    del x
    # end to end: page-locked host samples in place (2 GiB each way per call)
    Th = 16384
    h = torch.randn((Cn, Th, 2), dtype=torch.float32).pin_memory()
    f = qpsk_b200.Fir(taps, Cn, device=ctx.local)
    L = capi.lib()
    for _ in range(2):
        capi.check(L.qpsk_b200_fir_process_host(f.h, ctypes.c_void_p(h.data_ptr()), Th))
    t0 = time.perf_counter()
    for _ in range(3):
        capi.check(L.qpsk_b200_fir_process_host(f.h, ctypes.c_void_p(h.data_ptr()), Th))
    dt = (time.perf_counter() - t0) / 3
    recs[0]["e2e"] = {"value": Cn * Th / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": Cn * Th * 8, "d2h_bytes_per_step": Cn * Th * 8, "ms": dt * 1e3,
                      "note": "qpsk_b200_fir_process_host on 16,384 x 16,384 samples, 8 B per sample each way: bound by the PCIe link"}
    f.close()
    del h
    if want_cpu:
        from oracle import Ref
        if Ref.available("fir256"):
            rate = pool.rate("fir", [("fir256", 65536, 7 + i) for i in range(len(pool.cores))], 3.0) / 1e6
            recs[0]["cpu_baseline"] = {"value": rate, "unit": "Msamples/s", "cores": len(pool.cores), "kind": "reference",
                                       "sample": "unmodified reference rrc_fir (rrc_fir.c:17-30) built with NTAPS=256 (oracle/_ref/libref_fir256.so, gcc -O2), one "
                                                 "process pinned to each core, each ~3 s over 65,536-sample blocks of one channel"}
    return recs


def bench_config4(ctx, pool, want_cpu):
    """configs[4]: FFT timing/frequency estimator sweep, 256..8192 points + argmax, 1 M bursts over 8 GPUs
    (131,072 bursts per GPU, sharded by qpsk_b200.shard.partition; one all-reduce of the counters)."""
    torch, qpsk_b200, np = ctx.torch, ctx.qpsk_b200, ctx.np
    total = FFT_BURSTS_PER_GPU * ctx.world
    start, nb = qpsk_b200.shard.partition(total, ctx.world, ctx.rank)
    recs = []
    bins = torch.empty(nb, dtype=torch.int32, device=ctx.dev)
    mag = torch.empty(nb, dtype=torch.float32, device=ctx.dev)
    for n in FFT_SIZES:
        g = torch.Generator(device=ctx.dev)
        g.manual_seed(1000 + n)
        tone_all = torch.randint(0, n, (total,), generator=g, device=ctx.dev)      # every rank draws the same list and takes its block
        tone = tone_all[start:start + nb]
        x = torch.empty((nb, n, 2), device=ctx.dev, dtype=torch.float32)
        t = torch.arange(n, device=ctx.dev, dtype=torch.float32)
        chunk = max(1, (1 << 24) // n)
        for b0 in range(0, nb, chunk):                      # tone at a random bin + AWGN (SNR 0 dB per sample), in chunks
            b1 = min(nb, b0 + chunk)
            ph = (tone[b0:b1, None].float() * t[None, :] % n) * (2.0 * np.pi / n)
            x[b0:b1, :, 0] = torch.cos(ph)
            x[b0:b1, :, 1] = torch.sin(ph)
            x[b0:b1] += torch.randn((b1 - b0, n, 2), device=ctx.dev, dtype=torch.float32) * 0.7071
            del ph
        f = qpsk_b200.Fft(n, device=ctx.local)
        ms, iters, clocks = timed_loop(ctx, lambda: f.argmax_device(x.data_ptr(), nb, bins.data_ptr(), mag.data_ptr(), ctx.stream), min_seconds=0.06, min_iters=10)
        torch.cuda.synchronize()
        km = f.kernel_ms()
        ok = int((bins.long() == tone).sum().item())
        chk = int(bins.long().sum().item())
        tot = qpsk_b200.shard.reduce_stats([nb, ok, chk], device=ctx.dev)
        ms_max = qpsk_b200.shard.max_over_ranks(ms, device=ctx.dev)
        if ctx.rank == 0:
            by = nb * (8 * n + 8)
            traffic, traffic_src = profiled_traffic("fft_kernel", "r0*_fft%d_v*final.summary.csv" % n)      # captured at 131,072 bursts
            recs.append({"config": "configs[4]", "workload": "FFT + |X|^2 argmax, n = %d, %d bursts per GPU x %d GPU(s) (tone at a random bin + AWGN), complex float in HBM"
                                                            % (n, nb, ctx.world),
                         "n": n, "ms": ms_max, "iters": iters, "value": tot[0] * n / ms_max / 1e3, "unit": "Msamples/s", "bursts_per_s": tot[0] / (ms_max * 1e-3),
                         "bursts": int(tot[0]), "argmax_equals_tone": int(tot[1]), "argmax_checksum": int(tot[2]),
                         # the timed region holds nothing but this kernel's launches: its average launch duration is ms itself
                         "roofline": {"bound": "hbm", "kernel": "fft_kernel<%d,estimator>" % int(np.log2(n)), "achieved": by / (ms * 1e-3) / 1e9, "peak": ctx.hbm_peak,
                                      "unit": "GB/s", "frac": by / (ms * 1e-3) / 1e9 / ctx.hbm_peak, "kernel_ms": ms, "single_launch_ms": km,
                                      "bytes_per_burst": 8 * n + 8, "peak_kind": ctx.hbm_kind,
                                      "traffic": traffic * (nb / 131072.0) if traffic else None, "traffic_source": traffic_src,
                                      "gflops": 5.0 * n * np.log2(n) * nb / (ms * 1e-3) / 1e9},
                         "l2": "%.2f GiB of bursts per GPU exceeds L2" % (nb * n * 8 / 2 ** 30), "clocks": clocks})
            if want_cpu:
                from oracle import RefAlg
                if RefAlg.available():
                    rate = pool.rate("fft", [(n, 3 + i) for i in range(len(pool.cores))], 1.0)
                    recs[-1]["cpu_baseline"] = {"value": rate * n / 1e6, "unit": "Msamples/s", "bursts_per_s": rate, "cores": len(pool.cores), "kind": "reference",
                                                "sample": "unmodified reference fftn (fft.c:110-120, complex double, oracle/_ref/libref_alg_O2.so) + numpy argmax, one "
                                                          "process pinned to each core, ~1 s each"}
        f.close()
        del x, tone, tone_all
    return recs


def bench_stream(ctx):
    """SURVEY 8(f) row 3, streaming ingest: the continuous receiver over one raw s16le file per channel (the reference's
    on-disk format and read loop, qpsk.h:14, qpsk.c:339-354), 1,024 channels x 1,000 frames from the page cache."""
    import tempfile
    torch, qpsk_b200 = ctx.torch, ctx.qpsk_b200
    Cn, F = 1024, 1000
    pcm = synth_pcm_gpu(torch, qpsk_b200, Cn, F * FRAME, ctx.dev, ctx.local, seed=23).cpu().numpy()
    with tempfile.TemporaryDirectory() as d:
        paths = []
        for c in range(Cn):
            pth = os.path.join(d, "ch%04d.raw" % c)
            pcm[c].astype("<i2").tofile(pth)
            paths.append(pth)
        stats = {}
        qpsk_b200.receive_files(paths, frames_per_call=250, device=ctx.local, keep=False, stats=stats)      # warm-up: page cache, contexts
        stats = {}
        qpsk_b200.receive_files(paths, frames_per_call=250, device=ctx.local, keep=False, stats=stats)
    return {"config": "8(f)-3 streaming ingest", "workload": "qpsk_b200_stream_run: 1,024 raw s16le files (one per channel) x 1,000 frames, batches of 250 frames read by host "
                                                             "threads while the GPU works on the previous batch (submit_host / wait), dibits delivered to a sink",
            "value": stats["samples_per_s"] / 1e6, "unit": "Msamples/s", "seconds": stats["seconds"], "read_seconds": stats["read_seconds"],
            "wait_seconds": stats["wait_seconds"], "reader_threads": stats["readers"], "frames": stats["frames"],
            "note": "wall clock of the whole run, files in the page cache; read_seconds / wait_seconds are the host's time in fread and in qpsk_b200_rx_wait"}


def bench_config0(ctx):
    """configs[0]: the reference's own `qpsk` binary (1 channel, 2400 baud, +50 Hz loop-back, 2,000 frames), wall time,
    next to the same unmodified qpsk.c linked against libqpsk_b200.so instead of rrc_fir.c / costas_loop.c."""
    out = {"config": "configs[0]", "workload": "the reference's qpsk binary (Makefile:6-7, qpsk.c:289-359): single-channel 2400-baud loop-back with a +50 Hz carrier offset, wall time"}
    import tempfile
    for name in ("qpsk_stock", "qpsk_dropin"):
        exe = os.path.join(ROOT, "oracle", "_ref", name)
        if not os.path.exists(exe):
            out[name] = "not built (reference tree absent at build time)"
            continue
        best = None
        with tempfile.TemporaryDirectory() as d:
            for _ in range(2):
                t0 = time.perf_counter()
                p = subprocess.run([exe], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
                dt = time.perf_counter() - t0
                if p.returncode == 0 and (best is None or dt < best):
                    best = dt
        out[name] = {"wall_s": best, "msamples_s": (2000 * FRAME / best / 1e6) if best else None}
    out["note"] = ("one channel, one 512-sample frame per call: the drop-in pays a launch and two copies per rx_frame, so this configuration measures call latency, "
                   "not throughput; the byte-identical scatter output is checked in tests/test_dropin_gpu.py")
    return out


def host_topology(ctx):
    """This rank's GPU on the PCIe tree: bus id, NUMA node and the bridges above it (from sysfs)."""
    info = {"rank": ctx.rank, "local": ctx.local}
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = ctx.local
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            idx = int(vis.split(",")[ctx.local])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        info["pci_bus_id"] = bus
        try:
            info["pcie_gen_width"] = "gen%d x%d" % (pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h), pynvml.nvmlDeviceGetCurrPcieLinkWidth(h))
        except Exception:
            pass
        sysbus = bus.lower()
        if len(sysbus.split(":")[0]) == 8:
            sysbus = sysbus[4:]
        p = "/sys/bus/pci/devices/" + sysbus
        if os.path.exists(p + "/numa_node"):
            info["numa_node"] = int(open(p + "/numa_node").read().strip())
        real = os.path.realpath(p)
        info["upstream"] = [x for x in real.split("/") if ":" in x][:-1]       # root port and switch ports above the GPU
    except Exception as ex:
        info["error"] = str(ex)
    return info


def bind_to_gpu_numa(topo):
    """First-touch placement: run this process on the cores of the GPU's real NUMA node (sysfs, not NVML's affinity mask,
    which a container may blank) before the pinned staging buffers are allocated and touched."""
    node = topo.get("numa_node", -1)
    if node is None or node < 0:
        return None
    path = "/sys/devices/system/node/node%d/cpulist" % node
    if not os.path.exists(path):
        return None
    cores = set()
    for part in open(path).read().strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            cores.update(range(int(a), int(b) + 1))
        elif part:
            cores.add(int(part))
    try:
        allowed = sorted(cores & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


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
    ap.add_argument("--no-configs", action="store_true", help="headline only: skip the records of the other BASELINE.json configurations")
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

    ctx = Ctx()
    ctx.torch, ctx.qpsk_b200, ctx.capi, ctx.np = torch, qpsk_b200, capi, np
    ctx.rank, ctx.world, ctx.local, ctx.dev = rank, world, local, dev
    ctx.hbm_peak, ctx.hbm_kind = load_peaks()
    # host side of the end-to-end leg: keep this rank's pinned buffers on the GPU's own NUMA node
    topo = host_topology(ctx)
    numa_cores = bind_to_gpu_numa(topo) if world > 1 else None
    topo["bound_cores"] = len(numa_cores) if numa_cores else None

    mode = capi.MODE_EXACT if args.mode == "exact" else capi.MODE_FAST
    nsamp = NFRAMES * FRAME
    pcm = synth_pcm_gpu(torch, qpsk_b200, NCHAN, nsamp, dev, local, seed=97 + rank)
    # the full pipeline of configs[2]: ... -> slicer -> descramble/de-interleave/CRC16 per frame
    # plus the FFT frequency estimator as a stage of every call (configs[2] names it; the reference itself never
    # calls its fftn): 65,536 bursts of 1,024 symbols -> 4th power -> FFT -> argmax per step
    rx = qpsk_b200.Receiver(NCHAN, NFRAMES, rs=2400.0, mode=mode, device=local, decode_frames=True,
                            estimate_offset=not bool(int(os.environ.get("QPSK_BENCH_NO_ESTIMATOR", "0"))),
                            no_fuse=bool(int(os.environ.get("QPSK_BENCH_NO_FUSE", "0"))),
                            transient_symbols=TRANSIENT)
    torch.cuda.synchronize()
    # a real (non-default) stream: the C-ABI treats a NULL stream as "the context's own stream", and
    # torch.cuda.Event only sees work on the stream it is recorded on
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    ctx.stream = stream
    # the FP32 pipe's own ceiling on this GPU, measured now (the denominator of every fp32 roofline below)
    ctx.fp32_peak = {"exact": capi.probe_fp32(local, False), "fast": capi.probe_fp32(local, True)}

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
    total_ms = qpsk_b200.shard.max_over_ranks(total_ms, device=dev)
    ms_per_step = total_ms / args.steps
    samples_per_step = NCHAN * nsamp
    value = world * samples_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- strong scaling of ONE 65,536-channel set (SURVEY 8(d) C3 asks for both): every rank takes its contiguous block of the
    # channel set (qpsk_b200.shard.partition) out of the PCM it already holds; reported next to the weak-scaling headline
    strong = None
    if world > 1 and not args.strong and NCHAN % world == 0:
        sc = NCHAN // world
        rxs = qpsk_b200.Receiver(sc, NFRAMES, rs=2400.0, mode=mode, device=local, decode_frames=True, estimate_offset=True, transient_symbols=TRANSIENT)
        sub = pcm[:sc]                                   # rows are independent channels: any block is as good as the rank's own
        for _ in range(W):
            rxs.process_device(sub.data_ptr(), NFRAMES, stream)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            rxs.process_device(sub.data_ptr(), NFRAMES, stream)
        s1.record()
        barrier()
        sms = qpsk_b200.shard.max_over_ranks(s0.elapsed_time(s1), device=dev) / args.steps
        strong = {"channels_total": NCHAN, "channels_per_gpu": sc, "ms_per_step": sms, "value": NCHAN * nsamp / (sms * 1e-3) / 1e6, "unit": "Msamples/s",
                  "speedup_vs_one_gpu_weak_step": ms_per_step / sms, "ideal": world,
                  "efficiency": ms_per_step / sms / world,
                  "note": "one set of 65,536 channels split over the ranks; fewer channels per GPU are an awkward number of waves of "
                          "whole-stream CTAs (32,768 channels = 3.46 waves of 2 CTAs x 148 SMs, 16,384 = 1.73, 8,192 = 0.86), so the launch "
                          "policy cuts the frames into blocks and relays the loop state from CTA to CTA (profiles/r02_notes.md, last section)"}
        rxs.close()

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

        def host_leg(fn):
            for _ in range(2):
                capi.check(fn(rx.h, ctypes.c_void_p(h_pcm.data_ptr()), NFRAMES, ctypes.c_void_p(h_out.data_ptr())))
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                capi.check(fn(rx.h, ctypes.c_void_p(h_pcm.data_ptr()), NFRAMES, ctypes.c_void_p(h_out.data_ptr())))
            torch.cuda.synchronize()
            dt = qpsk_b200.shard.max_over_ranks(time.perf_counter() - t0, device=dev)
            return world * samples_per_step * args.steps / dt / 1e6
        # the same host<->device copies (slices, streams, events) with no kernel launched: the ingest ceiling of this run, taken
        # before AND after the end-to-end leg (on boxes whose GPUs share uplinks or host memory bandwidth the rate drifts by
        # 20 % within seconds: a 4-GPU run measured e2e 75 against a copy-only 62 taken afterwards); the ceiling is the better one
        copy_before = host_leg(L.qpsk_b200_rx_probe_copy_host)
        e2e_rate = host_leg(L.qpsk_b200_rx_process_host)
        copy_after = host_leg(L.qpsk_b200_rx_probe_copy_host)
        copy_rate = max(copy_before, copy_after)
        e2e = {"value": e2e_rate, "unit": "Msamples/s", "h2d_bytes_per_step": int(h_pcm.numel() * 2), "d2h_bytes_per_step": int(h_out.numel()),
               "copy_only": {"value": copy_rate, "unit": "Msamples/s", "gbytes_s_h2d": copy_rate * 2e6 / 1e9,
                             "before_after": [copy_before, copy_after],
                             "what": "qpsk_b200_rx_probe_copy_host: the identical copies with no kernels, all ranks at once, before and after the e2e leg (the better one)"},
               "frac_of_copy_only": e2e_rate / copy_rate}
        del h_pcm, h_out
    topos = [topo]
    if world > 1:
        topos = [None] * world
        dist.all_gather_object(topos, topo)

    # ---- the other BASELINE.json configurations
    want_cpu = (not args.no_cpu_baseline) and world == 1 and rank == 0
    pool = CpuPool() if want_cpu else None
    cpu_rows = pcm[: len(pool.cores)].cpu().numpy() if want_cpu else None
    rx.close()
    del pcm
    torch.cuda.empty_cache()
    configs = []
    if not args.no_configs:
        ctx.flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)      # 256 MiB > 126 MB L2
        try:
            # the FFT sweep first, straight after the PCIe-bound end-to-end legs: its large transforms keep the FP32 pipe ~70 % busy
            # next to 4-5 TB/s of HBM traffic, and behind a second of the 256-tap filter the board's power cap clips their clock
            # (every record carries the clocks and reasons it was measured under)
            fft_recs = bench_config4(ctx, pool, want_cpu)
            if world == 1:
                configs.append(bench_config0(ctx))
                configs.append(bench_config1(ctx, pool, want_cpu))
                configs.extend(bench_config3(ctx, pool, want_cpu))
            configs.extend(fft_recs)
            if world == 1:
                configs.append(bench_stream(ctx))
        except Exception as ex:      # the headline stands on its own
            configs.append({"error": "%s: %s" % (type(ex).__name__, ex)})

    if rank == 0:
        k_front = sum(m[0] for m in front_ms) / len(front_ms)
        k_costas = sum(m[1] for m in front_ms) / len(front_ms)
        # algorithmic bytes of the fused pipeline: 2 B PCM in + 2 bits per symbol out per sample
        alg_bytes = samples_per_step * (2.0 + 2.0 / SPS / 8.0)
        achieved = alg_bytes / (k_front * 1e-3) / 1e9
        roof = fp32_roofline(ctx, "rx_front_kernel<127,4,%s>" % args.mode, samples_per_step * NTAPS, k_front, args.mode)
        traffic, traffic_src = profiled_traffic("rx_front_kernel", "r02_rx_front_v*_%s.summary.csv" % ("transient" if TRANSIENT else "default"))
        roof["traffic"] = traffic * (NCHAN * NFRAMES / (65536.0 * 64.0)) if traffic else None
        roof["traffic_source"] = traffic_src
        roof["hbm"] = {"achieved": achieved, "peak": ctx.hbm_peak, "unit": "GB/s", "frac": achieved / ctx.hbm_peak, "bytes_per_sample": 2.0625,
                       "peak_kind": ctx.hbm_kind}
        sm_mhz = clocks.get("sm_mhz") or 1965
        roof["nominal"] = {"peak_tap_updates_per_s": (64.0 if args.mode == "fast" else 32.0) * 148 * sm_mhz * 1e6,
                           "frac": samples_per_step * NTAPS / (k_front * 1e-3) / ((64.0 if args.mode == "fast" else 32.0) * 148 * sm_mhz * 1e6),
                           "peak_kind": "128 FP32 lanes/clk/SM x 148 SM x sampled SM clock, one multiply or add per lane-cycle (exact) / one FMA (fast)"}
        line = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if (args.strong and world > 1) else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "channels_per_gpu": NCHAN, "frames_per_step": NFRAMES, "arithmetic": args.mode,
                       "stages": "mixer->FIR(127 taps)->timing->Costas->slicer->descramble/deinterleave/CRC16 + FFT(1024)/argmax frequency estimator",
                       "l2": "inputs (4 GiB PCM per step) exceed L2; no flush needed", "transient_symbols": TRANSIENT,
                       "decoded_mbit_s": value / 2.0, "parallelism": "channels sharded over %d GPU(s), no data-path collective" % world},
            "roofline": roof,
            "kernels_ms": {"rx_front": k_front, "costas": k_costas},
            "clocks": clocks, "gpu_launches": int(launches),
            "stats": {"symbols": stats[0], "frames_crc_checked": stats[1], "frames_crc_ok": stats[2], "channels_locked": stats[3],
                      "note": "random payload: CRC passes are chance (2^-16); counters show K4 ran over every frame"},
        }
        if strong is not None:
            line["strong_scaling"] = strong
        if e2e is not None:
            e2e["ranks"] = topos
            line["e2e"] = e2e
        if want_cpu:      # reported at N = 1 only (rank 0), per the measurement contract
            try:
                cb = headline_cpu_baseline(pool, cpu_rows, 8.0, "the first %d channels of this run's PCM, 64 frames each, repeated" % len(cpu_rows))
                if cb is not None:
                    line["cpu_baseline"] = cb
            except Exception as ex:  # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "error": str(ex)}
        if configs:
            line["configs"] = configs
        print(json.dumps(line))
    if pool is not None:
        pool.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
