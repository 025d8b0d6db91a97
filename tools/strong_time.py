"""Strong-scaling shapes alone (65,536 channels split over 2 / 4 / 8 GPUs = 32,768 / 16,384 / 8,192 channels x 64 frames on one
GPU, device-resident, exact arithmetic): ms per call without relayed frame blocks (QPSK_B200_RELAY=0), with the cost model's choice, and with forced
block counts.  (The costas_follow_kernel sweep of r02_strong_shapes_v1-3.txt was this tool with QPSK_B200_FOLLOW / _FOLLOW_FB.)
usage: python tools/strong_time.py [channels ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, qpsk_b200
dev = torch.device("cuda", 0)
F = 64
shapes = [int(a) for a in sys.argv[1:]] or [8192, 16384, 32768, 65536]
pcm_all = bench.synth_pcm_gpu(torch, qpsk_b200, max(shapes), F * 512, dev, 0, seed=5)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
for C in shapes:
    pcm = pcm_all[:C]
    ref = None
    variants = [("base", {"QPSK_B200_RELAY": "0"}), ("relay", {})] + [("relay%d" % n, {"QPSK_B200_RELAY": str(n)}) for n in (4, 8)]
    for label, env in variants:
        for k in ("QPSK_B200_FOLLOW", "QPSK_B200_FOLLOW_FB", "QPSK_B200_RELAY"):
            os.environ.pop(k, None)
        env = dict(env)
        no_chunk = env.pop("no_chunk", None) is not None
        os.environ.update(env)
        rx = qpsk_b200.Receiver(C, F, device=0, decode_frames=True, estimate_offset=True, transient_symbols=True, no_chunk=no_chunk)
        for _ in range(3):
            rx.process_device(pcm.data_ptr(), F, st.cuda_stream)
        torch.cuda.synchronize()
        ms = []
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); rx.process_device(pcm.data_ptr(), F, st.cuda_stream); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        try:
            rx.sync()                                    # reports a watchdog exit of the loop kernel, if any
            rx.reset()
            rx.process_device(pcm.data_ptr(), F, st.cuda_stream); rx.sync()
            chk = int(rx.dibits().astype(np.int64).sum())
        except Exception as ex:
            print("C %6d %-6s FAILED: %s" % (C, label, ex), flush=True)
            chk = -1
        if ref is None: ref = chk
        print("C %6d %-6s ms median %.3f min %.3f  kernels (front, loop) %s  dibit checksum %d %s" % (
            C, label, float(np.median(ms)), min(ms), rx.kernel_ms(), chk, "OK" if chk == ref else "MISMATCH"), flush=True)
        rx.close()
