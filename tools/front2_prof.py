"""Phase sums of rx_front2_kernel from an instrumented build (tools/build_variant.sh prof -DQPSK_FRONT_PROF)."""
import ctypes, os, sys, collections
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, qpsk_b200
from qpsk_b200 import capi
NCH = int(os.environ.get("PROF_CHANNELS", "65536")); NF = 64
dev = torch.device("cuda", 0)
pcm = bench.synth_pcm_gpu(torch, qpsk_b200, NCH, NF * 512, dev, 0, seed=97)
rx = qpsk_b200.Receiver(NCH, NF, rs=2400.0, device=0, decode_frames=True, estimate_offset=True, transient_symbols=True)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
for _ in range(4):
    rx.process_device(pcm.data_ptr(), NF, st.cuda_stream)
    torch.cuda.synchronize()
print("front kernel ms", rx.kernel_ms())
buf = np.zeros((4096, 48), dtype=np.uint64)
lib = capi.lib()
assert lib.qpsk_b200_debug_front_prof(buf.ctypes.data_as(ctypes.c_void_p)) == 0
nb = NCH // 32
a = buf[:nb].astype(np.int64)
items = NF * 32
print("filter warps, cycles per item (mean over CTAs), by warp:")
for name, off in (("strip", 8), ("store", 16), ("wait mixed", 24), ("wait raw", 32)):
    print("  %-10s" % name, " ".join("%7.0f" % (a[:, off + w] / np.maximum(a[:, 40 + w], 1)).mean() for w in range(8)))
print("  items     ", " ".join("%7.0f" % a[:, 40 + w].mean() for w in range(8)))
print("filter life (w0) cycles %.0f = %.0f per tile" % (a[:, 6].mean(), a[:, 6].mean() / (NF * 4)))
print("producer: life %.0f, waiting for ring space %.0f (%.1f %%), per chunk busy %.0f" % (a[:, 7].mean(), a[:, 3].mean(), 100 * a[:, 3].mean() / a[:, 7].mean(), (a[:, 7] - a[:, 3]).mean() / (items + 8)))
b = buf[2048:2048 + nb].astype(np.int64)
print("timing warp I: life %.0f, waiting for items %.0f (%.1f %%), frame ends %.0f, prefetched items %.0f of %d" % (b[:, 3].mean(), b[:, 0].mean(), 100 * b[:, 0].mean() / b[:, 3].mean(), b[:, 1].mean(), b[:, 2].mean(), items))
t0 = a[:, 0].min()
print("cta entry->filter end ms %.3f, timing end %.3f" % (((a[:, 2] - a[:, 0]) / 1e6).mean(), ((b[:, 4] - a[:, 0]) / 1e6).mean()))
# ---- per-SM timeline
entry = (a[:, 0] - t0) / 1e6; fend = (a[:, 2] - t0) / 1e6; tend = (b[:, 4] - t0) / 1e6
bysm = collections.defaultdict(list)
for i in range(nb): bysm[int(a[i, 1])].append((entry[i], max(fend[i], tend[i]), i))
fin = np.array([max(x[1] for x in v) for v in bysm.values()])
print("kernel span %.3f ms; SM finish: min %.2f mean %.2f max %.2f; CTAs per SM %s" % (fin.max(), fin.min(), fin.mean(), fin.max(), dict(collections.Counter(len(v) for v in bysm.values()))))
lives = np.array([x[1] - x[0] for v in bysm.values() for x in v])
print("CTA life (entry -> last of filter/timing end) mean %.3f ms min %.3f max %.3f" % (lives.mean(), lives.min(), lives.max()))
for sm in (0, 77):
    print(sm, [(round(x, 2), round(y, 2)) for x, y, _ in sorted(bysm[sm])])
# ---- per-item trace of the CTAs on SM 0 / 77
trace = np.zeros(128 * 256 * 10 * 2, dtype=np.int64); hdr = np.zeros((128, 4), dtype=np.uint64); cnt = ctypes.c_int(0)
assert lib.qpsk_b200_debug_front_trace(trace.ctypes.data_as(ctypes.c_void_p), hdr.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cnt)) == 0
np.savez_compressed("gpurun_out/front2_trace.npz", trace=trace[:64 * 2048 * 4].reshape(64, 2048, 4), hdr=hdr[:64], count=cnt.value)
print("trace slots used", cnt.value)
