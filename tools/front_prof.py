"""Timeline of the front-end kernel from an instrumented build (tools/build_variant.sh prof -DQPSK_FRONT_PROF):
QPSK_B200_LIB=$PWD/tools/bin/libq_prof.so python tools/front_prof.py > gpurun_out/front_prof.txt"""
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
t0 = a[:, 0].min()
entry = (a[:, 0] - t0) / 1e6
ends = (a[:, 2:6] - t0) / 1e6
last = ends.max(axis=1)
print("kernel span ms %.3f" % last.max())
print("role end - entry (ms): fir0 %.3f fir7 %.3f timing %.3f costas %.3f ; cta life %.3f" % (*(ends - entry[:, None]).mean(axis=0), (last - entry).mean()))
print("costas end - fir end: mean %.3f max %.3f" % ((ends[:, 3] - ends[:, :2].max(axis=1)).mean(), (ends[:, 3] - ends[:, :2].max(axis=1)).max()))
tiles = NF * 4
print("fir loop cycles per tile (w0): %.0f" % (a[:, 6] / tiles).mean())
for name, off in (("strip", 8), ("post", 16), ("fill", 24), ("empty", 32)):
    print("%-6s per tile by warp:" % name, " ".join("%6.0f" % (a[:, off + w] / tiles).mean() for w in range(8)))
bysm = collections.defaultdict(list)
for b in range(nb):
    bysm[int(a[b, 1])].append((entry[b], last[b], b))
gaps = []; lives = []
for sm, v in bysm.items():
    v.sort()
    # two chains per SM: successor = first CTA entering after an end
    ends_sorted = sorted(x[1] for x in v)
    for e in v[2:]:
        prev = max(x for x in ends_sorted if x <= e[0] + 1e-3) if any(x <= e[0] + 1e-3 for x in ends_sorted) else None
        if prev is not None: gaps.append(e[0] - prev)
    lives += [x[1] - x[0] for x in v]
print("CTAs per SM", collections.Counter(len(v) for v in bysm.values()))
print("gap between a CTA's end and the next entry on its SM (ms): mean %.4f max %.4f" % (np.mean(gaps), np.max(gaps)))
fin = np.array([max(x[1] for x in v) for v in bysm.values()])
print("SM finish (ms): min %.2f mean %.2f max %.2f" % (fin.min(), fin.mean(), fin.max()))
for sm in (0, 77):
    print(sm, [(round(x, 2), round(y, 2)) for x, y, _ in bysm[sm]])
period = (a[:, 40] - a[:, 7]) / 1e3
print("tile period around tile 64 (us): mean %.2f" % period[a[:, 7] > 0].mean())
# ---- per-tile trace of the CTAs on SM 0 / SM 77
trace = np.zeros((128, 256, 10, 2), dtype=np.int64); hdr = np.zeros((128, 4), dtype=np.uint64); cnt = ctypes.c_int(0)
assert lib.qpsk_b200_debug_front_trace(trace.ctypes.data_as(ctypes.c_void_p), hdr.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cnt)) == 0
np.savez_compressed("gpurun_out/front_trace.npz", trace=trace, hdr=hdr, count=cnt.value, t0=t0)
print("trace slots used", cnt.value)
