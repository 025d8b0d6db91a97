import sys, time, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from oracle import Oracle
from synth import make_pcm
import qpsk_b200
from qpsk_b200 import capi
for rs, C, F in ((2400.0, 5, 12), (1200.0, 40, 9), (2400.0, 70, 6)):
    o = Oracle(rs=rs)
    pcm, dfs = make_pcm(C, F, rs=rs, seed=int(rs)+C, esn0_db=20.0, oracle=o)
    ref = o.rx_run(pcm)
    rx = qpsk_b200.Receiver(C, F, rs=rs, keep_fir=True, keep_symbols=True)
    packed = rx.rx_frames(pcm)
    got = {"fir": rx.read(capi.OUT_FIR), "index": rx.read(capi.OUT_INDEX), "dec": rx.read(capi.OUT_DEC),
           "costas": rx.read(capi.OUT_SYMBOLS), "dibit": qpsk_b200.unpack_dibits(packed), "track": rx.read(capi.OUT_TRACK)}
    print("rs", rs, "C", C, "F", F, "taps", np.array_equal(rx.read(capi.OUT_TAPS).view(np.uint32), o.taps.view(np.uint32)))
    for k in ("fir", "index", "dec", "costas", "dibit"):
        a, b = got[k], ref[k]
        eq = np.array_equal(a.view(np.uint8), b.view(np.uint8))
        print("  %-7s bit-exact=%s  mismatches=%d / %d" % (k, eq, int(np.sum(a != b)), a.size))
    print("  phase", np.array_equal(got["track"][..., 0].view(np.uint32), ref["phase"].view(np.uint32)),
          "freq", np.array_equal(got["track"][..., 1].view(np.uint32), ref["freq"].view(np.uint32)))
    print("  index hist", np.bincount(ref["index"].ravel(), minlength=8), "kernel ms", rx.kernel_ms())
    # streaming: same PCM in two calls must give the same result
    rx.reset()
    h = (F // 2) * 512
    p1 = rx.rx_frames(pcm[:, :h]); p2 = rx.rx_frames(pcm[:, h:])
    print("  split-call equal", np.array_equal(np.concatenate([p1, p2], axis=1), packed))
    rx.close()
