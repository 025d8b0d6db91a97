"""Regenerates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built by
oracle/Makefile from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures are small, committed, and are what travels to the GPU box: nothing in the test
suite reads /root/reference at run time.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import Ref, RefAlg  # noqa: E402


def kat_bits(npackets=8):
    """RNG-independent bit pattern bit k = ((k*k + k/3) >> 1) & 1."""
    k = np.arange(npackets * 512, dtype=np.uint64)
    return (((k * k + k // np.uint64(3)) >> np.uint64(1)) & np.uint64(1)).astype(np.int32)


def modem(flavour, carrier, npackets, seed):
    r = Ref(flavour)
    assert r.layout_ok(), "parity flavour must have the Makefile global layout"
    bits = kat_bits(npackets)
    r.tx_reset(carrier)
    pcm = np.concatenate([r.packet_mod(bits[i * 512:(i + 1) * 512]) for i in range(npackets)])
    # a second channel with noise so the timing index moves around
    rng = np.random.default_rng(seed)
    noisy = np.clip(pcm + rng.normal(0, 900, pcm.shape), -32768, 32767).astype(np.int16)
    both = np.stack([pcm, noisy])
    out = r.rx_run(both)
    return dict(pcm=both, taps=r.taps, fir=out["fir"], dec=out["dec"], costas=out["costas"], dibit=out["dibit"],
                phase=out["phase"], freq=out["freq"], bits=bits)


def main():
    g = modem("2400", 1550.0, 8, 7)
    np.savez_compressed(os.path.join(HERE, "rx_2400.npz"), **g)
    g = modem("1200", 1530.0, 4, 8)
    np.savez_compressed(os.path.join(HERE, "rx_1200.npz"), **g)

    a = RefAlg()
    rng = np.random.default_rng(5)
    alg = {}
    msgs = [b"", b"123456789", bytes([0]), bytes(22), bytes(rng.integers(0, 256, 64, dtype=np.uint8))]
    alg["crc_in"] = np.array([np.frombuffer(m.ljust(64, b"\0"), np.uint8) for m in msgs])
    alg["crc_len"] = np.array([len(m) for m in msgs], np.int32)
    alg["crc_out"] = np.array([a.crc16(np.frombuffer(m, np.uint8)) for m in msgs], np.uint16)
    sizes = [1, 2, 3, 4, 8, 22, 32, 43, 44, 64]
    alg["il_sizes"] = np.array(sizes, np.int32)
    for n in sizes:
        buf = ((37 * np.arange(n) + 11) % 256).astype(np.uint8)
        alg["il_fwd_%d" % n] = a.interleave(buf, 0)
        alg["il_inv_%d" % n] = a.interleave(buf, 1)
    alg["il_debug_in"] = np.array([0xAA, 0xAA, 0xAA, 0xAA, 0, 0, 0, 0], np.uint8)      # interleave.c:105
    alg["il_debug_out"] = a.interleave(alg["il_debug_in"], 0)
    alg["scr_zero"] = a.scramble_stream(np.zeros(512, np.uint8), 0)
    d = rng.integers(0, 4, 300, dtype=np.uint8)
    alg["scr_in"] = d
    alg["scr_out"] = a.scramble_stream(d, 0)
    for n in (8, 256, 512):
        x = (rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex128)
        alg["fft_in_%d" % n] = x
        alg["fft_out_%d" % n] = a.fftn(x)
        alg["ifft_out_%d" % n] = a.fftn(x, inverse=True)
    alg["fft_ramp8"] = a.fftn(np.arange(1, 9).astype(np.complex128))
    np.savez_compressed(os.path.join(HERE, "algorithms.npz"), **alg)

    f = Ref("fir256")
    x = (rng.normal(size=600) + 1j * rng.normal(size=600)).astype(np.complex64)
    mem = np.zeros(256, np.complex64)
    y = x.copy()
    f.fir(mem, y)
    np.savez_compressed(os.path.join(HERE, "fir256.npz"), taps=f.taps, x=x, y=y, mem=mem)
    for n in ("rx_2400", "rx_1200", "algorithms", "fir256"):
        print(n, os.path.getsize(os.path.join(HERE, n + ".npz")))


if __name__ == "__main__":
    main()
