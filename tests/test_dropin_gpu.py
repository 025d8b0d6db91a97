"""The reference's single-channel API exported by libqpsk_b200.so (include/qpsk_dropin.h).

1. qpsk_b200/dropin_demo (C, written against the reference API only) repeats the reference's
   loop-back experiment; every observable must equal the reference golden vectors bit for bit.
2. oracle/_ref/qpsk_dropin is the reference's UNMODIFIED qpsk.c linked against libqpsk_b200.so
   instead of rrc_fir.c + costas_loop.c; its scatter output must be byte-identical to
   oracle/_ref/qpsk_stock (the reference's own binary, config 0).  Both are prebuilt by
   oracle/Makefile where /root/reference exists; time() is pinned so rand() repeats."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_demo_program_against_reference_golden(golden, tmp_path):
    exe = os.path.join(ROOT, "qpsk_b200", "dropin_demo")
    assert os.path.exists(exe), "run `make -C qpsk_b200`"
    out = tmp_path / "demo.bin"
    subprocess.run([exe, str(out)], check=True, timeout=300)
    raw = out.read_bytes()
    g = golden["rx_2400"]
    pos = 0

    def take(dtype, count):
        nonlocal pos
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=pos)
        pos += a.nbytes
        return a

    pcm = take(np.int16, 8192)
    assert np.array_equal(pcm, g["pcm"][0])                                  # qpsk_packet_mod == reference TX
    for k in range(16):
        sym, pf, bits = take(np.complex64, 128), take(np.float32, 3), take(np.int32, 256)
        assert np.array_equal(sym.view(np.uint32), g["costas"][0, k * 128:(k + 1) * 128].view(np.uint32)), k
        assert pf[0] == g["phase"][0, k] and pf[1] == g["freq"][0, k]
        assert pf[2] == np.float32(np.float64(pf[1]) * 2400.0 / (2 * np.pi))
        assert np.array_equal((bits[0::2] | (bits[1::2] << 1)).astype(np.uint8), g["dibit"][0, k * 128:(k + 1) * 128])
    x, mem = take(np.complex64, 200), take(np.complex64, 127)
    h = (g["taps"].astype(np.float64) * 1.85).astype(np.float32)
    assert np.array_equal(x[:127].real, h) and not x[127:].any() and not x.imag.any()
    assert not mem.any()                                                     # 200 samples later the impulse has left the 127-entry delay line
    out8 = take(np.complex128, 8)
    want8 = golden["algorithms"]["fft_ramp8"]
    assert np.max(np.abs(out8 - want8)) <= 1e-5 * np.max(np.abs(want8))      # FP32 transform behind a complex double API
    a, b, c = take(np.complex128, 512), take(np.complex128, 512), take(np.complex128, 512)
    assert np.max(np.abs(b - np.fft.fft(a) / 512)) <= 1e-5 * np.max(np.abs(b))
    assert np.max(np.abs(c - a)) <= 2e-5 * np.max(np.abs(a))
    assert take(np.uint16, 1)[0] == 0x29B1
    assert take(np.uint8, 8).tolist() == [0b10000010, 0b00100000, 0b00001000, 0b10000010, 0b00101000, 0b10001010, 0b10100010, 0b00101000]
    assert take(np.uint8, 8).tolist() == [0xAA] * 4 + [0] * 4
    assert "".join(map(str, take(np.uint8, 32))) == "00033321001003200300131011031203"
    assert take(np.int32, 1)[0] == -1 and take(np.uint8, 1)[0] == 3
    assert np.array_equal(take(np.int16, 1024), g["pcm"][0, :1024])          # tx_frame(symbols) == qpsk_packet_mod(bits)
    assert take(np.int32, 2).tolist() == [0, 0]
    assert pos == len(raw)


def test_unmodified_reference_main_linked_against_the_library(tmp_path):
    stock = os.path.join(ROOT, "oracle", "_ref", "qpsk_stock")
    dropin = os.path.join(ROOT, "oracle", "_ref", "qpsk_dropin")
    if not (os.path.exists(stock) and os.path.exists(dropin)):
        pytest.skip("oracle/_ref binaries not prebuilt (reference tree absent at build time)")
    a = subprocess.run([stock], capture_output=True, timeout=600, cwd=tmp_path)
    b = subprocess.run([dropin], capture_output=True, timeout=1200, cwd=tmp_path)
    assert a.returncode == 0 and b.returncode == 0, b.stderr[-2000:]
    assert a.stderr.count(b"\n") == 256000                                    # 2000 frames x 128 scatter lines
    assert a.stderr == b.stderr
