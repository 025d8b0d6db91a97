"""The CPU restatement (oracle/qpsk_oracle.c) against the golden vectors generated from the
unmodified reference (tests/golden/make_golden.py), the one vector the reference ships
(interleave.c:97-103) and the known answers of SURVEY.md Appendix B.  CPU only."""
import numpy as np
import pytest

from conftest import bits_equal


@pytest.mark.parametrize("name,rs", [("rx_2400", 2400.0), ("rx_1200", 1200.0)])
def test_rx_pipeline_matches_reference_golden(oracle_lib, golden, name, rs):
    g = golden[name]
    o = oracle_lib.Oracle(rs=rs)
    assert bits_equal(o.taps, g["taps"])
    out = o.rx_run(g["pcm"])
    for k in ("fir", "dec", "costas", "dibit", "phase", "freq"):
        assert bits_equal(out[k], g[k]), k


def test_tx_matches_reference_golden(oracle_lib, golden):
    g = golden["rx_2400"]
    o = oracle_lib.Oracle()
    tx = o.new_tx(1550.0)
    bits = g["bits"]
    pcm = np.concatenate([o.packet_mod(tx, bits[i * 512:(i + 1) * 512]) for i in range(len(bits) // 512)])
    assert np.array_equal(pcm, g["pcm"][0])
    # SURVEY Appendix B known answers
    assert list(pcm[:8]) == [4, -5, -3, 5, 3, 0, 5, 5]
    assert list(pcm[300:304]) == [-8227, -7540, 10105, 13816]
    assert o.crc16(pcm.view(np.uint8)) == 0x8B84


def test_appendix_b_known_answers(oracle_lib, golden):
    o = oracle_lib.Oracle()
    t = o.taps
    assert np.float32(t.sum(dtype=np.float32)) == pytest.approx(1.84999957, abs=2e-7)
    assert t[0] == t[126] == np.float32(0.000275141007)
    assert t[63] == np.float32(0.506498635)
    assert np.array_equal(t, t[::-1])                      # bitwise symmetric at 127 taps
    assert o.p.loop0.alpha == np.float32(0.162623003) and o.p.loop0.beta == np.float32(0.0144503005)
    assert o.p.rx_rect.re == np.float32(0.555570245) and o.p.rx_rect.im == np.float32(-0.831469595)
    out = o.rx_run(golden["rx_2400"]["pcm"][:1])
    assert out["fir"][0, 200] == np.complex64(0.61877203 - 0.984675944j)
    assert out["costas"][0, 128 + 50] == np.complex64(-0.932297945 - 1.12734723j)
    assert out["costas"][0, 128 + 126] == np.complex64(-1.03308845 + 1.03392613j)
    assert out["costas"][0, 128 + 127] == 0                # the Makefile-layout aliasing read (finding 1)
    assert out["phase"][0, 1] == np.float32(7.45614452e-05) and out["freq"][0, 1] == np.float32(0.130905882)
    assert out["phase"][0, 0] == 0 and out["freq"][0, 0] == 0
    t12 = oracle_lib.Oracle(rs=1200.0).taps
    assert t12[0] == np.float32(0.000371473026) and t12[63] == np.float32(0.253798127)


def test_fir256_matches_reference_golden(oracle_lib, golden):
    g = golden["fir256"]
    o = oracle_lib.Oracle()
    taps = o.rrc_make(256, 9600.0, 1200.0, 0.35)
    assert bits_equal(taps, g["taps"])
    assert taps[0] == np.float32(-7.23024132e-05) and taps[128] == np.float32(0.253261864)
    y = g["x"].copy()
    mem = np.zeros(256, np.complex64)
    o.fir(taps, mem, y)
    assert bits_equal(y, g["y"]) and bits_equal(mem, g["mem"])


def test_bit_stages_match_reference_golden(oracle_lib, golden):
    g = golden["algorithms"]
    o = oracle_lib.Oracle()
    for row, n, want in zip(g["crc_in"], g["crc_len"], g["crc_out"]):
        assert o.crc16(row[:n]) == want
    assert o.crc16(np.frombuffer(b"123456789", np.uint8)) == 0x29B1
    assert o.crc16(np.zeros(0, np.uint8)) == 0xFFFF
    assert o.crc16(np.zeros(22, np.uint8)) == 0x9FB4
    for n in g["il_sizes"]:
        buf = ((37 * np.arange(n) + 11) % 256).astype(np.uint8)
        assert np.array_equal(o.interleave(buf, 0), g["il_fwd_%d" % n])
        assert np.array_equal(o.interleave(buf, 1), g["il_inv_%d" % n])
        assert np.array_equal(o.interleave(o.interleave(buf, 0), 1), buf)
    # the reference's only shipped golden vector, interleave.c:97-103 (printed MSB first)
    want = [0b10000010, 0b00100000, 0b00001000, 0b10000010, 0b00101000, 0b10001010, 0b10100010, 0b00101000]
    assert list(o.interleave(g["il_debug_in"], 0)) == want == list(g["il_debug_out"])
    z, _ = o.scramble_stream(np.zeros(512, np.uint8))
    assert np.array_equal(z, g["scr_zero"])
    assert "".join(map(str, z[:32])) == "00033321001003200300131011031203"
    s, _ = o.scramble_stream(g["scr_in"])
    assert np.array_equal(s, g["scr_out"])
    back, _ = o.scramble_stream(s)
    assert np.array_equal(back, g["scr_in"])               # additive: tx -> rx is the identity


def test_fft_matches_reference_golden(oracle_lib, golden):
    g = golden["algorithms"]
    o = oracle_lib.Oracle()
    for n in (8, 256, 512):
        x = g["fft_in_%d" % n]
        assert bits_equal(o.fftn(x), g["fft_out_%d" % n])
        assert bits_equal(o.fftn(x, inverse=True), g["ifft_out_%d" % n])
        assert np.max(np.abs(o.fftn(o.fftn(x), inverse=True) - x)) < 1e-14
    r = o.fftn(np.arange(1, 9).astype(np.complex128))
    assert bits_equal(r, g["fft_ramp8"])
    assert abs(r[0] - 4.5) < 1e-15 and abs(r[1] - (-0.5 + 1.2071067811865475j)) < 1e-12
    k, m = o.fft_argmax(r)
    assert k == 0 and m == pytest.approx(20.25)
