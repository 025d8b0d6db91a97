"""The CPU restatement against the unmodified reference compiled into oracle/_ref, on fresh
seeded inputs (beyond the committed golden vectors).  Skipped where oracle/_ref was not built
(it is built by oracle/Makefile whenever /root/reference exists, and travels to the GPU box)."""
import ctypes as C

import numpy as np
import pytest

from conftest import bits_equal
from synth import make_pcm


def _need(oracle_lib, flavour):
    if not oracle_lib.Ref.available(flavour):
        pytest.skip("oracle/_ref/libref_%s.so not built" % flavour)
    return oracle_lib.Ref(flavour)


@pytest.mark.parametrize("flavour,rs,esn0", [("2400", 2400.0, None), ("2400", 2400.0, 12.0), ("1200", 1200.0, 20.0), ("1200", 1200.0, 6.0)])
def test_rx_against_compiled_reference(oracle_lib, flavour, rs, esn0):
    r = _need(oracle_lib, flavour)
    assert r.layout_ok()
    o = oracle_lib.Oracle(rs=rs)
    pcm, _ = make_pcm(3, 40, rs=rs, seed=int(rs) + 3, esn0_db=esn0, oracle=o)
    a, b = o.rx_run(pcm), r.rx_run(pcm)
    for k in ("fir", "dec", "costas", "dibit", "phase", "freq"):
        assert bits_equal(a[k], b[k]), k
    if rs == 2400.0:
        assert a["index"].max() >= 4      # the aliasing read is exercised


def test_o2_reference_build_differs_only_through_the_ub(oracle_lib):
    """-O2 reorders the globals (SURVEY finding 1): identical until the first out-of-frame read."""
    r0, r2 = _need(oracle_lib, "2400"), _need(oracle_lib, "2400_O2")
    o = oracle_lib.Oracle()
    pcm, _ = make_pcm(1, 8, seed=11, oracle=o)
    a, b = r0.rx_run(pcm), r2.rx_run(pcm)
    assert bits_equal(a["fir"], b["fir"])
    r1, r12 = _need(oracle_lib, "1200"), _need(oracle_lib, "1200_O2")
    o12 = oracle_lib.Oracle(rs=1200.0)
    pcm, _ = make_pcm(1, 8, rs=1200.0, seed=12, oracle=o12)
    a, b = r1.rx_run(pcm), r12.rx_run(pcm)
    for k in ("fir", "dec", "costas", "dibit", "phase", "freq"):
        assert bits_equal(a[k], b[k]), k   # no UB at 1200 baud: every build agrees


@pytest.mark.parametrize("fs,rs,alpha", [(9600.0, 2400.0, 0.35), (9600.0, 1200.0, 0.35), (9600.0, 2400.0, 0.5), (9600.0, 2400.0, 1.0),
                                         (8000.0, 1000.0, 0.25), (48000.0, 2400.0, 0.2)])
def test_rrc_make_all_branches(oracle_lib, fs, rs, alpha):
    r = _need(oracle_lib, "2400")
    r.L.rrc_make(fs, rs, alpha)
    want = r.taps
    r.L.rrc_make(9600.0, 2400.0, 0.35)
    got = oracle_lib.Oracle().rrc_make(127, fs, rs, alpha)
    assert bits_equal(got, want)


def test_costas_loop_functions(oracle_lib):
    r = _need(oracle_lib, "2400")
    L = oracle_lib.lib()
    rng = np.random.default_rng(3)
    for bw in (np.float32(2 * np.pi / 100), np.float32(2 * np.pi / 200), np.float32(0.01)):
        r.L.create_control_loop(bw, -1.0, 1.0)
        l = oracle_lib._Loop()
        L.orc_loop_create(C.byref(l), bw, -1.0, 1.0)
        assert (l.alpha, l.beta, l.damping, l.loop_bw) == (r.L.get_alpha(), r.L.get_beta(), r.L.get_damping_factor(), r.L.get_loop_bandwidth())
        for _ in range(2000):
            s = oracle_lib._CF(float(rng.normal()), float(rng.normal()))
            e = L.orc_phase_detector(s)
            assert e == r.L.ref_phase_detector(s.re, s.im)
            r.L.advance_loop(e); r.L.phase_wrap(); r.L.frequency_limit()
            L.orc_loop_advance(C.byref(l), e); L.orc_loop_phase_wrap(C.byref(l)); L.orc_loop_frequency_limit(C.byref(l))
            assert (l.phase, l.freq) == (r.L.get_phase(), r.L.get_frequency())
    assert L.orc_phase_detector(oracle_lib._CF(0.0, 0.0)) == r.L.ref_phase_detector(0.0, 0.0) == 0.0
    for f in (0.5, 3.0, -7.0):
        r.L.set_frequency(f); L.orc_loop_set_frequency(C.byref(l), f)
        assert l.freq == r.L.get_frequency()
    for p in (1.0, 20.0, -13.0):
        r.L.set_phase(p); L.orc_loop_set_phase(C.byref(l), p)
        assert l.phase == r.L.get_phase()


def test_bit_stages_and_fft_random(oracle_lib):
    if not oracle_lib.RefAlg.available():
        pytest.skip("oracle/_ref/libref_alg.so not built")
    a, o = oracle_lib.RefAlg(), oracle_lib.Oracle()
    rng = np.random.default_rng(9)
    for n in list(range(1, 70)) + [100, 255, 1000]:
        buf = rng.integers(0, 256, n, dtype=np.uint8)
        assert a.crc16(buf) == o.crc16(buf)
        for d in (0, 1):
            assert np.array_equal(a.interleave(buf, d), o.interleave(buf, d)), (n, d)
    d = rng.integers(0, 4, 5000, dtype=np.uint8)
    assert np.array_equal(a.scramble_stream(d, 0), o.scramble_stream(d)[0])
    assert np.array_equal(a.scramble_stream(d, 1), o.scramble_stream(d)[0])
    for n in (1, 2, 4, 64, 1024, 8192):
        x = (rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex128)
        assert bits_equal(a.fftn(x), o.fftn(x)), n
        assert bits_equal(a.fftn(x, inverse=True), o.fftn(x, inverse=True)), n


def test_tx_any_length_against_compiled_reference(oracle_lib):
    """qpsk_packet_mod / tx_frame (qpsk.c:225-285) accept any length; calls of 1, 7, 33, 255 ... symbols with the filter memory and
    the phasor carried between them, oracle vs the compiled reference."""
    for flavour, rs in (("2400", 2400.0), ("1200", 1200.0)):
        r = _need(oracle_lib, flavour)
        o = oracle_lib.Oracle(rs=rs)
        rng = np.random.default_rng(int(rs))
        r.tx_reset(1531.0)
        st = o.new_tx(1531.0)
        for n in (1, 7, 33, 255, 256, 40, 3, 129):
            bits = rng.integers(0, 2, 2 * n, dtype=np.int32)
            assert np.array_equal(o.packet_mod(st, bits), r.packet_mod(bits)), (flavour, n)


def test_glibc_sincos_restatement_matches_host_libm(oracle_lib):
    """oracle.orc_glibc_{sinf,cosf} restate glibc 2.39's FMA-variant sinf/cosf, which the device NCO
    follows.  Strided sweep over every float in [-7, 7] (the loop phase never leaves [-TAU, TAU]);
    the exhaustive sweep result is recorded in DESIGN.md."""
    L = oracle_lib.lib()
    L.orc_glibc_check.restype = C.c_long
    L.orc_glibc_check.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    hi = np.float32(7.0).view(np.uint32)
    bad = L.orc_glibc_check(0, int(hi), 61)
    assert bad == 0, "%d mismatches against the host libm (is this CPU without FMA?)" % bad
