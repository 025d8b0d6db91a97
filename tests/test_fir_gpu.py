"""Channel-batched rrc_fir()/rrc_make() against the oracle and the reference golden vector:
exact mode is 0 ULP on outputs and on the delay line; fast mode within 1e-5 (max-norm relative)."""
import numpy as np
import pytest

from conftest import bits_equal


def test_rrc_make_host_matches_oracle(oracle_lib):
    import qpsk_b200
    o = oracle_lib.Oracle()
    for ntaps, fs, rs, alpha in [(127, 9600.0, 2400.0, 0.35), (127, 9600.0, 1200.0, 0.35), (256, 9600.0, 1200.0, 0.35),
                                 (127, 9600.0, 2400.0, 0.5), (127, 9600.0, 2400.0, 1.0), (63, 8000.0, 1000.0, 0.25)]:
        assert bits_equal(qpsk_b200.rrc_make(ntaps, fs, rs, alpha), o.rrc_make(ntaps, fs, rs, alpha)), (ntaps, fs, rs, alpha)


@pytest.mark.gpu
def test_fir256_reference_golden(golden):
    import qpsk_b200
    g = golden["fir256"]
    f = qpsk_b200.Fir(g["taps"], 1)
    y = f.filter(g["x"].copy().reshape(1, -1))
    assert bits_equal(y[0], g["y"])
    assert bits_equal(f.memory[0], g["mem"])
    f.close()


@pytest.mark.gpu
@pytest.mark.parametrize("ntaps,rs,nchan,lengths", [
    (127, 2400.0, 70, (512, 1, 127, 300, 1024)),      # ragged call lengths incl. shorter than the filter
    (256, 1200.0, 33, (700, 5, 256, 129)),            # long-tap profile (config 3), two halo tiles
    (127, 2400.0, 1, (1024,)),                        # the reference's own TX call shape
    (127, 2400.0, 40, (4099, 2000, 6144)),            # long calls are cut into time blocks (saved halos); odd length = 8-byte path
    (256, 1200.0, 20, (3001, 4096)),
])
def test_fir_exact_vs_oracle_with_carried_delay_line(oracle_lib, ntaps, rs, nchan, lengths):
    import qpsk_b200
    o = oracle_lib.Oracle()
    taps = qpsk_b200.rrc_make(ntaps, 9600.0, rs, 0.35)
    rng = np.random.default_rng(ntaps + nchan)
    f = qpsk_b200.Fir(taps, nchan)
    mem = np.zeros((nchan, ntaps), np.complex64)
    for n in lengths:
        x = (rng.normal(size=(nchan, n)) + 1j * rng.normal(size=(nchan, n))).astype(np.complex64)
        want = x.copy()
        for c in range(nchan):
            o.fir(taps, mem[c], want[c])
        got = f.filter(x.copy())
        assert bits_equal(got, want), n
        assert bits_equal(f.memory, mem), n
    # set_memory / reset behave like writing the reference's memory[] array
    f.reset()
    assert not f.memory.any()
    f.memory = mem
    assert bits_equal(f.memory, mem)
    f.close()


@pytest.mark.gpu
@pytest.mark.parametrize("ntaps,rs", [(127, 2400.0), (256, 1200.0)])
def test_fir_subnormal_products_bit_exact(oracle_lib, ntaps, rs):
    """rrc_fir is a general filter (rrc_fir.c:24-26 multiplies and adds whatever it is given, nothing is flushed): inputs so
    small that tap products, partial sums or the inputs themselves are subnormal must still come out bit for bit -- the
    general entry points use the non-flushing FMUL2 + FFMA2(acc, 1, p) form, not the receiver's FMUL2.FTZ + FADD2."""
    import qpsk_b200
    o = oracle_lib.Oracle()
    taps = qpsk_b200.rrc_make(ntaps, 9600.0, rs, 0.35)
    rng = np.random.default_rng(ntaps)
    nchan, n = 9, 900
    f = qpsk_b200.Fir(taps, nchan)
    mem = np.zeros((nchan, ntaps), np.complex64)
    for scale in (1e-34, 1e-37, 3e-39, 1e-42):          # products subnormal / sums subnormal / inputs subnormal / a few ulps of the smallest
        x = ((rng.normal(size=(nchan, n)) + 1j * rng.normal(size=(nchan, n))) * scale).astype(np.complex64)
        x[0, ::7] = 0                                    # signed zeros and exact cancellations ride along
        x[1, 1::5] = -x[1, 0:-1:5]
        want = x.copy()
        for c in range(nchan):
            o.fir(taps, mem[c], want[c])
        got = f.filter(x.copy())
        assert np.any(want != 0) or scale < 1e-41
        assert bits_equal(got, want), scale
        assert bits_equal(f.memory, mem), scale
    f.close()


@pytest.mark.gpu
def test_fir_impulse_and_linearity(oracle_lib):
    import qpsk_b200
    taps = qpsk_b200.rrc_make(127, 9600.0, 2400.0, 0.35)
    f = qpsk_b200.Fir(taps, 2)
    x = np.zeros((2, 256), np.complex64)
    x[0, 0] = 1.0
    x[1, 0] = 1j
    y = f.filter(x)
    h = (taps.astype(np.float64) * 1.85).astype(np.float32)       # impulse response = taps * GAIN (rrc_fir.c:28)
    assert np.array_equal(y[0, :127].real, h) and np.array_equal(y[1, :127].imag, h)
    assert y[0, 0].real == np.float32(0.000509010861) and y[0, 63].real == np.float32(0.937022448)
    assert not y[:, 127:].any()
    f.close()


@pytest.mark.gpu
def test_fir_fast_mode_tolerance(oracle_lib):
    import qpsk_b200
    from qpsk_b200 import capi
    o = oracle_lib.Oracle()
    taps = qpsk_b200.rrc_make(256, 9600.0, 1200.0, 0.35)
    rng = np.random.default_rng(2)
    x = (rng.normal(size=(40, 2000)) + 1j * rng.normal(size=(40, 2000))).astype(np.complex64)
    want = x.copy()
    for c in range(40):
        o.fir(taps, np.zeros(256, np.complex64), want[c])
    f = qpsk_b200.Fir(taps, 40, mode=capi.MODE_FAST)
    got = f.filter(x.copy())
    err = np.max(np.abs(got - want), axis=1) / np.max(np.abs(want), axis=1)
    assert err.max() <= 1e-5
    f.close()


@pytest.mark.gpu
def test_large_host_call_is_sliced_and_equals_the_device_path(oracle_lib):
    """A bank above 2 x 32 M complex samples goes through the host entry point in channel slices over three streams;
    the bytes must equal the unsliced device path, a strided subset the oracle, and the delay lines must carry on."""
    import torch
    import qpsk_b200
    o = oracle_lib.Oracle()
    taps = qpsk_b200.rrc_make(127, 9600.0, 2400.0, 0.35)
    C, T = 2100, 32768                                           # 68.8 M samples: three slices of 1,024 / 1,024 / 52 channels
    rng = np.random.default_rng(8)
    x = (rng.standard_normal((C, T), dtype=np.float32) + 1j * rng.standard_normal((C, T), dtype=np.float32)).astype(np.complex64)
    f = qpsk_b200.Fir(taps, C)
    got = f.filter(x.copy())
    mem = f.memory
    g = qpsk_b200.Fir(taps, C)
    d = torch.from_numpy(x.view(np.float32).reshape(C, T, 2)).cuda()
    g.filter_device(d.data_ptr(), T)
    torch.cuda.synchronize()
    want = d.cpu().numpy().reshape(C, T * 2).view(np.complex64)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(mem.view(np.uint32), g.memory.view(np.uint32))
    for c in (0, 1023, 1024, 2047, 2048, 2099):
        ref = x[c].copy()
        o.fir(taps, np.zeros(127, np.complex64), ref)
        assert np.array_equal(got[c].view(np.uint32), ref.view(np.uint32)), c
    # a second (short, unsliced) call continues from the carried delay lines
    y = x[:, :300].copy()
    got2 = f.filter(y.copy())
    ref = np.concatenate([x[7], y[7]])
    o.fir(taps, np.zeros(127, np.complex64), ref)
    assert np.array_equal(got2[7].view(np.uint32), ref[T:].view(np.uint32))
    f.close(); g.close()
