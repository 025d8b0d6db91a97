"""Batched FFT + argmax against the oracle's fftn/ifftn (bit-identical to the reference's complex
double recursion) and the reference golden vectors.  FP32 on the GPU: max-norm relative error
<= 1e-5 (north_star); argmax bins must be identical on tone-plus-noise bursts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5   # north_star: FFT outputs within 1e-5 relative error in FP32 (max |y - y_ref| / max |y_ref| per burst)


def rel_err(got, want):
    return np.max(np.abs(got - want), axis=-1) / np.max(np.abs(want), axis=-1)


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_forward_and_inverse_vs_oracle(oracle_lib, n):
    import qpsk_b200
    o = oracle_lib.Oracle()
    rng = np.random.default_rng(n)
    nb = 37 if n <= 1024 else 5
    x = (rng.normal(size=(nb, n)) + 1j * rng.normal(size=(nb, n))).astype(np.complex64)
    f = qpsk_b200.Fft(n)
    fwd, inv = f.transform(x), f.transform(x, inverse=True)
    want_f = np.array([o.fftn(r.astype(np.complex128)) for r in x])
    want_i = np.array([o.fftn(r.astype(np.complex128), inverse=True) for r in x])
    assert rel_err(fwd, want_f).max() <= TOL
    assert rel_err(inv, want_i).max() <= TOL
    back = f.transform(fwd, inverse=True)                     # ifft(fft(x)) == x (forward carries the 1/n)
    assert rel_err(back, x).max() <= TOL
    f.close()


def test_reference_golden_vectors(golden):
    import qpsk_b200
    g = golden["algorithms"]
    for n in (8, 256, 512):
        f = qpsk_b200.Fft(n)
        x = g["fft_in_%d" % n].astype(np.complex64).reshape(1, -1)
        assert rel_err(f.transform(x), g["fft_out_%d" % n].reshape(1, -1)).max() <= TOL
        assert rel_err(f.transform(x, inverse=True), g["ifft_out_%d" % n].reshape(1, -1)).max() <= TOL
        f.close()
    f = qpsk_b200.Fft(8)
    r = f.transform(np.arange(1, 9).astype(np.complex64).reshape(1, -1))[0]
    assert abs(r[0] - 4.5) < 1e-6 and abs(r[1] - (-0.5 + 1.2071068j)) < 1e-6
    f.close()


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096, 8192])
def test_argmax_tone_bursts(oracle_lib, n):
    """Config 4 shape: each burst = a tone at a random bin + AWGN; output = (bin, |X|^2)."""
    import qpsk_b200
    o = oracle_lib.Oracle()
    rng = np.random.default_rng(n + 1)
    nb = 300 if n <= 1024 else 40
    bins = rng.integers(0, n, nb)
    t = np.arange(n)
    x = np.exp(2j * np.pi * bins[:, None] * t[None, :] / n) * rng.uniform(0.5, 2.0, (nb, 1))
    x = (x + 0.3 * (rng.normal(size=x.shape) + 1j * rng.normal(size=x.shape))).astype(np.complex64)
    f = qpsk_b200.Fft(n)
    got_bin, got_mag = f.argmax(x)
    assert np.array_equal(got_bin, bins)
    for i in range(0, nb, 7):
        k, m = o.fft_argmax(o.fftn(x[i].astype(np.complex128)))
        assert k == got_bin[i] and abs(m - got_mag[i]) <= 1e-4 * m
    f.close()


def test_argmax_first_strict_maximum_on_ties():
    import qpsk_b200
    n = 256
    f = qpsk_b200.Fft(n)
    x = np.zeros((3, n), np.complex64)
    x[0, 0] = 1.0                                  # flat spectrum: every bin equal -> bin 0
    t = np.arange(n)
    x[1] = np.exp(2j * np.pi * 5 * t / n) + np.exp(2j * np.pi * 200 * t / n)    # two equal peaks -> the lower bin
    x[2] = 0                                       # all zero -> bin 0, magnitude 0
    b, m = f.argmax(x)
    assert list(b) == [0, 5, 0] and m[2] == 0.0
    f.close()


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("nb", [1, 5, 37])
def test_argmax_exact_ties_pick_the_lowest_bin_at_every_size(n, nb):
    """An impulse at sample 0 has a perfectly flat spectrum in any arithmetic (every butterfly adds zeros), so all n bins tie
    exactly: the estimator must return bin 0 -- across the threads of a warp, the warps of a transform and the two halves of
    the split 8192-point kernel (whose even and odd bins live in different warps) -- with |X|^2 = (a / n)^2; an all-zero
    burst gives (0, 0).  Burst counts that do not fill the last pass of a CTA exercise the masked slots."""
    import qpsk_b200
    f = qpsk_b200.Fft(n)
    amp = np.float32(1.0) + np.arange(nb, dtype=np.float32)
    x = np.zeros((nb, n), np.complex64)
    x[:, 0] = amp
    x[nb // 2] = 0
    b, m = f.argmax(x)
    want = (amp / np.float32(n)) ** 2
    want[nb // 2] = 0
    assert np.array_equal(b, np.zeros(nb, np.int32))
    assert np.allclose(m, want, rtol=1e-6, atol=0)
    # a single bin raised by one ulp-sized step above a flat spectrum must win wherever it sits (every thread / warp / half)
    rng = np.random.default_rng(n + nb)
    for k in rng.integers(0, n, 6):
        t = np.arange(n)
        y = np.zeros((1, n), np.complex64)
        y[0, 0] = n
        y[0] += (0.01 * np.exp(2j * np.pi * k * t / n)).astype(np.complex64)
        bb, _ = f.argmax(y)
        assert bb[0] == k, (n, k, bb[0])
    f.close()


@pytest.mark.parametrize("n", [16384, 65536, 1 << 19])
def test_long_transforms_four_step(oracle_lib, n):
    """fftn / ifftn of the reference take any power of two (fft.c:110-136); beyond 8192 points the library runs a four-step
    decomposition over the batched kernel (qpsk_b200_fft_big_host, what the drop-in's fftn calls): <= 1e-5 against the oracle."""
    import ctypes as C
    from qpsk_b200 import capi
    o = oracle_lib.Oracle()
    rng = np.random.default_rng(n)
    x = (rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex64)
    x[n // 3] += 40.0
    L = capi.lib()
    for inverse in (0, 1):
        got = np.empty_like(x)
        capi.check(L.qpsk_b200_fft_big_host(x.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p), n, inverse, 0))
        want = o.fftn(x.astype(np.complex128), inverse=bool(inverse))
        assert np.max(np.abs(got - want)) / np.max(np.abs(want)) <= 1e-5, (n, inverse)
    assert L.qpsk_b200_fft_big_host(x.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), 8192, 0, 0) == -1
    assert L.qpsk_b200_fft_big_host(x.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), 20000, 0, 0) == -1


def test_fft_rejects_bad_length():
    import ctypes as C
    from qpsk_b200 import capi
    L = capi.lib()
    h = C.c_void_p()
    for n in (0, 1, 3, 100, 16384):
        assert L.qpsk_b200_fft_create(n, 0, C.byref(h)) == -1
