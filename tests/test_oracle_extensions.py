"""CPU checks of the oracle's restatements of this project's extensions (parity unpinned: the reference has no
rotation resolution, channel noise or square-law timing statistic).  They pin the definitions the GPU tests compare
against to simple mathematical facts."""
import numpy as np


def test_rotation_map_is_the_constellations_quarter_turn(oracle_lib):
    o = oracle_lib.Oracle()
    pts = np.array([1, 1j, -1j, -1])                              # constellation[], qpsk.c:58-63
    for d in range(4):
        for r in range(4):
            z = pts[d] * (1j ** r)
            assert pts[o.rotate_dibits([d], r)[0]] == z
    rng = np.random.default_rng(0)
    payload = rng.integers(0, 256, 30, dtype=np.uint8)
    dib = o.frame_encode(payload, 32)
    for r in range(4):
        frame, got = o.frame_decode_rotated(o.rotate_dibits(dib, r), 32)
        assert got == r and np.array_equal(frame[:30], payload)
    frame, got = o.frame_decode_rotated(rng.integers(0, 4, 128).astype(np.uint8), 32)
    assert got == -1


def test_test_noise_is_a_pure_function_of_seed_channel_and_sample(oracle_lib):
    o = oracle_lib.Oracle()
    z = np.zeros((3, 100000), np.int16)
    a = o.awgn(z, 500.0, seed=99)
    assert np.all(np.abs(a.std(axis=1) / 500.0 - 1.0) < 0.02) and np.all(np.abs(a.mean(axis=1)) < 6.0)
    assert np.abs(np.corrcoef(a[0].astype(float), a[1].astype(float))[0, 1]) < 0.02
    # pieces, shards and repeated calls agree; another seed does not
    assert np.array_equal(o.awgn(z[:, :5000], 500.0, seed=99, first_sample=20000), a[:, 20000:25000])
    assert np.array_equal(o.awgn(z[1:], 500.0, seed=99, first_channel=1), a[1:])
    assert not np.array_equal(o.awgn(z, 500.0, seed=100), a)
    # saturation like the reference's (int16_t) conversion would need: clamp, never wrap
    hot = np.full((1, 1000), 32767, np.int16)
    assert o.awgn(hot, 3000.0, seed=1).max() == 32767 and o.awgn(-hot - 1, 3000.0, seed=1).min() == -32768


def test_timing_statistic_points_at_the_pulse_peak(oracle_lib):
    """A train of raised-cosine-like pulses peaking at sample k0 mod sps: -arg(S) sps / (2 pi) = k0."""
    for rs, sps in ((2400.0, 4), (1200.0, 8)):
        o = oracle_lib.Oracle(rs=rs)
        n = np.arange(512)
        for k0 in range(sps):
            env = 1.0 + 0.8 * np.cos(2 * np.pi * (n - k0) / sps)          # |y|^2 with a line at the symbol rate
            y = (np.sqrt(env / 2) * (1 + 1j)).astype(np.complex64)
            S = complex(o.timing_sum(y[None, :])[0])
            tau = (-np.angle(S) * sps / (2 * np.pi)) % sps
            assert abs((tau - k0 + sps / 2) % sps - sps / 2) < 1e-3, (rs, k0, tau)
