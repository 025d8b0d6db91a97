"""Extension modes (never part of the reference-parity path): UB_PHASE decimation +
diagonal slicer + CRC-resolved rotation make a framed loop-back decodable.  Parity here is against the oracle's
restatement of the same extensions (parity unpinned: the reference has no such modes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _framed_pcm(qpsk_b200, payload, rs, carriers, seed, esn0_db=20.0):
    """payload uint8 [C, F, nbytes] -> int16 PCM [C, (F+2)*512]: frames_encode -> filler so that the two filters'
    group delay completes a whole frame -> library transmit path -> AWGN."""
    from qpsk_b200 import bits
    C, F, nbytes = payload.shape
    sps = int(9600.0 / rs)
    nsym = 512 // sps
    delay = (126 - 126 % sps) // sps                       # symbols of group delay of the two 127-tap filters
    dib = qpsk_b200.unpack_dibits(bits.frames_encode(payload))
    sym = np.concatenate([np.zeros((C, nsym - delay), np.uint8), dib, np.zeros((C, nsym + delay), np.uint8)], axis=1)
    assert sym.shape[1] == (F + 2) * nsym
    tx = qpsk_b200.Transmitter(carriers, rs=rs)
    pcm = tx.modulate(sym)
    tx.close()
    rng = np.random.default_rng(seed)
    p = np.mean(pcm.astype(np.float64) ** 2)
    sigma = np.sqrt(p * sps / (2.0 * 10.0 ** (esn0_db / 10.0)))
    return np.clip(np.trunc(pcm + rng.normal(0.0, sigma, pcm.shape)), -32768, 32767).astype(np.int16)


@pytest.mark.parametrize("rs,nbytes", [(2400.0, 32), (1200.0, 16)])
def test_framed_loopback_decodes_and_matches_oracle(oracle_lib, rs, nbytes):
    import qpsk_b200
    from qpsk_b200 import capi
    C, F = 48, 14
    rng = np.random.default_rng(int(rs))
    payload = rng.integers(0, 256, (C, F, nbytes), dtype=np.uint8)
    carriers = (1500.0 + rng.uniform(-60, 60, C)).astype(np.float32)
    pcm = _framed_pcm(qpsk_b200, payload, rs, carriers, seed=3)
    nfr = pcm.shape[1] // 512
    rx = qpsk_b200.Receiver(C, nfr, rs=rs, ub_mode=capi.UB_PHASE, slice_diagonal=True, decode_frames=True, resolve_rotation=True)
    rx.rx_frames(pcm)
    dib, idx = rx.dibits(), rx.read(capi.OUT_INDEX)
    frames, ok, rot = rx.read(capi.OUT_FRAMES), rx.read(capi.OUT_CRC_OK), rx.read(capi.OUT_ROTATION)
    rx.close()
    # bit-exact against the oracle's restatement of the same extension modes
    o = oracle_lib.Oracle(rs=rs, ub_mode=2, slice_diagonal=True)
    want = o.rx_run(pcm, want=("index", "dibit"))
    assert np.array_equal(idx, want["index"]) and np.array_equal(dib, want["dibit"])
    nsym = 512 // int(9600.0 / rs)
    wd = want["dibit"].reshape(C, nfr, nsym)
    for c in range(0, C, 5):
        for f in range(nfr):
            wf, wr = o.frame_decode_rotated(wd[c, f], nbytes)
            assert np.array_equal(frames[c, f], wf) and int(rot[c, f]) == (wr if wr >= 0 else 255)
    # the payload comes back: frame k of the transmitter is frame k+2 of the receiver (filler + one-frame loop delay)
    got = frames[:, 2:2 + F, :nbytes - 2]
    good = ok[:, 2:2 + F].astype(bool) & (got == payload[..., :nbytes - 2]).all(axis=2)
    assert good.mean() > 0.6, good.mean()                  # acquisition and cycle slips of the reference loop cost the rest
    assert (rot[:, 2:2 + F][good] < 4).all() and len(np.unique(rot[:, 2:2 + F][good])) > 1   # the ambiguity is real
    # the reference-faithful slicer cannot do this: one bit sits on a decision boundary (SURVEY finding 3)
    rx = qpsk_b200.Receiver(C, nfr, rs=rs, ub_mode=capi.UB_PHASE, decode_frames=True, resolve_rotation=True)
    rx.rx_frames(pcm)
    assert rx.read(capi.OUT_CRC_OK)[:, 2:2 + F].mean() < 0.05
    rx.close()


@pytest.mark.parametrize("rs,nbytes", [(2400.0, 32), (1200.0, 16)])
def test_closed_loop_estimators_decode_nearly_every_frame(oracle_lib, rs, nbytes):
    """SURVEY 8(f)-4 with the estimators in the loop: UB_TAU samples at round(tau) of the square-law timing estimate instead of
    the reference's amplitude histogram, PREROTATE_OFFSET seeds every channel's d_freq from the 4th-power FFT estimate before
    the Costas loop runs.  With the diagonal slicer and CRC-resolved rotation the framed loop-back then decodes > 95 % of the
    frames at Es/N0 = 20 dB (60-85 % without the two).  Bit-exact against the oracle's restatement of the same modes, seeded
    with the same frequency."""
    import qpsk_b200
    from qpsk_b200 import capi
    C, F = 48, 14
    rng = np.random.default_rng(int(rs) + 7)
    payload = rng.integers(0, 256, (C, F, nbytes), dtype=np.uint8)
    carriers = (1500.0 + rng.uniform(-60, 60, C)).astype(np.float32)
    pcm = _framed_pcm(qpsk_b200, payload, rs, carriers, seed=5)
    nfr = pcm.shape[1] // 512
    rx = qpsk_b200.Receiver(C, nfr, rs=rs, ub_mode=capi.UB_TAU, slice_diagonal=True, decode_frames=True, resolve_rotation=True,
                            prerotate_offset=True, estimate_timing=True)
    rx.rx_frames(pcm)
    dib, idx = rx.dibits(), rx.read(capi.OUT_INDEX)
    frames, ok = rx.read(capi.OUT_FRAMES), rx.read(capi.OUT_CRC_OK)
    hz, tau = rx.read(capi.OUT_OFFSET_HZ), rx.read(capi.OUT_TIMING_TAU)
    track = rx.read(capi.OUT_TRACK)
    rx.close()
    sps = int(9600.0 / rs)
    # the estimate is the carrier offset, to the estimator's resolution rs / (4 n) (plus the loop's own 45-degree ambiguity: none in frequency)
    n_est = 1024 if nfr * (512 // sps) >= 1024 else 512
    assert np.all(np.abs(hz - (carriers - 1500.0)) < 1.5 * rs / (4 * n_est) + 0.5), np.abs(hz - (carriers - 1500.0)).max()
    # the sampling phase is the rounded estimate, and it does not wander from frame to frame once the signal is there
    assert np.array_equal(idx[:, 2:], np.rint(tau[:, 2:]).astype(np.int64) % sps)
    assert (idx[:, 3:-1] == idx[:, 2:3]).mean() > 0.98
    # bit-exact against the oracle in the same modes, its loop seeded like set_frequency(TAU * offset_hz / RS)
    o = oracle_lib.Oracle(rs=rs, ub_mode=3, slice_diagonal=True)
    st = o.new_states(C)
    for c in range(C):
        st[c].freq = float(np.float32(2.0 * np.pi * np.float64(hz[c]) / np.float64(rs)))
    want = o.rx_run(pcm, states=st, want=("index", "dibit", "freq"))
    assert np.array_equal(idx, want["index"]) and np.array_equal(dib, want["dibit"])
    assert np.array_equal(track[:, :, 1], want["freq"])
    # and the payload comes back
    got = frames[:, 2:2 + F, :nbytes - 2]
    good = ok[:, 2:2 + F].astype(bool) & (got == payload[..., :nbytes - 2]).all(axis=2)
    assert good.mean() > 0.95, good.mean()


def test_bad_ub_mode_is_rejected():
    import qpsk_b200
    with pytest.raises(RuntimeError):
        qpsk_b200.Receiver(4, 2, ub_mode=4)
