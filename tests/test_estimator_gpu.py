"""Extension: FFT-based carrier-offset estimator (4th power of the decimated symbols -> batched FFT -> |X|^2
argmax).  Parity unpinned by construction -- the reference has no estimator -- so the checks are (i) the same
argmax bin as numpy's FFT of the 4th power of the ORACLE's decimated symbols (which the CUDA path reproduces bit
for bit) and (ii) the physical answer: the known carrier offset of each synthetic channel within one bin."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_offset_estimate_matches_numpy_and_truth(oracle_lib):
    import qpsk_b200
    from synth import make_pcm
    o = oracle_lib.Oracle()
    C, F = 48, 16
    pcm, dfs = make_pcm(C, F, seed=44, max_df=70.0, esn0_db=25.0, oracle=o)
    rx = qpsk_b200.Receiver(C, F)
    rx.rx_frames(pcm)
    log2n = 11                                             # 2,048 symbols: 0.29 Hz per bin
    hz, bins = rx.estimate_offset(log2n)
    dec = o.rx_run(pcm, want=("dec",))["dec"][:, :1 << log2n].astype(np.complex128)
    spec = np.abs(np.fft.fft(dec ** 4, axis=1)) ** 2
    want_bins = np.argmax(spec, axis=1)
    # float32 FFT against float64: identical bins unless two bins are within rounding of each other
    close = np.abs(bins - want_bins)
    assert np.count_nonzero(np.minimum(close, (1 << log2n) - close) > 1) == 0
    assert np.count_nonzero(bins == want_bins) >= C - 2
    res = 2400.0 / (4 * (1 << log2n))
    assert np.max(np.abs(hz - dfs)) <= 3 * res + 0.5       # the transmit carrier offsets, to within a few bins
    with pytest.raises(qpsk_b200.QpskB200Error):
        rx.estimate_offset(13)                             # more symbols than the last call produced
    rx.close()


def test_estimator_inside_the_process_call(oracle_lib):
    """QPSK_B200_ESTIMATE_OFFSET: the same estimator as a stage of every process call (device-resident results,
    sliced with the channels on the host-buffer path); equals the explicit call at the same burst length and does
    not disturb the receive decisions."""
    import qpsk_b200
    from qpsk_b200 import capi
    from synth import make_pcm
    o = oracle_lib.Oracle()
    C, F = 70, 9                                           # 1,152 symbols per call -> bursts of 1,024
    pcm, dfs = make_pcm(C, F, seed=45, max_df=70.0, esn0_db=25.0, oracle=o)
    rx = qpsk_b200.Receiver(C, F, estimate_offset=True)
    got = rx.rx_frames(pcm)
    bins, hz = rx.read(capi.OUT_OFFSET_BIN), rx.read(capi.OUT_OFFSET_HZ)
    hz2, bins2 = rx.estimate_offset(10)
    assert np.array_equal(bins, bins2) and np.array_equal(hz, hz2)
    assert np.max(np.abs(hz - dfs)) <= 3 * 2400.0 / 4096 + 0.5
    want = o.rx_run(pcm, want=("dibit",))["dibit"]
    assert np.array_equal(qpsk_b200.unpack_dibits(got), want)
    # a second, shorter call re-plans the burst length (2 frames = 256 symbols)
    rx.rx_frames(pcm[:, :1024])
    assert rx.read(capi.OUT_OFFSET_BIN).max() < 256
    rx.close()
    plain = qpsk_b200.Receiver(4, 2)
    plain.rx_frames(pcm[:4, :1024])
    with pytest.raises(qpsk_b200.QpskB200Error):
        plain.read(capi.OUT_OFFSET_HZ)
    plain.close()


@pytest.mark.parametrize("rs", [2400.0, 1200.0])
def test_square_law_timing_statistic(oracle_lib, rs):
    """QPSK_B200_ESTIMATE_TIMING (extension): S = sum y^2 e^{-2 pi i n / sps} per frame from the timing warps is
    bit-identical to the oracle's restatement on the oracle's filter output, its phase sits where the two 127-tap
    filters put the eye (sample 126 mod sps), and switching it on changes no decision."""
    import qpsk_b200
    from qpsk_b200 import capi
    from synth import make_pcm
    o = oracle_lib.Oracle(rs=rs)
    C, F = 40, 7
    pcm, _ = make_pcm(C, F, rs=rs, seed=46, max_df=60.0, esn0_db=25.0, oracle=o)
    want = o.rx_run(pcm, want=("fir", "dibit", "index"))
    rx = qpsk_b200.Receiver(C, F, rs=rs, estimate_timing=True)
    got = rx.rx_frames(pcm)
    S, tau = rx.read(capi.OUT_TIMING_SUM), rx.read(capi.OUT_TIMING_TAU)
    assert np.array_equal(qpsk_b200.unpack_dibits(got), want["dibit"]) and np.array_equal(rx.read(capi.OUT_INDEX), want["index"])
    rx.close()
    Sw = o.timing_sum(want["fir"].reshape(C, F, 512))
    assert np.array_equal(S.view(np.uint32), Sw.view(np.uint32))
    sps = int(9600.0 / rs)
    eye = 126 % sps
    err = (tau[:, 1:] - eye + sps / 2) % sps - sps / 2            # circular error, first frame is the filter filling up
    assert np.max(np.abs(err)) < 0.35, np.max(np.abs(err))
    plain = qpsk_b200.Receiver(4, 2, rs=rs)
    plain.rx_frames(pcm[:4, :1024])
    with pytest.raises(qpsk_b200.QpskB200Error):
        plain.read(capi.OUT_TIMING_TAU)
    plain.close()
