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
