"""Streaming ingest of raw s16le files (the reference's loop-back file format, qpsk.h:14): files are fed in
calls of a few frames each and the channel state carried in HBM makes the result identical to one pass."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_receive_files_equals_oracle(oracle_lib, tmp_path):
    import qpsk_b200
    from synth import make_pcm
    o = oracle_lib.Oracle()
    pcm, _ = make_pcm(3, 21, seed=17, esn0_db=18.0, oracle=o)
    paths = []
    for c in range(3):
        p = tmp_path / ("chan%d.raw" % c)
        extra = np.arange(100 + 37 * c, dtype=np.int16)                 # a ragged tail that does not fill a frame
        np.concatenate([pcm[c], extra]).astype("<i2").tofile(p)
        paths.append(str(p))
    seen = []
    got = qpsk_b200.receive_files(paths, frames_per_call=8, on_chunk=lambda f0, chunk, rx: seen.append((f0, chunk.shape[1])))
    want = o.rx_run(pcm, want=("dibit",))["dibit"]
    assert np.array_equal(got, want)
    assert seen == [(0, 8 * 128), (8, 8 * 128), (16, 5 * 128)]


def test_reference_loopback_file(oracle_lib, tmp_path):
    """The reference's own experiment as files: modulate 8 packets at CENTER + 50 Hz into a raw file (library
    transmit path), stream it back through the receiver, compare with the golden decisions."""
    import qpsk_b200
    o = oracle_lib.Oracle()
    k = np.arange(8 * 512, dtype=np.uint64)
    bits = (((k * k + k // np.uint64(3)) >> np.uint64(1)) & np.uint64(1)).astype(np.int32)
    tx = qpsk_b200.Transmitter([1550.0])
    pcm = tx.modulate(qpsk_b200.bits_to_symbols(bits).reshape(1, -1))
    tx.close()
    path = tmp_path / "spectrum-filtered.raw"
    pcm[0].astype("<i2").tofile(path)
    got = qpsk_b200.receive_files([str(path)], frames_per_call=5)
    assert np.array_equal(got, o.rx_run(pcm, want=("dibit",))["dibit"])
