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


def test_long_stream_many_channels_from_files(oracle_lib, tmp_path):
    """The continuous receiver at scale: 1,024 channels x 2,000 frames (the reference's own run length, qpsk.c:339-354) from
    one raw s16le file per channel, read ahead by host threads while the GPU works on the previous batch
    (qpsk_b200_stream_run over qpsk_b200_rx_submit_host / _wait).  The result equals one pass over the whole PCM, and the
    oracle on a few channels; the sustained rate from the page cache is printed (pytest -s) and reported by bench.py."""
    import torch
    import qpsk_b200
    C, F = 1024, 2000
    rng = np.random.default_rng(5)
    carriers = (1500.0 + rng.uniform(-75, 75, C)).astype(np.float32)
    tx = qpsk_b200.Transmitter(carriers)
    sym = torch.randint(0, 4, (C, F * 128), dtype=torch.uint8, device="cuda")
    pcm_d = torch.empty((C, F * 512), dtype=torch.int16, device="cuda")
    tx.modulate_device(sym.data_ptr(), F * 128, pcm_d.data_ptr())
    torch.cuda.synchronize()
    tx.close()
    pcm = pcm_d.cpu().numpy()
    del pcm_d, sym
    paths = []
    for c in range(C):
        p = tmp_path / ("ch%04d.raw" % c)
        pcm[c].astype("<i2").tofile(p)
        paths.append(str(p))
    stats = {}
    got = qpsk_b200.receive_files(paths, frames_per_call=250, stats=stats)
    assert stats["frames"] == F and got.shape == (C, F * 128)
    print("streamed %d channels x %d frames from files: %.2f Gsamples/s sustained (read %.2f s, wait %.2f s of %.2f s, %d readers)"
          % (C, F, stats["samples_per_s"] / 1e9, stats["read_seconds"], stats["wait_seconds"], stats["seconds"], stats["readers"]))
    rx = qpsk_b200.Receiver(C, F)
    one_pass = qpsk_b200.unpack_dibits(rx.rx_frames(pcm))
    rx.close()
    assert np.array_equal(got, one_pass)
    o = oracle_lib.Oracle()
    pick = [0, 511, 1023]
    assert np.array_equal(got[pick], o.rx_run(pcm[pick], want=("dibit",))["dibit"])
