"""BASELINE.json's full sizes, checked through size-independent properties (the oracle only has to
run a small base set):

* config 2 size: 65,536 channels.  Every channel is one of 64 base channels; all replicas of a
  base channel must produce exactly the oracle's decisions for it (channels are independent, so
  any cross-channel leakage, slicing or padding error shows up as a replica that differs).
* config 3 size: rrc_fir over 16,384 channels, 256 taps: scaling the input by a power of two
  scales every output bit-exactly (rounding commutes with exact scaling), and a strided subset
  equals the oracle.
* config 4 size: 2^20 bursts of 256 points: a pure tone at a known bin per burst must come back as
  that bin, for every burst.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_config2_size_replicated_channels(oracle_lib):
    import torch
    import qpsk_b200
    from qpsk_b200 import capi
    from synth import make_pcm
    o = oracle_lib.Oracle()
    F, NB, C = 8, 64, 65536
    base, _ = make_pcm(NB, F, seed=2024, esn0_db=18.0, oracle=o)
    want = o.rx_run(base, want=("dibit", "phase", "freq"))
    rng = np.random.default_rng(1)
    which = rng.integers(0, NB, C)
    which[:NB] = np.arange(NB)
    d_base = torch.from_numpy(base).cuda()
    d_pcm = d_base[torch.from_numpy(which).cuda()].contiguous()            # [C, F*512] on the device
    rx = qpsk_b200.Receiver(C, F, decode_frames=True)
    rx.process_device(d_pcm.data_ptr(), F)
    rx.sync()
    got = rx.read(capi.OUT_DIBITS)
    track = rx.read(capi.OUT_TRACK)
    packed_want = np.zeros((NB, F * 128 // 4), np.uint8)
    for k in range(4):
        packed_want |= (want["dibit"][:, k::4] << (2 * k)).astype(np.uint8)
    assert np.array_equal(got, packed_want[which])
    assert np.array_equal(track[..., 0], want["phase"][which]) and np.array_equal(track[..., 1], want["freq"][which])
    n, _ = rx.crc_counters()
    assert n == C * F
    # the host-buffer entry point (sliced, pipelined) gives the same bytes
    rx.reset()
    host = rx.rx_frames(d_pcm.cpu().numpy())
    assert np.array_equal(host, got)
    rx.close()


@pytest.mark.parametrize("C", [8192, 16384, 32768])
def test_strong_scaling_shapes_replicated_channels(oracle_lib, C):
    """The shapes one rank sees when 65,536 channels are split over 8 / 4 / 2 GPUs, 64 frames per call as in the bench, two
    calls back to back on the device-resident entry point.  Each shape takes a different plan (one wave of whole-stream CTAs
    with the loop fused; 1.73 waves; frame chunks with the loop of chunk k under the front end of chunk k+1 -- the plan whose
    chasing loop once occupied every SM and starved its own front end until the watchdog): every replica of a base channel
    must decide exactly like the oracle, and the call must end without a watchdog report."""
    import torch
    import qpsk_b200
    from qpsk_b200 import capi
    from synth import make_pcm
    o = oracle_lib.Oracle()
    F, NB = 64, 32
    base, _ = make_pcm(NB, 2 * F, seed=4096 + C, esn0_db=18.0, oracle=o)
    want = o.rx_run(base, want=("dibit", "phase", "freq"))
    rng = np.random.default_rng(C)
    which = rng.integers(0, NB, C)
    which[:NB] = np.arange(NB)
    d_base = torch.from_numpy(base).cuda()
    idx = torch.from_numpy(which).cuda()
    rx = qpsk_b200.Receiver(C, F, decode_frames=True, transient_symbols=True)
    for call in range(2):
        d_pcm = d_base[:, call * F * 512:(call + 1) * F * 512][idx].contiguous()
        rx.process_device(d_pcm.data_ptr(), F)
        rx.sync()                                                           # raises on a watchdog exit
        chunks, blocks, mode = rx.last_plan()
        assert (chunks, mode) == (1, capi.LOOP_RELAYED) and blocks > 1      # the launch policy relays all three shapes
        got = qpsk_b200.unpack_dibits(rx.read(capi.OUT_DIBITS))
        track = rx.read(capi.OUT_TRACK)
        w = want["dibit"][:, call * F * 128:(call + 1) * F * 128]
        assert np.array_equal(got, w[which])
        assert np.array_equal(track[..., 0], want["phase"][which, call * F:(call + 1) * F])
        assert np.array_equal(track[..., 1], want["freq"][which, call * F:(call + 1) * F])
    n, _ = rx.crc_counters()
    assert n == C * 2 * F
    rx.close()


@pytest.mark.parametrize("host_chunks", [None, "1", "3"])
def test_multi_slice_host_call_in_frame_chunks(oracle_lib, monkeypatch, host_chunks):
    """The host-buffer entry point with more channels than one slice (> 18,944) and >= 32 frames per call: the call runs as
    channel slices x frame chunks (a quarter of the frames per job by default, so that the job behind the last copy is short),
    loop fused, state carried from chunk to chunk.  Two calls (40 and 33 frames: ragged chunks), replicated base channels,
    every decision equal to the oracle's; QPSK_B200_HOST_CHUNKS=1 is whole calls per slice, 3 a ragged split."""
    import qpsk_b200
    from synth import make_pcm
    if host_chunks is not None:
        monkeypatch.setenv("QPSK_B200_HOST_CHUNKS", host_chunks)
    o = oracle_lib.Oracle()
    F1, F2, NB, C = 40, 33, 16, 19200 + 7
    base, _ = make_pcm(NB, F1 + F2, seed=77, esn0_db=18.0, oracle=o)
    want = o.rx_run(base, want=("dibit",))["dibit"]
    rng = np.random.default_rng(5)
    which = rng.integers(0, NB, C)
    which[:NB] = np.arange(NB)
    from qpsk_b200 import capi
    rx = qpsk_b200.Receiver(C, F1, decode_frames=True, transient_symbols=True, estimate_offset=True)
    got1 = qpsk_b200.unpack_dibits(rx.rx_frames(np.ascontiguousarray(base[which, :F1 * 512])))
    assert np.array_equal(got1, want[which, :F1 * 128])
    # the in-call estimator runs behind the first frame chunk here, not at the end of the call: same bins as a small receiver's
    small = qpsk_b200.Receiver(NB, F1, estimate_offset=True, transient_symbols=True)
    small.rx_frames(np.ascontiguousarray(base[:, :F1 * 512]))
    assert np.array_equal(rx.read(capi.OUT_OFFSET_BIN), small.read(capi.OUT_OFFSET_BIN)[which])
    small.close()
    got2 = qpsk_b200.unpack_dibits(rx.rx_frames(np.ascontiguousarray(base[which, F1 * 512:])))
    assert np.array_equal(got2[:, :F2 * 128], want[which, F1 * 128:])
    n, _ = rx.crc_counters()
    assert n == C * (F1 + F2)
    rx.close()


def test_config3_size_fir_scaling_property(oracle_lib):
    import torch
    import qpsk_b200
    o = oracle_lib.Oracle()
    C, T, ntaps = 16384, 4096, 256
    taps = qpsk_b200.rrc_make(ntaps, 9600.0, 1200.0, 0.35)
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    x = torch.randn((C, T, 2), generator=g, device="cuda")
    x2 = (x * 4.0).contiguous()
    keep = x[::997].cpu().numpy().copy()
    f = qpsk_b200.Fir(taps, C)
    f.filter_device(x.data_ptr(), T)
    f.reset()
    f.filter_device(x2.data_ptr(), T)
    torch.cuda.synchronize()
    assert torch.equal(x2, x * 4.0)                    # power-of-two scaling commutes with every rounding step
    y = x[::997].cpu().numpy()
    for i in range(0, keep.shape[0], 3):
        ref = np.ascontiguousarray(keep[i]).view(np.complex64).reshape(-1).copy()
        o.fir(taps, np.zeros(ntaps, np.complex64), ref)
        assert np.array_equal(y[i].view(np.complex64).reshape(-1).view(np.uint32), ref.view(np.uint32))
    f.close()


def test_config4_size_tone_bins():
    import torch
    import qpsk_b200
    n, nb = 256, 1 << 20
    g = torch.Generator(device="cuda")
    g.manual_seed(9)
    bins = torch.randint(0, n, (nb,), generator=g, device="cuda")
    t = torch.arange(n, device="cuda")
    ph = (2 * np.pi / n) * (bins[:, None] * t[None, :]).float()
    x = torch.stack([torch.cos(ph), torch.sin(ph)], dim=-1).contiguous()      # [nb, n, 2]
    out_bin = torch.empty(nb, dtype=torch.int32, device="cuda")
    out_mag = torch.empty(nb, dtype=torch.float32, device="cuda")
    f = qpsk_b200.Fft(n)
    f.argmax_device(x.data_ptr(), nb, out_bin.data_ptr(), out_mag.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(out_bin.long(), bins)
    assert torch.allclose(out_mag, torch.ones_like(out_mag), atol=1e-4)          # unit tone, forward scaled by 1/n
    f.close()


def test_config1_shape_all_channels_bit_exact(oracle_lib):
    """BASELINE config 1: 1,024 independent 1200-baud channels (10 m profile), AWGN + per-channel carrier offset.
    No undefined behaviour at 1200 baud, so every channel must agree with the oracle bit for bit (32 frames here;
    the frame-split grid and the stand-alone Costas kernel are the code path this size takes)."""
    import qpsk_b200
    from qpsk_b200 import capi
    from synth import make_pcm
    o = oracle_lib.Oracle(rs=1200.0)
    base, _ = make_pcm(128, 32, rs=1200.0, seed=1200, esn0_db=10.0, oracle=o)
    rng = np.random.default_rng(3)
    # 1,024 distinct channels: each base channel with eight different additive noise realisations
    pcm = np.repeat(base, 8, axis=0).astype(np.int32) + rng.integers(-300, 301, (1024, base.shape[1]))
    pcm = np.clip(pcm, -32768, 32767).astype(np.int16)
    want = o.rx_run(pcm, want=("index", "dibit", "phase", "freq"))
    rx = qpsk_b200.Receiver(1024, 32, rs=1200.0)
    got = qpsk_b200.unpack_dibits(rx.rx_frames(pcm))
    assert np.array_equal(got, want["dibit"])
    assert np.array_equal(rx.read(capi.OUT_INDEX), want["index"])
    track = rx.read(capi.OUT_TRACK)
    assert np.array_equal(track[..., 0], want["phase"]) and np.array_equal(track[..., 1], want["freq"])
    rx.close()


def test_config2_every_channel_bit_exact(oracle_lib, monkeypatch, capsys):
    """SURVEY 8(d) "all channels once": the bench workload itself, 65,536 distinct channels x 64 frames, against the
    oracle on all host cores (tools/full_parity.py; ~20 s on 16 cores): dibits, timing indices and loop tracks."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import full_parity
    monkeypatch.setattr(sys, "argv", ["full_parity.py", "65536", "64"])
    full_parity.main()
    assert "bit-exact on every channel" in capsys.readouterr().out
