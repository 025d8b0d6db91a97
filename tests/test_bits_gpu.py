"""Bit stages on the GPU against the oracle (itself pinned to the reference's crc16 / interleave /
scramble and its one shipped golden vector), plus the frame codec against the oracle's composition
(parity unpinned: the reference never chains these stages) and round-trip properties."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_crc16_known_answers_and_random(oracle_lib, golden):
    from qpsk_b200 import bits
    o = oracle_lib.Oracle()
    g = golden["algorithms"]
    for row, n, want in zip(g["crc_in"], g["crc_len"], g["crc_out"]):
        if n:
            assert bits.crc16(row[:n].reshape(1, -1))[0] == want
    assert bits.crc16(np.frombuffer(b"123456789", np.uint8).reshape(1, -1))[0] == 0x29B1
    assert bits.crc16(np.zeros((1, 22), np.uint8))[0] == 0x9FB4
    assert bits.crc16(np.zeros((3, 0), np.uint8)).tolist() == [0xFFFF] * 3        # empty frames
    rng = np.random.default_rng(1)
    for nbytes in (1, 2, 22, 30, 255, 1000):
        f = rng.integers(0, 256, (257, nbytes), dtype=np.uint8)
        assert np.array_equal(bits.crc16(f), np.array([o.crc16(r) for r in f], np.uint16))


def test_interleave_golden_and_random(oracle_lib, golden):
    from qpsk_b200 import bits
    o = oracle_lib.Oracle()
    g = golden["algorithms"]
    # the reference's only shipped vector, interleave.c:97-103
    want = [0b10000010, 0b00100000, 0b00001000, 0b10000010, 0b00101000, 0b10001010, 0b10100010, 0b00101000]
    assert bits.interleave(g["il_debug_in"].reshape(1, -1), 0)[0].tolist() == want
    for n in g["il_sizes"]:
        buf = ((37 * np.arange(n) + 11) % 256).astype(np.uint8).reshape(1, -1)
        assert np.array_equal(bits.interleave(buf, 0)[0], g["il_fwd_%d" % n])
        assert np.array_equal(bits.interleave(buf, 1)[0], g["il_inv_%d" % n])
    rng = np.random.default_rng(2)
    for nbytes in (1, 5, 16, 22, 32, 43, 44, 64, 347, 1000):     # 347 bytes: b divides nbits, not a bijection
        f = rng.integers(0, 256, (19, nbytes), dtype=np.uint8)
        for d in (0, 1):
            assert np.array_equal(bits.interleave(f, d), np.array([o.interleave(r, d) for r in f])), (nbytes, d)
        if nbytes != 347:
            assert np.array_equal(bits.interleave(bits.interleave(f, 0), 1), f)


def test_scramble_keystream_and_round_trip(oracle_lib, golden):
    from qpsk_b200 import bits
    o = oracle_lib.Oracle()
    g = golden["algorithms"]
    z = bits.scramble(np.zeros((2, 512), np.uint8))
    assert np.array_equal(z[0], g["scr_zero"]) and np.array_equal(z[1], g["scr_zero"])
    assert "".join(map(str, z[0][:32])) == "00033321001003200300131011031203"
    assert np.array_equal(bits.scramble(g["scr_in"].reshape(1, -1))[0], g["scr_out"])
    rng = np.random.default_rng(3)
    d = rng.integers(0, 4, (50, 128), dtype=np.uint8)
    s = bits.scramble(d)
    assert np.array_equal(s, np.array([o.scramble_stream(r)[0] for r in d]))
    assert np.array_equal(bits.scramble(s), d)                                       # additive: tx then rx is the identity


@pytest.mark.parametrize("nbytes", [16, 32])
def test_frame_codec_vs_oracle_composition(oracle_lib, nbytes):
    import qpsk_b200
    from qpsk_b200 import bits
    o = oracle_lib.Oracle()
    rng = np.random.default_rng(nbytes)
    C, F = 70, 9
    payload = rng.integers(0, 256, (C, F, nbytes), dtype=np.uint8)
    packed = bits.frames_encode(payload)
    dib = qpsk_b200.unpack_dibits(packed).reshape(C, F, 4 * nbytes)
    for c in range(0, C, 9):
        for f in range(F):
            assert np.array_equal(dib[c, f], o.frame_encode(payload[c, f, :nbytes - 2], nbytes))
    frames, ok = bits.frames_decode(packed, nbytes)
    assert ok.all() and np.array_equal(frames[..., :nbytes - 2], payload[..., :nbytes - 2])
    # any corrupted dibit must fail the CRC and decode exactly like the oracle
    bad = packed.copy()
    bad[3, 5] ^= 0x10
    frames_b, ok_b = bits.frames_decode(bad, nbytes)
    f_hit = 5 // nbytes
    assert ok_b[3, f_hit] == 0 and ok_b.sum() == C * F - 1
    want, want_ok = o.frame_decode(qpsk_b200.unpack_dibits(bad[3:4])[0][f_hit * 4 * nbytes:(f_hit + 1) * 4 * nbytes], nbytes)
    assert not want_ok and np.array_equal(frames_b[3, f_hit], want)


def test_receiver_decodes_frames_on_device(oracle_lib):
    """DECODE_FRAMES: K4 runs on the slicer output in HBM; verdicts and frames equal the oracle's
    composition applied to the oracle's dibits, and the counters add up."""
    import qpsk_b200
    from qpsk_b200 import capi
    from synth import make_pcm
    o = oracle_lib.Oracle()
    pcm, _ = make_pcm(40, 6, seed=77, esn0_db=20.0, oracle=o)
    want_dibits = o.rx_run(pcm, want=("dibit",))["dibit"].reshape(40, 6, 128)
    rx = qpsk_b200.Receiver(40, 6, decode_frames=True)
    rx.rx_frames(pcm)
    frames, ok = rx.read(capi.OUT_FRAMES), rx.read(capi.OUT_CRC_OK)
    for c in range(0, 40, 7):
        for f in range(6):
            wf, wok = o.frame_decode(want_dibits[c, f], 32)
            assert np.array_equal(frames[c, f], wf) and bool(ok[c, f]) == wok
    n, passes = rx.crc_counters()
    assert n == 40 * 6 and passes == int(ok.sum())
    rx.close()


@pytest.mark.parametrize("nbytes", [16, 32])
def test_rotation_resolved_on_the_crc(oracle_lib, nbytes):
    """RESOLVE_ROTATION: every frame turned by 0..3 quarter turns (the Costas loop's ambiguity) decodes to
    the same payload and reports the turn; garbage reports 255 and the rotation-0 decode, like the oracle."""
    import qpsk_b200
    from qpsk_b200 import bits
    o = oracle_lib.Oracle()
    rng = np.random.default_rng(100 + nbytes)
    C, F = 37, 11
    payload = rng.integers(0, 256, (C, F, nbytes), dtype=np.uint8)
    dib = qpsk_b200.unpack_dibits(bits.frames_encode(payload)).reshape(C, F, 4 * nbytes)
    turns = rng.integers(0, 4, (C, F))
    ahead = np.array([[0, 1, 2, 3], [1, 3, 0, 2], [3, 2, 1, 0], [2, 0, 3, 1]], np.uint8)   # rho^r(d), qpsk.c:58-63
    turned = np.stack([[ahead[turns[c, f]][dib[c, f]] for f in range(F)] for c in range(C)]).astype(np.uint8)
    assert np.array_equal(turned[0, 0], o.rotate_dibits(dib[0, 0], turns[0, 0]))
    turned[5, 3] = rng.integers(0, 4, 4 * nbytes)                  # one frame of noise
    flat = turned.reshape(C, F * 4 * nbytes)
    packed = (flat[:, 0::4] | (flat[:, 1::4] << 2) | (flat[:, 2::4] << 4) | (flat[:, 3::4] << 6)).astype(np.uint8)
    frames, ok, rot = bits.frames_decode(packed, nbytes, resolve_rotation=True)
    for c in range(C):
        for f in range(F):
            wf, wr = o.frame_decode_rotated(turned[c, f], nbytes)
            assert int(rot[c, f]) == (wr if wr >= 0 else 255) and bool(ok[c, f]) == (wr >= 0)
            assert np.array_equal(frames[c, f], wf)
    good = np.ones((C, F), bool)
    good[5, 3] = False
    assert np.array_equal(rot[good], turns[good].astype(np.uint8)) and ok[good].all()
    assert np.array_equal(frames[good][:, :nbytes - 2], payload[good][:, :nbytes - 2])
    # without the flag only the unturned frames pass
    _, ok0 = bits.frames_decode(packed, nbytes)
    assert np.array_equal(ok0.astype(bool) & good, (turns == 0) & good)


def test_receiver_resolves_rotation_on_device(oracle_lib):
    """DECODE_FRAMES | RESOLVE_ROTATION in the receiver: OUT_ROTATION / OUT_FRAMES / OUT_CRC_OK equal the oracle's
    rotated decode of the oracle's dibits; a flagless receiver refuses OUT_ROTATION."""
    import qpsk_b200
    from qpsk_b200 import capi
    from synth import make_pcm
    o = oracle_lib.Oracle()
    pcm, _ = make_pcm(33, 5, seed=78, esn0_db=20.0, oracle=o)
    want_dibits = o.rx_run(pcm, want=("dibit",))["dibit"].reshape(33, 5, 128)
    rx = qpsk_b200.Receiver(33, 5, decode_frames=True, resolve_rotation=True)
    rx.rx_frames(pcm)
    frames, ok, rot = rx.read(capi.OUT_FRAMES), rx.read(capi.OUT_CRC_OK), rx.read(capi.OUT_ROTATION)
    for c in range(33):
        for f in range(5):
            wf, wr = o.frame_decode_rotated(want_dibits[c, f], 32)
            assert np.array_equal(frames[c, f], wf) and int(rot[c, f]) == (wr if wr >= 0 else 255) and bool(ok[c, f]) == (wr >= 0)
    rx.close()
    rx = qpsk_b200.Receiver(4, 2, decode_frames=True)
    rx.rx_frames(pcm[:4, :1024])
    with pytest.raises(RuntimeError):
        rx.read(capi.OUT_ROTATION)
    rx.close()
