"""Batched transmit path against the oracle's restatement of qpsk_packet_mod/tx_frame (pinned to the
reference by tests/test_oracle_golden.py) and the reference golden PCM: bit-exact int16."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_reference_golden_pcm(golden):
    import qpsk_b200
    g = golden["rx_2400"]
    tx = qpsk_b200.Transmitter([1550.0])
    pcm = tx.modulate(qpsk_b200.bits_to_symbols(g["bits"]).reshape(1, -1))
    assert np.array_equal(pcm[0], g["pcm"][0])
    assert pcm[0, :8].tolist() == [4, -5, -3, 5, 3, 0, 5, 5] and pcm[0, 300:304].tolist() == [-8227, -7540, 10105, 13816]
    tx.close()


@pytest.mark.parametrize("rs,nchan,npackets", [(2400.0, 37, 3), (1200.0, 64, 2), (2400.0, 1, 5)])
def test_pcm_bit_exact_vs_oracle_with_streaming(oracle_lib, rs, nchan, npackets):
    import qpsk_b200
    o = oracle_lib.Oracle(rs=rs)
    rng = np.random.default_rng(int(rs) + nchan)
    carriers = (1500.0 + rng.uniform(-75, 75, nchan)).astype(np.float32)
    bits = rng.integers(0, 2, (nchan, npackets, 512), dtype=np.int32)
    want = np.zeros((nchan, npackets * 256 * o.sps), np.int16)
    for c in range(nchan):
        st = o.new_tx(float(carriers[c]))
        want[c] = np.concatenate([o.packet_mod(st, bits[c, k]) for k in range(npackets)])
    tx = qpsk_b200.Transmitter(carriers, rs=rs)
    syms = qpsk_b200.bits_to_symbols(bits.reshape(nchan, -1))
    whole = tx.modulate(syms)
    assert np.array_equal(whole, want)
    # the same stream in ragged calls: filter history, phasor and packet position carry over
    tx.reset()
    step = 128 // o.sps
    cuts = [0, step, 5 * step, 256, syms.shape[1]]
    parts = [tx.modulate(syms[:, a:b]) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    assert np.array_equal(np.concatenate(parts, axis=1), want)
    tx.close()


@pytest.mark.parametrize("rs", [2400.0, 1200.0])
def test_any_length(oracle_lib, rs):
    """tx_frame / qpsk_packet_mod take any length (qpsk.c:225-264), not only whole 128-sample tiles: calls of 1, 7, 33, 255 ...
    symbols, each its own packet (phasor renormalised at the end of every call, qpsk.c:253), history carried between them."""
    import qpsk_b200
    o = oracle_lib.Oracle(rs=rs)
    nchan, lengths = 5, [1, 7, 33, 255, 256, 40, 3, 129]
    rng = np.random.default_rng(int(rs) + 1)
    carriers = (1500.0 + rng.uniform(-75, 75, nchan)).astype(np.float32)
    bits = [rng.integers(0, 2, (nchan, 2 * n), dtype=np.int32) for n in lengths]
    tx = qpsk_b200.Transmitter(carriers, rs=rs, packet_symbols=1 << 20)
    states = [o.new_tx(float(c)) for c in carriers]
    for n, b in zip(lengths, bits):
        want = np.stack([o.packet_mod(states[c], b[c]) for c in range(nchan)])
        got = tx.modulate(qpsk_b200.bits_to_symbols(b))
        tx.end_packet()
        assert got.shape == (nchan, n * o.sps)
        assert np.array_equal(got, want), "length %d" % n
    tx.close()


def test_packet_boundary_inside_a_ragged_call(oracle_lib):
    """packets of 256 symbols fed in calls that do not divide them: the phasor is renormalised at exactly the packet's last sample"""
    import qpsk_b200
    o = oracle_lib.Oracle()
    rng = np.random.default_rng(9)
    bits = rng.integers(0, 2, (3, 4, 512), dtype=np.int32)
    carriers = np.array([1450.0, 1500.0, 1571.5], np.float32)
    want = np.zeros((3, 4 * 1024), np.int16)
    for c in range(3):
        st = o.new_tx(float(carriers[c]))
        want[c] = np.concatenate([o.packet_mod(st, bits[c, k]) for k in range(4)])
    tx = qpsk_b200.Transmitter(carriers)
    syms = qpsk_b200.bits_to_symbols(bits.reshape(3, -1))
    cuts = [0, 100, 101, 300, 555, 1024]
    got = np.concatenate([tx.modulate(syms[:, a:b]) for a, b in zip(cuts[:-1], cuts[1:])], axis=1)
    assert np.array_equal(got, want)
    tx.close()


def test_tx_then_rx_loopback_is_what_the_reference_does(oracle_lib):
    """The reference's experiment (qpsk.c:289-359) on the GPU end to end: modulate at CENTER+50 Hz,
    receive, and compare the decisions with the oracle receiving the oracle's PCM."""
    import qpsk_b200
    o = oracle_lib.Oracle()
    rng = np.random.default_rng(50)
    bits = rng.integers(0, 2, (4, 8 * 512), dtype=np.int32)
    tx = qpsk_b200.Transmitter([1550.0] * 4)
    pcm = tx.modulate(qpsk_b200.bits_to_symbols(bits))
    rx = qpsk_b200.Receiver(4, pcm.shape[1] // 512)
    got = qpsk_b200.unpack_dibits(rx.rx_frames(pcm))
    want = o.rx_run(pcm, want=("dibit", "freq"))
    assert np.array_equal(got, want["dibit"])
    hz = want["freq"][:, -1] * 2400.0 / (2 * np.pi)
    assert np.all(np.abs(hz - 50.0) < 2.0)          # the loop acquires the +50 Hz offset (f = 0.1309 rad/sym)
    tx.close(); rx.close()


def test_tx_rejects_empty_calls():
    import ctypes as C
    import qpsk_b200
    from qpsk_b200 import capi
    tx = qpsk_b200.Transmitter([1500.0])
    s = np.zeros((1, 33), np.uint8)
    pcm = np.zeros((1, 33 * 4), np.int16)
    assert capi.lib().qpsk_b200_tx_process_host(tx.h, s.ctypes.data_as(C.c_void_p), 0, pcm.ctypes.data_as(C.c_void_p)) == -1
    assert capi.lib().qpsk_b200_tx_process_host(tx.h, s.ctypes.data_as(C.c_void_p), 33, pcm.ctypes.data_as(C.c_void_p)) == 0
    tx.close()


@pytest.mark.gpu
def test_channel_noise_is_reproduced_by_the_oracle(oracle_lib):
    """Generator extension: counter-based AWGN on PCM in HBM equals the oracle's restatement sample for sample
    (any split of the stream, any shard of the channel set), and has the requested power."""
    import torch
    import qpsk_b200
    o = oracle_lib.Oracle()
    rng = np.random.default_rng(3)
    C, T = 37, 4097                                               # odd length: the last pair is half used
    pcm = rng.integers(-20000, 20000, (C, T), dtype=np.int16)
    pcm[0, :50] = 32767                                           # saturation on both sides
    pcm[1, :50] = -32768
    sigma = rng.uniform(10.0, 3000.0, C).astype(np.float32)
    d = torch.from_numpy(pcm).cuda()
    qpsk_b200.awgn_device(d.data_ptr(), C, T, sigma, seed=0x1234567890ABCDEF)
    got = d.cpu().numpy()
    want = o.awgn(pcm, sigma, 0x1234567890ABCDEF)
    assert np.array_equal(got, want)
    assert np.abs(got[0, :50]).max() == 32767 and got[1, :50].min() == -32768
    # the same stream generated in two pieces and as the upper shard of a larger channel set
    d2 = torch.from_numpy(pcm[5:, 1000:]).contiguous().cuda()
    qpsk_b200.awgn_device(d2.data_ptr(), C - 5, T - 1000, sigma[5:], seed=0x1234567890ABCDEF, first_sample=1000, first_channel=5)
    assert np.array_equal(d2.cpu().numpy(), want[5:, 1000:])
    # unit variance of the underlying noise
    z = torch.zeros((4, 1 << 20), dtype=torch.int16, device="cuda")
    qpsk_b200.awgn_device(z.data_ptr(), 4, 1 << 20, 1000.0, seed=7)
    zz = z.cpu().numpy().astype(np.float64)
    assert np.all(np.abs(zz.std(axis=1) / 1000.0 - 1.0) < 0.01) and np.all(np.abs(zz.mean(axis=1)) < 5.0)
    assert np.abs(np.corrcoef(zz[0], zz[1])[0, 1]) < 0.01


@pytest.mark.gpu
def test_carrier_steps_are_phase_continuous_and_match_the_oracle(oracle_lib):
    """set_carrier between packets (a stepped Doppler ramp) = the reference's fbb_tx_rect assignment (qpsk.c:320)
    with fbb_tx_phase carried over: bit-exact PCM against the oracle driven the same way."""
    import qpsk_b200
    o = oracle_lib.Oracle()
    rng = np.random.default_rng(4)
    C, npkt = 5, 4
    f0 = (1500.0 + rng.uniform(-50, 50, C)).astype(np.float32)
    sym = rng.integers(0, 4, (C, npkt * 256), dtype=np.uint8)
    tx = qpsk_b200.Transmitter(f0)
    got = []
    for k in range(npkt):
        tx.set_carrier((f0 + 0.5 * k).astype(np.float32))
        got.append(tx.modulate(sym[:, k * 256:(k + 1) * 256]))
    tx.close()
    got = np.concatenate(got, axis=1)
    for c in range(C):
        t = o.new_tx(float(f0[c]))
        parts = []
        for k in range(npkt):
            o.set_tx_carrier(t, float(np.float32(f0[c] + np.float32(0.5 * k))))
            s = sym[c, k * 256:(k + 1) * 256]
            bits = np.zeros(512, np.int32)
            bits[0::2] = s >> 1
            bits[1::2] = s & 1
            parts.append(o.packet_mod(t, bits))
        assert np.array_equal(got[c], np.concatenate(parts)), c
