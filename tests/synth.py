"""Synthetic multi-channel QPSK PCM for the parity tests (test infrastructure).

Bits come from a seeded numpy generator; the waveform is produced by the oracle's restatement
of qpsk_packet_mod/tx_frame (qpsk.c:225-285) at a per-channel carrier CENTER + df, optionally
with AWGN at a given Es/N0 added before the int16 conversion clips.
"""
import numpy as np

from oracle import Oracle


def make_pcm(nchan, nframes, rs=2400.0, seed=1234, max_df=75.0, esn0_db=None, oracle=None):
    o = oracle or Oracle(rs=rs)
    rng = np.random.default_rng(seed)
    nsamp = nframes * o.frame_size
    pkt_syms = 256                                   # qpsk.c:329 packets of FRAME_SIZE/2 symbols
    pkt_samp = pkt_syms * o.sps
    npkt = (nsamp + pkt_samp - 1) // pkt_samp
    pcm = np.zeros((nchan, npkt * pkt_samp), np.int16)
    dfs = rng.uniform(-max_df, max_df, nchan)
    for c in range(nchan):
        tx = o.new_tx(1500.0 + dfs[c])
        bits = rng.integers(0, 2, size=(npkt, 2 * pkt_syms), dtype=np.int32)
        for k in range(npkt):
            pcm[c, k * pkt_samp:(k + 1) * pkt_samp] = o.packet_mod(tx, bits[k])
    pcm = pcm[:, :nsamp]
    if esn0_db is not None:
        # symbol energy of the real passband signal ~ mean(pcm^2) * sps; noise variance per sample follows
        p = np.mean(pcm.astype(np.float64) ** 2)
        sigma = np.sqrt(p * o.sps / (2.0 * 10.0 ** (esn0_db / 10.0)))
        noisy = pcm.astype(np.float64) + rng.normal(0.0, sigma, pcm.shape)
        pcm = np.clip(np.trunc(noisy), -32768, 32767).astype(np.int16)
    return np.ascontiguousarray(pcm), dfs
