#!/usr/bin/env python
"""A small pass over every kernel (for compute-sanitizer): receiver (fused and unfused Costas, both bauds),
FIR, FFT, bit stages, transmit path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qpsk_b200
from qpsk_b200 import capi, bits

rng = np.random.default_rng(0)
for rs, C, F in ((2400.0, 33, 3), (1200.0, 40, 2)):
    pcm = rng.integers(-8000, 8000, (C, F * 512)).astype(np.int16)
    rx = qpsk_b200.Receiver(C, F, rs=rs, keep_fir=True, keep_symbols=True, decode_frames=True)
    rx.rx_frames(pcm); rx.rx_frames(pcm)
    for w in (capi.OUT_FIR, capi.OUT_INDEX, capi.OUT_DEC, capi.OUT_SYMBOLS, capi.OUT_TRACK, capi.OUT_FRAMES, capi.OUT_CRC_OK):
        rx.read(w)
    rx.close()
taps = qpsk_b200.rrc_make(256, 9600.0, 1200.0, 0.35)
f = qpsk_b200.Fir(taps, 35)
f.filter((rng.normal(size=(35, 300)) + 1j * rng.normal(size=(35, 300))).astype(np.complex64)); f.close()
for n in (8, 256, 4096):
    ff = qpsk_b200.Fft(n)
    x = (rng.normal(size=(9, n)) + 1j * rng.normal(size=(9, n))).astype(np.complex64)
    ff.argmax(x); ff.transform(x); ff.transform(x, inverse=True); ff.close()
bits.crc16(rng.integers(0, 256, (7, 30), dtype=np.uint8)); bits.interleave(rng.integers(0, 256, (7, 22), dtype=np.uint8), 0)
bits.scramble(rng.integers(0, 4, (5, 128), dtype=np.uint8))
p = rng.integers(0, 256, (5, 3, 32), dtype=np.uint8); bits.frames_decode(bits.frames_encode(p), 32)
tx = qpsk_b200.Transmitter([1500.0, 1550.0, 1480.0]); tx.modulate(rng.integers(0, 4, (3, 256), dtype=np.uint8)); tx.close()
print("sanity pass done")
