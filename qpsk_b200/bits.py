"""Batched bit stages (reference algorithms/bit-scramble.h, interleave.h, crc16.h) and the frame codec."""
import ctypes as C

import numpy as np

from . import _capi as capi


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def crc16(frames, device=0):
    """crc16() (crc16.c:11-23) of every row of uint8 [nframes, nbytes] -> uint16 [nframes]."""
    f = np.ascontiguousarray(frames, np.uint8)
    out = np.zeros(f.shape[0], np.uint16)
    capi.check(capi.lib().qpsk_b200_bits_crc16(_p(f) if f.size else None, f.shape[1], f.shape[0], _p(out), device))
    return out


def interleave(frames, direction, device=0):
    """interleave(row, nbytes, dir) (interleave.c:43-78) of every row; 0 = INTERLEAVE, 1 = DEINTERLEAVE."""
    f = np.array(frames, np.uint8, copy=True, order="C")
    capi.check(capi.lib().qpsk_b200_bits_interleave(_p(f), f.shape[1], f.shape[0], direction, device))
    return f


def scramble(dibits, device=0):
    """scramble() (bit-scramble.c:74-84) over every row of dibits, register reset to SEED per row."""
    d = np.array(dibits, np.uint8, copy=True, order="C")
    capi.check(capi.lib().qpsk_b200_bits_scramble(_p(d), d.shape[1], d.shape[0], device))
    return d


def frames_encode(payload, device=0):
    """payload uint8 [C, F, nbytes] (last two bytes ignored) -> packed dibits uint8 [C, F*nbytes]."""
    p = np.ascontiguousarray(payload, np.uint8)
    Cn, F, nb = p.shape
    out = np.zeros((Cn, F * nb), np.uint8)
    capi.check(capi.lib().qpsk_b200_frames_encode(_p(p), nb, Cn, F, _p(out), device))
    return out


def frames_decode(packed, nbytes, device=0, resolve_rotation=False):
    """packed dibits uint8 [C, F*nbytes] -> (frames uint8 [C, F, nbytes], crc_ok uint8 [C, F]); with
    resolve_rotation also the quarter turns undone per frame (uint8 [C, F], 255 = no CRC match)."""
    d = np.ascontiguousarray(packed, np.uint8)
    Cn, F = d.shape[0], d.shape[1] // nbytes
    frames = np.zeros((Cn, F, nbytes), np.uint8)
    ok = np.zeros((Cn, F), np.uint8)
    if resolve_rotation:
        rot = np.zeros((Cn, F), np.uint8)
        capi.check(capi.lib().qpsk_b200_frames_decode_rotated(_p(d), nbytes, Cn, F, _p(frames), _p(ok), _p(rot), device))
        return frames, ok, rot
    capi.check(capi.lib().qpsk_b200_frames_decode(_p(d), nbytes, Cn, F, _p(frames), _p(ok), device))
    return frames, ok
