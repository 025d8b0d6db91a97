"""Host-side mirror of the reference receive path over many channels.

`Receiver.rx_frames(pcm)` is rx_frame() of the reference (qpsk.c:88-218) applied to every row of
`pcm` frame after frame, with the per-channel globals of qpsk.c:36-53 / costas_loop.c:13-23 held
in HBM between calls.  All compute happens in libqpsk_b200.so on a B200.
"""
import ctypes as C

import numpy as np

from . import _capi as capi


class Receiver:
    def __init__(self, nchan, max_frames, rs=2400.0, mode=capi.MODE_EXACT, ub_mode=capi.UB_ALIAS,
                 keep_fir=False, keep_symbols=False, decode_frames=False, no_fuse=False, resolve_rotation=False, slice_diagonal=False, estimate_offset=False, estimate_timing=False, device=0, loop_bw=None, center=1500.0,
                 no_chunk=False, prerotate_offset=False, transient_symbols=False):
        self.L = capi.lib()
        cfg = capi.RxConfig()
        self.L.qpsk_b200_rx_default_config(C.byref(cfg))
        cfg.rs = rs
        cfg.center = center
        cfg.mode = mode
        cfg.ub_mode = ub_mode
        cfg.flags = ((capi.KEEP_FIR if keep_fir else 0) | (capi.KEEP_SYMBOLS if keep_symbols else 0)
                     | (capi.DECODE_FRAMES if decode_frames else 0) | (capi.NO_FUSE if no_fuse else 0)
                     | (capi.RESOLVE_ROTATION if resolve_rotation else 0) | (capi.SLICE_DIAGONAL if slice_diagonal else 0)
                     | (capi.ESTIMATE_OFFSET if estimate_offset else 0) | (capi.ESTIMATE_TIMING if estimate_timing else 0)
                     | (capi.NO_CHUNK if no_chunk else 0) | (capi.PREROTATE_OFFSET if prerotate_offset else 0)
                     | (capi.TRANSIENT_SYMBOLS if transient_symbols else 0))
        cfg.device = device
        if loop_bw is not None:
            cfg.loop_bw = loop_bw
        self.cfg = cfg
        self.h = C.c_void_p()
        capi.check(self.L.qpsk_b200_rx_create(C.byref(cfg), nchan, max_frames, C.byref(self.h)))
        self.nchan, self.max_frames = nchan, max_frames
        self.frame_size = cfg.frame_size
        self.sps = int(cfg.fs / cfg.rs)
        self.nsym = self.frame_size // self.sps
        self.last_frames = 0

    def close(self):
        if self.h:
            self.L.qpsk_b200_rx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        capi.check(self.L.qpsk_b200_rx_reset(self.h))

    # -- device-resident path ---------------------------------------------------------------
    def process_device(self, d_pcm_ptr, nframes, stream=None):
        """d_pcm_ptr: device address of int16 [C][nframes*frame_size]; asynchronous."""
        capi.check(self.L.qpsk_b200_rx_process_device(self.h, C.c_void_p(d_pcm_ptr), nframes,
                                                      C.c_void_p(stream) if stream else None))
        self.last_frames = nframes

    def sync(self):
        capi.check(self.L.qpsk_b200_rx_sync(self.h))

    # -- host path (what a user of the reference calls) ---------------------------------------
    def rx_frames(self, pcm, want_dibits=True):
        """pcm: int16 ndarray [C, F*frame_size] in host memory.  Returns packed dibits uint8 [C, F*nsym/4]."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        assert pcm.ndim == 2 and pcm.shape[0] == self.nchan and pcm.shape[1] % self.frame_size == 0
        F = pcm.shape[1] // self.frame_size
        out = np.empty((self.nchan, F * self.nsym // 4), np.uint8) if want_dibits else None
        capi.check(self.L.qpsk_b200_rx_process_host(self.h, pcm.ctypes.data_as(C.c_void_p), F,
                                                    out.ctypes.data_as(C.c_void_p) if want_dibits else None))
        self.last_frames = F
        return out

    def submit(self, pcm, out):
        """Asynchronous rx_frames: enqueue one batch (pcm int16 [C, F*frame_size], out uint8 [C, F*nsym/4], both C-contiguous
        and alive until the matching wait(); page-locked buffers overlap best).  Up to two batches may be in flight."""
        assert pcm.dtype == np.int16 and pcm.flags.c_contiguous and out.dtype == np.uint8 and out.flags.c_contiguous
        F = pcm.shape[1] // self.frame_size
        capi.check(self.L.qpsk_b200_rx_submit_host(self.h, pcm.ctypes.data_as(C.c_void_p), F, out.ctypes.data_as(C.c_void_p)))
        self.last_frames = F

    def wait(self):
        """Block until the oldest submitted batch is complete."""
        capi.check(self.L.qpsk_b200_rx_wait(self.h))

    def read(self, what):
        F, Cn, S, N = self.last_frames, self.nchan, self.nsym, self.frame_size
        shapes = {
            capi.OUT_DIBITS: ((Cn, F * S // 4), np.uint8),
            capi.OUT_INDEX: ((Cn, F), np.int32),
            capi.OUT_TRACK: ((Cn, F, 2), np.float32),
            capi.OUT_DEC: ((Cn, F * S), np.complex64),
            capi.OUT_SYMBOLS: ((Cn, F * S), np.complex64),
            capi.OUT_FIR: ((Cn, F * N), np.complex64),
            capi.OUT_TAPS: ((self.cfg.ntaps,), np.float32),
            capi.OUT_FRAMES: ((Cn, F, S // 4), np.uint8),
            capi.OUT_CRC_OK: ((Cn, F), np.uint8),
            capi.OUT_ROTATION: ((Cn, F), np.uint8),
            capi.OUT_OFFSET_BIN: ((Cn,), np.int32),
            capi.OUT_OFFSET_HZ: ((Cn,), np.float32),
            capi.OUT_TIMING_SUM: ((Cn, F), np.complex64),
            capi.OUT_TIMING_TAU: ((Cn, F), np.float32),
        }
        shape, dt = shapes[what]
        out = np.empty(shape, dt)
        capi.check(self.L.qpsk_b200_rx_read(self.h, what, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def dibits(self):
        """Unpacked dibits uint8 [C, F*nsym] (bits[0] | bits[1] << 1 of qpsk_demod, qpsk.c:74-79)."""
        p = self.read(capi.OUT_DIBITS)
        return unpack_dibits(p)

    def crc_counters(self):
        a, b = C.c_ulonglong(), C.c_ulonglong()
        capi.check(self.L.qpsk_b200_rx_crc_counters(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def estimate_offset(self, log2n):
        """Extension: per-channel carrier offset (Hz) from the 4th-power spectrum of the first 2^log2n decimated
        symbols of the last call (batched FFT + argmax kernel).  Returns (offset_hz float32 [C], bin int32 [C])."""
        hz = np.empty(self.nchan, np.float32)
        bins = np.empty(self.nchan, np.int32)
        capi.check(self.L.qpsk_b200_rx_estimate_offset(self.h, log2n, hz.ctypes.data_as(C.c_void_p), bins.ctypes.data_as(C.c_void_p)))
        return hz, bins

    def kernel_ms(self):
        a, b = C.c_float(), C.c_float()
        capi.check(self.L.qpsk_b200_rx_last_kernel_ms(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_plan(self):
        """(frame chunks, frame blocks per channel group, loop mode = capi.LOOP_*) of the most recent call."""
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        capi.check(self.L.qpsk_b200_rx_last_plan(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def launch_count(self):
        return int(self.L.qpsk_b200_rx_launch_count(self.h))


def unpack_dibits(packed):
    """uint8 [..., n] with 4 dibits per byte (symbol i at bits 2*(i%4)) -> uint8 [..., 4n]."""
    p = np.asarray(packed, np.uint8)
    out = np.empty(p.shape + (4,), np.uint8)
    for k in range(4):
        out[..., k] = (p >> (2 * k)) & 3
    return out.reshape(p.shape[:-1] + (p.shape[-1] * 4,))
