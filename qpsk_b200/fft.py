"""Batched fftn()/ifftn() (reference algorithms/fft.h:46-49) plus the fused |X|^2 argmax estimator."""
import ctypes as C

import numpy as np

from . import _capi as capi


class Fft:
    def __init__(self, n, device=0):
        self.L = capi.lib()
        self.n = n
        self.h = C.c_void_p()
        capi.check(self.L.qpsk_b200_fft_create(n, device, C.byref(self.h)))

    def close(self):
        if self.h:
            self.L.qpsk_b200_fft_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def argmax(self, bursts):
        """bursts complex64 [B, n] (host) -> (bin int32 [B], |X[bin]|^2 float32 [B]); forward transform scaled by 1/n."""
        x = np.ascontiguousarray(bursts, np.complex64)
        assert x.ndim == 2 and x.shape[1] == self.n
        b = np.empty(x.shape[0], np.int32)
        m = np.empty(x.shape[0], np.float32)
        capi.check(self.L.qpsk_b200_fft_argmax_host(self.h, x.ctypes.data_as(C.c_void_p), x.shape[0],
                                                    b.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p)))
        return b, m

    def transform(self, bursts, inverse=False):
        """fftn (forward, scaled by 1/n) or ifftn (inverse, unscaled) of every row."""
        x = np.ascontiguousarray(bursts, np.complex64)
        assert x.ndim == 2 and x.shape[1] == self.n
        out = np.empty_like(x)
        capi.check(self.L.qpsk_b200_fft_transform_host(self.h, x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                                       x.shape[0], 1 if inverse else 0))
        return out

    def argmax_device(self, d_in, nbursts, d_bin, d_mag2, stream=None):
        capi.check(self.L.qpsk_b200_fft_argmax_device(self.h, C.c_void_p(d_in), nbursts, C.c_void_p(d_bin),
                                                      C.c_void_p(d_mag2) if d_mag2 else None, C.c_void_p(stream) if stream else None))

    def kernel_ms(self):
        ms = C.c_float()
        capi.check(self.L.qpsk_b200_fft_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value
