"""Channel-batched rrc_fir()/rrc_make() (reference rrc_fir.h:16-17) on the GPU."""
import ctypes as C

import numpy as np

from . import _capi as capi


def rrc_make(ntaps, fs, rs, alpha):
    """rrc_make(fs, rs, alpha) (rrc_fir.c:32-76) for `ntaps` taps; host-side design arithmetic."""
    taps = np.zeros(ntaps, np.float32)
    capi.check(capi.lib().qpsk_b200_rrc_make(taps.ctypes.data_as(C.c_void_p), ntaps, fs, rs, alpha))
    return taps


class Fir:
    """rrc_fir(memory, sample, length) over `nchan` channels; `memory` lives in HBM between calls."""

    def __init__(self, taps, nchan, mode=capi.MODE_EXACT, device=0):
        self.L = capi.lib()
        self.taps = np.ascontiguousarray(taps, np.float32)
        self.ntaps, self.nchan = len(self.taps), nchan
        self.h = C.c_void_p()
        capi.check(self.L.qpsk_b200_fir_create(self.taps.ctypes.data_as(C.c_void_p), self.ntaps, nchan, mode, device, C.byref(self.h)))

    def close(self):
        if self.h:
            self.L.qpsk_b200_fir_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        capi.check(self.L.qpsk_b200_fir_reset(self.h))

    def filter(self, samples):
        """samples: complex64 [C, T] host array, filtered in place (and returned)."""
        assert samples.dtype == np.complex64 and samples.flags.c_contiguous and samples.shape[0] == self.nchan
        capi.check(self.L.qpsk_b200_fir_process_host(self.h, samples.ctypes.data_as(C.c_void_p), samples.shape[1]))
        return samples

    def filter_device(self, d_ptr, nsamples, stream=None):
        capi.check(self.L.qpsk_b200_fir_process_device(self.h, C.c_void_p(d_ptr), nsamples, C.c_void_p(stream) if stream else None))

    @property
    def memory(self):
        m = np.zeros((self.nchan, self.ntaps), np.complex64)
        capi.check(self.L.qpsk_b200_fir_get_memory(self.h, m.ctypes.data_as(C.c_void_p)))
        return m

    @memory.setter
    def memory(self, m):
        m = np.ascontiguousarray(m, np.complex64)
        assert m.shape == (self.nchan, self.ntaps)
        capi.check(self.L.qpsk_b200_fir_set_memory(self.h, m.ctypes.data_as(C.c_void_p)))

    def kernel_ms(self):
        ms = C.c_float()
        capi.check(self.L.qpsk_b200_fir_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value
