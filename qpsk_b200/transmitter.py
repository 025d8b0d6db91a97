"""Host-side mirror of the reference transmit path (qpsk_packet_mod / tx_frame, qpsk.c:225-285) over many channels."""
import ctypes as C

import numpy as np

from . import _capi as capi


def bits_to_symbols(tx_bits):
    """tx_bits [..., 2n] (one bit per entry, qpsk.c:273) -> constellation index per symbol,
    (tx_bits[2k] << 1) | tx_bits[2k+1]  (qpsk.c:270 with dibit[0] = tx_bits[s+1], dibit[1] = tx_bits[s])."""
    b = np.asarray(tx_bits).astype(np.uint8) & 1
    return ((b[..., 0::2] << 1) | b[..., 1::2]).astype(np.uint8)


class Transmitter:
    def __init__(self, carrier_hz, rs=2400.0, fs=9600.0, rrc_alpha=0.35, packet_symbols=256, device=0):
        self.L = capi.lib()
        self.carrier = np.ascontiguousarray(carrier_hz, np.float32)
        self.nchan = len(self.carrier)
        self.sps = int(fs / rs)
        self.h = C.c_void_p()
        capi.check(self.L.qpsk_b200_tx_create(fs, rs, rrc_alpha, self.carrier.ctypes.data_as(C.c_void_p), self.nchan,
                                              packet_symbols, device, C.byref(self.h)))

    def close(self):
        if self.h:
            self.L.qpsk_b200_tx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        capi.check(self.L.qpsk_b200_tx_reset(self.h))

    def modulate(self, symbols):
        """symbols uint8 [C, nsym] (constellation indices) -> int16 PCM [C, nsym*sps]."""
        s = np.ascontiguousarray(symbols, np.uint8)
        assert s.shape[0] == self.nchan
        pcm = np.zeros((self.nchan, s.shape[1] * self.sps), np.int16)
        capi.check(self.L.qpsk_b200_tx_process_host(self.h, s.ctypes.data_as(C.c_void_p), s.shape[1], pcm.ctypes.data_as(C.c_void_p)))
        return pcm

    def modulate_device(self, d_symbols, nsym, d_pcm, stream=None):
        capi.check(self.L.qpsk_b200_tx_process_device(self.h, C.c_void_p(d_symbols), nsym, C.c_void_p(d_pcm),
                                                      C.c_void_p(stream) if stream else None))


    def end_packet(self):
        """End the current tx_frame call now: the up-mix phasor is renormalised (qpsk.c:253) and the packet position restarts."""
        capi.check(self.L.qpsk_b200_tx_end_packet(self.h))

    def set_carrier(self, carrier_hz):
        """New per-channel carriers from the next call on (phase-continuous): steps of a Doppler ramp."""
        c = np.ascontiguousarray(carrier_hz, np.float32)
        assert c.shape == (self.nchan,)
        capi.check(self.L.qpsk_b200_tx_set_carrier(self.h, c.ctypes.data_as(C.c_void_p)))


def awgn_device(d_pcm, nchan, nsamples, sigma, seed, first_sample=0, first_channel=0, device=0, stream=None):
    """Counter-based test-channel noise on int16 PCM in HBM, in place (qpsk_b200_channel_awgn_device)."""
    sg = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma, np.float32), (nchan,)))
    capi.check(capi.lib().qpsk_b200_channel_awgn_device(C.c_void_p(d_pcm), nchan, nsamples, sg.ctypes.data_as(C.c_void_p), int(seed),
                                                        int(first_sample), int(first_channel), device, C.c_void_p(stream) if stream else None))
