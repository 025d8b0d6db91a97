"""ctypes bindings for the CPU checker (oracle/).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  `Oracle` wraps liboracle.so (the restatement, qpsk_oracle.c); `Ref`
wraps oracle/_ref/libref_*.so (the unmodified reference compiled by oracle/Makefile).
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.environ.get("QPSK_REF_DIR", "/root/reference")
TAU = 2.0 * math.pi
MAX_TAPS, MAX_FRAME = 512, 512


def build(quiet=True):
    """Compile liboracle.so and, when the reference tree is present, oracle/_ref/*.so."""
    cmd = ["make", "-C", HERE, "CC=gcc", "REF_DIR=" + REF_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if not quiet:
        print(r.stdout)


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class _Loop(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("phase", "freq", "max_freq", "min_freq", "damping", "loop_bw", "alpha", "beta")]


class _CF(C.Structure):
    _fields_ = [("re", C.c_float), ("im", C.c_float)]


class _Profile(C.Structure):
    _fields_ = [("sps", C.c_int), ("frame_size", C.c_int), ("nsym", C.c_int), ("ntaps", C.c_int), ("ub_mode", C.c_int),
                ("fs", C.c_float), ("rs", C.c_float), ("center", C.c_float), ("taps", C.c_float * MAX_TAPS),
                ("rx_rect", _CF), ("rot45", _CF), ("loop0", _Loop)]


class _RxState(C.Structure):
    _fields_ = [("rx_phase", _CF), ("fir_mem", _CF * MAX_TAPS), ("dec", _CF * MAX_FRAME), ("phase", C.c_float), ("freq", C.c_float)]


class _TxState(C.Structure):
    _fields_ = [("tx_phase", _CF), ("tx_rect", _CF), ("fir_mem", _CF * MAX_TAPS)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_glibc_sinf.restype = C.c_float
        L.orc_glibc_sinf.argtypes = [C.c_float]
        L.orc_glibc_cosf.restype = C.c_float
        L.orc_glibc_cosf.argtypes = [C.c_float]
        L.orc_crc16.restype = C.c_uint16
        L.orc_phase_detector.restype = C.c_float
        L.orc_phase_detector.argtypes = [_CF]
        L.orc_fft_argmax.restype = C.c_int
        L.orc_interleave_prime.restype = C.c_int
        L.orc_profile_init.argtypes = [C.POINTER(_Profile), C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_float, C.c_int]
        L.orc_rrc_make.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_float, C.c_float, C.c_float]
        L.orc_loop_create.argtypes = [C.POINTER(_Loop), C.c_float, C.c_float, C.c_float]
        L.orc_loop_advance.argtypes = [C.POINTER(_Loop), C.c_float]
        L.orc_loop_set_frequency.argtypes = [C.POINTER(_Loop), C.c_float]
        L.orc_loop_set_phase.argtypes = [C.POINTER(_Loop), C.c_float]
        L.orc_tx_state_init.argtypes = [C.POINTER(_TxState), C.c_float, C.c_float]
        _lib = L
    return _lib


class Oracle:
    """One modem profile of the restatement.  Defaults are the reference's (qpsk.h:16-23, rrc_fir.h:13)."""

    def __init__(self, fs=9600.0, rs=2400.0, center=1500.0, rrc_alpha=0.35, ntaps=127, frame_size=512,
                 loop_bw=None, ub_mode=0, slice_diagonal=False):
        self.L = lib()
        self.p = _Profile()
        if loop_bw is None:
            loop_bw = np.float32(TAU / 100.0)  # qpsk.c:302
        self.L.orc_profile_init(C.byref(self.p), fs, rs, center, rrc_alpha, ntaps, frame_size, float(loop_bw), ub_mode)
        if slice_diagonal:
            self.L.orc_profile_slice_diagonal(C.byref(self.p), 1)
        self.sps, self.frame_size, self.nsym, self.ntaps = self.p.sps, self.p.frame_size, self.p.nsym, self.p.ntaps
        self.fs = fs

    @property
    def taps(self):
        return np.ctypeslib.as_array(self.p.taps)[: self.ntaps].copy()

    def new_states(self, nchan):
        st = (_RxState * nchan)()
        for i in range(nchan):
            self.L.orc_rx_state_init(C.byref(self.p), C.byref(st[i]))
        return st

    def rx_run(self, pcm, states=None, want=("fir", "index", "dec", "costas", "dibit", "phase", "freq")):
        """pcm int16 [C, F*frame_size] -> dict of per-stage outputs (complex64 / int32 / uint8 / float32)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        nchan, n = pcm.shape
        F = n // self.frame_size
        assert F * self.frame_size == n
        if states is None:
            states = self.new_states(nchan)
        out = {}
        if "fir" in want: out["fir"] = np.zeros((nchan, F * self.frame_size), np.complex64)
        if "index" in want: out["index"] = np.zeros((nchan, F), np.int32)
        if "dec" in want: out["dec"] = np.zeros((nchan, F * self.nsym), np.complex64)
        if "costas" in want: out["costas"] = np.zeros((nchan, F * self.nsym), np.complex64)
        if "dibit" in want: out["dibit"] = np.zeros((nchan, F * self.nsym), np.uint8)
        if "phase" in want: out["phase"] = np.zeros((nchan, F), np.float32)
        if "freq" in want: out["freq"] = np.zeros((nchan, F), np.float32)
        g = lambda k: (out[k].ctypes.data_as(C.c_void_p) if k in out else None)
        self.L.orc_rx_run(C.byref(self.p), states, pcm.ctypes.data_as(C.c_void_p), nchan, F,
                          g("fir"), g("index"), g("dec"), g("costas"), g("dibit"), g("phase"), g("freq"))
        out["states"] = states
        return out

    def fir(self, taps, memory, samples):
        """rrc_fir semantics on complex64 arrays, in place (memory: [ntaps], samples: [n])."""
        taps = np.ascontiguousarray(taps, np.float32)
        self.L.orc_rrc_fir(taps.ctypes.data_as(C.c_void_p), len(taps), memory.ctypes.data_as(C.c_void_p),
                           samples.ctypes.data_as(C.c_void_p), len(samples))

    def rrc_make(self, ntaps, fs, rs, alpha):
        t = np.zeros(ntaps, np.float32)
        self.L.orc_rrc_make(_ptr(t, C.c_float), ntaps, fs, rs, alpha)
        return t

    def new_tx(self, carrier_hz):
        s = _TxState()
        self.L.orc_tx_state_init(C.byref(s), carrier_hz, self.fs)
        return s

    def set_tx_carrier(self, tx, carrier_hz):
        self.L.orc_tx_set_carrier.argtypes = [C.c_void_p, C.c_float, C.c_float]
        self.L.orc_tx_set_carrier.restype = None
        self.L.orc_tx_set_carrier(C.byref(tx), carrier_hz, self.fs)

    def packet_mod(self, tx, bits):
        """bits: int32 [2*length] (one bit per int, qpsk.c:273) -> int16 [length*sps]."""
        bits = np.ascontiguousarray(bits, np.int32)
        length = len(bits) // 2
        out = np.zeros(length * self.sps, np.int16)
        n = self.L.orc_qpsk_packet_mod(C.byref(self.p), C.byref(tx), out.ctypes.data_as(C.c_void_p),
                                       bits.ctypes.data_as(C.c_void_p), length)
        assert n == len(out)
        return out

    # -- bit stages / FFT -------------------------------------------------------------------
    def crc16(self, data):
        data = np.ascontiguousarray(data, np.uint8)
        return int(self.L.orc_crc16(data.ctypes.data_as(C.c_void_p), len(data)))

    def interleave(self, data, direction):
        buf = np.array(data, np.uint8, copy=True)
        self.L.orc_interleave(buf.ctypes.data_as(C.c_void_p), len(buf), direction)
        return buf

    def scramble_stream(self, dibits, reg=0x4A80):
        d = np.array(dibits, np.uint8, copy=True)
        r = C.c_uint16(reg)
        one = C.c_uint8()
        for i in range(len(d)):
            one.value = int(d[i])
            self.L.orc_scramble_dibit(C.byref(r), C.byref(one))
            d[i] = one.value
        return d, r.value

    def frame_encode(self, payload, nbytes):
        """payload uint8 [nbytes-2] -> scrambled, interleaved dibits uint8 [4*nbytes] (unpinned composition)."""
        payload = np.ascontiguousarray(payload, np.uint8)
        assert len(payload) == nbytes - 2
        out = np.zeros(4 * nbytes, np.uint8)
        self.L.orc_frame_encode(payload.ctypes.data_as(C.c_void_p), nbytes, out.ctypes.data_as(C.c_void_p))
        return out

    def frame_decode(self, dibits, nbytes):
        dibits = np.ascontiguousarray(dibits, np.uint8)
        assert len(dibits) == 4 * nbytes
        frame = np.zeros(nbytes, np.uint8)
        ok = self.L.orc_frame_decode(dibits.ctypes.data_as(C.c_void_p), nbytes, frame.ctypes.data_as(C.c_void_p))
        return frame, bool(ok)

    def awgn(self, pcm, sigma, seed, first_sample=0, first_channel=0):
        """pcm int16 [C, T] -> noisy copy; the generator's counter-based noise (extension)."""
        out = np.ascontiguousarray(pcm, np.int16).copy()
        sg = np.broadcast_to(np.asarray(sigma, np.float32), (out.shape[0],))
        self.L.orc_awgn.argtypes = [C.c_void_p, C.c_longlong, C.c_float, C.c_uint64, C.c_longlong, C.c_int]
        self.L.orc_awgn.restype = None
        for c in range(out.shape[0]):
            self.L.orc_awgn(out[c].ctypes.data_as(C.c_void_p), out.shape[1], float(sg[c]), int(seed), int(first_sample), first_channel + c)
        return out

    def timing_sum(self, fir_frames):
        """fir complex64 [..., frame_size] (rx_run's "fir") -> the extension's timing statistic complex64 [...]."""
        x = np.ascontiguousarray(fir_frames, np.complex64)
        flat = x.reshape(-1, x.shape[-1])
        out = np.zeros(flat.shape[0], np.complex64)
        for i in range(flat.shape[0]):
            self.L.orc_timing_sum(flat[i].ctypes.data_as(C.c_void_p), flat.shape[1], self.sps,
                                  out[i:i + 1].ctypes.data_as(C.c_void_p))
        return out.reshape(x.shape[:-1])

    def frame_decode_rotated(self, dibits, nbytes):
        """-> (frame, quarter turns undone 0..3 or -1 when no rotation passes the CRC)."""
        dibits = np.ascontiguousarray(dibits, np.uint8)
        assert len(dibits) == 4 * nbytes
        frame = np.zeros(nbytes, np.uint8)
        self.L.orc_frame_decode_rotated.restype = C.c_int
        r = self.L.orc_frame_decode_rotated(dibits.ctypes.data_as(C.c_void_p), nbytes, frame.ctypes.data_as(C.c_void_p))
        return frame, int(r)

    def rotate_dibits(self, dibits, quarter_turns):
        self.L.orc_rotate_dibit.restype = C.c_uint8
        self.L.orc_rotate_dibit.argtypes = [C.c_uint8, C.c_int]
        return np.array([self.L.orc_rotate_dibit(int(d), int(quarter_turns)) for d in np.asarray(dibits).ravel()],
                        np.uint8).reshape(np.shape(dibits))

    def fftn(self, x, inverse=False):
        x = np.ascontiguousarray(x, np.complex128)
        out = np.zeros_like(x)
        fn = self.L.orc_ifftn if inverse else self.L.orc_fftn
        fn(x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), len(x))
        return out

    def fft_argmax(self, X):
        X = np.ascontiguousarray(X, np.complex128)
        m = C.c_double()
        k = self.L.orc_fft_argmax(X.ctypes.data_as(C.c_void_p), len(X), C.byref(m))
        return k, m.value


class Ref:
    """The unmodified reference compiled into oracle/_ref (present only where it was prebuilt)."""

    def __init__(self, flavour="2400"):
        path = os.path.join(HERE, "_ref", "libref_%s.so" % flavour)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.L = L = C.CDLL(path)
        L.ref_rx_time.restype = C.c_double
        L.ref_phase_detector.restype = C.c_float
        L.ref_phase_detector.argtypes = [C.c_float, C.c_float]
        L.ref_qpsk_demod.argtypes = [C.c_float, C.c_float, C.c_void_p]
        L.ref_tx_reset.argtypes = [C.c_double]
        L.rrc_make.argtypes = [C.c_float, C.c_float, C.c_float]
        for n in ("get_phase", "get_frequency", "get_alpha", "get_beta", "get_loop_bandwidth", "get_damping_factor",
                  "get_max_freq", "get_min_freq"):
            getattr(L, n).restype = C.c_float
        for n in ("set_phase", "set_frequency", "set_alpha", "set_beta", "set_loop_bandwidth", "set_damping_factor",
                  "set_max_freq", "set_min_freq", "advance_loop"):
            getattr(L, n).argtypes = [C.c_float]
        L.create_control_loop.argtypes = [C.c_float, C.c_float, C.c_float]
        self.sps, self.ntaps, self.frame_size = L.ref_sps(), L.ref_ntaps(), L.ref_frame_size()
        self.nsym = self.frame_size // self.sps
        L.ref_init()

    @staticmethod
    def available(flavour="2400"):
        return os.path.exists(os.path.join(HERE, "_ref", "libref_%s.so" % flavour))

    def layout_ok(self):
        return bool(self.L.ref_layout_ok())

    @property
    def taps(self):
        t = np.zeros(self.ntaps, np.float32)
        self.L.ref_get_taps(t.ctypes.data_as(C.c_void_p))
        return t

    def rx_run(self, pcm):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        nchan, n = pcm.shape
        F = n // self.frame_size
        out = {"fir": np.zeros((nchan, F * self.frame_size), np.complex64),
               "dec": np.zeros((nchan, F * self.nsym), np.complex64),
               "costas": np.zeros((nchan, F * self.nsym), np.complex64),
               "dibit": np.zeros((nchan, F * self.nsym), np.uint8),
               "phase": np.zeros((nchan, F), np.float32), "freq": np.zeros((nchan, F), np.float32)}
        v = lambda k: out[k].ctypes.data_as(C.c_void_p)
        self.L.ref_rx_run(pcm.ctypes.data_as(C.c_void_p), nchan, F, v("fir"), v("dec"), v("costas"), v("dibit"), v("phase"), v("freq"))
        return out

    def rx_time(self, pcm_1ch, reps):
        pcm = np.ascontiguousarray(pcm_1ch, np.int16).reshape(-1)
        F = len(pcm) // self.frame_size
        return self.L.ref_rx_time(pcm.ctypes.data_as(C.c_void_p), F, reps)

    def fir(self, memory, samples):
        self.L.ref_rrc_fir(memory.ctypes.data_as(C.c_void_p), samples.ctypes.data_as(C.c_void_p), len(samples))

    def packet_mod(self, bits):
        bits = np.ascontiguousarray(bits, np.int32)
        length = len(bits) // 2
        out = np.zeros(length * self.sps, np.int16)
        n = self.L.ref_packet_mod(out.ctypes.data_as(C.c_void_p), bits.ctypes.data_as(C.c_void_p), length)
        assert n == len(out)
        return out

    def tx_reset(self, carrier_hz):
        self.L.ref_tx_reset(float(carrier_hz))


class RefAlg:
    """algorithms/{fft,crc16,interleave,bit-scramble}.c of the unmodified reference."""

    def __init__(self, flavour="alg"):
        path = os.path.join(HERE, "_ref", "libref_%s.so" % flavour)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.L = C.CDLL(path)
        self.L.crc16.restype = C.c_uint16

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, "_ref", "libref_alg.so"))

    def crc16(self, data):
        data = np.ascontiguousarray(data, np.uint8)
        return int(self.L.crc16(data.ctypes.data_as(C.c_void_p), len(data)))

    def interleave(self, data, direction):
        buf = np.array(data, np.uint8, copy=True)
        self.L.interleave(buf.ctypes.data_as(C.c_void_p), len(buf), direction)
        return buf

    def scramble_stream(self, dibits, which=0):
        d = np.array(dibits, np.uint8, copy=True)
        self.L.scramble_init(which)
        one = C.c_uint8()
        for i in range(len(d)):
            one.value = int(d[i])
            self.L.scramble(C.byref(one), which)
            d[i] = one.value
        return d

    def fftn(self, x, inverse=False):
        x = np.ascontiguousarray(x, np.complex128)
        out = np.zeros_like(x)
        fn = self.L.ifftn if inverse else self.L.fftn
        fn(x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), len(x))
        return out
