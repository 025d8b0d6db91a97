"""ctypes binding of libqpsk_b200.so (the C-ABI declared in include/qpsk_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is usable, the
calls raise.  The library is built in-tree by `make -C qpsk_b200` (see __graft_entry__.build()).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QPSK_B200_LIB") or os.path.join(HERE, "libqpsk_b200.so")      # the override is for A/B builds of the same library


class QpskB200Error(RuntimeError):
    pass


class RxConfig(C.Structure):
    _fields_ = [("fs", C.c_float), ("rs", C.c_float), ("center", C.c_float), ("rrc_alpha", C.c_float),
                ("loop_bw", C.c_float), ("ntaps", C.c_int), ("frame_size", C.c_int), ("mode", C.c_int),
                ("ub_mode", C.c_int), ("flags", C.c_int), ("device", C.c_int)]


class StreamStats(C.Structure):
    _fields_ = [("frames", C.c_longlong), ("seconds", C.c_double), ("read_seconds", C.c_double), ("wait_seconds", C.c_double),
                ("readers", C.c_int)]


STREAM_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.POINTER(C.c_uint8))

MODE_EXACT, MODE_FAST = 0, 1
UB_ALIAS, UB_CLAMP, UB_PHASE, UB_TAU = 0, 1, 2, 3
KEEP_FIR, KEEP_SYMBOLS, DECODE_FRAMES, NO_FUSE, RESOLVE_ROTATION, SLICE_DIAGONAL, ESTIMATE_OFFSET, ESTIMATE_TIMING, NO_CHUNK, PREROTATE_OFFSET, TRANSIENT_SYMBOLS = 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024
LOOP_STANDALONE, LOOP_FUSED, LOOP_RELAYED, LOOP_CHASING, LOOP_FOLLOWING = range(5)
OUT_DIBITS, OUT_INDEX, OUT_TRACK, OUT_DEC, OUT_SYMBOLS, OUT_FIR, OUT_TAPS, OUT_FRAMES, OUT_CRC_OK, OUT_ROTATION, OUT_OFFSET_BIN, OUT_OFFSET_HZ, OUT_TIMING_SUM, OUT_TIMING_TAU = range(14)

_lib = None


def lib():
    """Load libqpsk_b200.so; raises QpskB200Error when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QpskB200Error("%s not found: run `make -C qpsk_b200` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.qpsk_b200_last_error.restype = C.c_char_p
    L.qpsk_b200_rx_default_config.argtypes = [C.POINTER(RxConfig)]
    L.qpsk_b200_rx_default_config.restype = None
    L.qpsk_b200_rx_create.argtypes = [C.POINTER(RxConfig), C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.qpsk_b200_rx_destroy.argtypes = [C.c_void_p]
    L.qpsk_b200_rx_reset.argtypes = [C.c_void_p]
    L.qpsk_b200_rx_process_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.qpsk_b200_rx_process_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.qpsk_b200_rx_submit_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.qpsk_b200_rx_wait.argtypes = [C.c_void_p]
    L.qpsk_b200_rx_probe_copy_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.qpsk_b200_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.qpsk_b200_host_free.argtypes = [C.c_void_p]
    L.qpsk_b200_stream_open.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.qpsk_b200_stream_frames.argtypes = [C.c_void_p]
    L.qpsk_b200_stream_frames.restype = C.c_longlong
    L.qpsk_b200_stream_run.argtypes = [C.c_void_p, STREAM_SINK, C.c_void_p, C.POINTER(StreamStats)]
    L.qpsk_b200_stream_close.argtypes = [C.c_void_p]
    L.qpsk_b200_rx_sync.argtypes = [C.c_void_p]
    L.qpsk_b200_rx_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    L.qpsk_b200_rx_output_bytes.argtypes = [C.c_void_p, C.c_int]
    L.qpsk_b200_rx_output_bytes.restype = C.c_size_t
    L.qpsk_b200_rx_device_dibits.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    L.qpsk_b200_rx_launch_count.argtypes = [C.c_void_p]
    L.qpsk_b200_rx_launch_count.restype = C.c_longlong
    L.qpsk_b200_rx_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.qpsk_b200_rx_last_plan.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.qpsk_b200_debug_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.qpsk_b200_rx_estimate_offset.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.qpsk_b200_probe_fp32.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    _bind_fir(L)
    _bind_fft(L)
    _bind_bits(L)
    _bind_tx(L)
    _bind_channel(L)
    _lib = L
    return L


def probe_fp32(device=0, fused=False):
    """Measured FP32-pipe ceiling of the device: complex tap-updates per second (exact = FMUL2+FADD2, fused = FFMA2)."""
    rate, ms = C.c_double(), C.c_float()
    check(lib().qpsk_b200_probe_fp32(device, 1 if fused else 0, C.byref(rate), C.byref(ms)))
    return rate.value


def check(rc):
    if rc != 0:
        raise QpskB200Error("qpsk_b200 error %d: %s" % (rc, lib().qpsk_b200_last_error().decode()))


def _bind_fir(L):
    L.qpsk_b200_rrc_make.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float]
    L.qpsk_b200_fir_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.qpsk_b200_fir_destroy.argtypes = [C.c_void_p]
    L.qpsk_b200_fir_reset.argtypes = [C.c_void_p]
    L.qpsk_b200_fir_process_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.qpsk_b200_fir_process_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.qpsk_b200_fir_get_memory.argtypes = [C.c_void_p, C.c_void_p]
    L.qpsk_b200_fir_set_memory.argtypes = [C.c_void_p, C.c_void_p]
    L.qpsk_b200_fir_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]


def _bind_fft(L):
    L.qpsk_b200_fft_create.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.qpsk_b200_fft_destroy.argtypes = [C.c_void_p]
    L.qpsk_b200_fft_argmax_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.qpsk_b200_fft_argmax_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.qpsk_b200_fft_transform_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.qpsk_b200_fft_transform_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.qpsk_b200_fft_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.qpsk_b200_fft_big_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]


def _bind_bits(L):
    L.qpsk_b200_bits_crc16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.qpsk_b200_bits_interleave.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.qpsk_b200_bits_scramble.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.qpsk_b200_frames_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.qpsk_b200_frames_decode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.qpsk_b200_frames_decode_rotated.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.qpsk_b200_rx_crc_counters.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]


def _bind_channel(L):
    L.qpsk_b200_tx_set_carrier.argtypes = [C.c_void_p, C.c_void_p]
    L.qpsk_b200_channel_awgn_device.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_ulonglong, C.c_longlong, C.c_int,
                                                C.c_int, C.c_void_p]


def _bind_tx(L):
    L.qpsk_b200_tx_create.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.qpsk_b200_tx_destroy.argtypes = [C.c_void_p]
    L.qpsk_b200_tx_reset.argtypes = [C.c_void_p]
    L.qpsk_b200_tx_process_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.qpsk_b200_tx_process_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.qpsk_b200_tx_end_packet.argtypes = [C.c_void_p]
