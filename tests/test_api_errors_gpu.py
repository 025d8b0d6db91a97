"""Error behaviour of the C-ABI (include/qpsk_b200.h): every misuse returns a negative status and leaves a message in
qpsk_b200_last_error(); nothing crashes, nothing silently falls back.  The reference itself has no error channel
(everything is void); the drop-in symbols keep that, the batch API adds status codes."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ERR_ARG, ERR_CUDA, ERR_STATE = -1, -2, -3


def _lib():
    from qpsk_b200 import capi
    return capi, capi.lib()


def _cfg(capi, L, **kw):
    cfg = capi.RxConfig()
    L.qpsk_b200_rx_default_config(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def test_rx_create_rejects_unsupported_profiles():
    capi, L = _lib()
    h = C.c_void_p()
    bad = [dict(frame_size=256), dict(ntaps=255), dict(rs=4800.0), dict(mode=7), dict(ub_mode=-1), dict(device=99)]
    for kw in bad:
        rc = L.qpsk_b200_rx_create(C.byref(_cfg(capi, L, **kw)), 32, 4, C.byref(h))
        assert rc in (ERR_ARG, ERR_CUDA) and not h.value, kw
        assert L.qpsk_b200_last_error()
    assert L.qpsk_b200_rx_create(C.byref(_cfg(capi, L)), 0, 4, C.byref(h)) == ERR_ARG
    assert L.qpsk_b200_rx_create(C.byref(_cfg(capi, L)), 32, 0, C.byref(h)) == ERR_ARG
    assert L.qpsk_b200_rx_create(None, 32, 4, C.byref(h)) == ERR_ARG
    assert L.qpsk_b200_rx_create(C.byref(_cfg(capi, L)), 32, 4, None) == ERR_ARG
    assert L.qpsk_b200_rx_destroy(None) == 0                      # like free(NULL)


def test_rx_process_and_read_argument_checks():
    import qpsk_b200
    capi, L = _lib()
    rx = qpsk_b200.Receiver(33, 4)
    pcm = np.zeros((33, 5 * 512), np.int16)
    out = np.zeros((33, 5 * 32), np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    # reading before any call, more frames than the context was made for, empty calls, null buffers
    assert L.qpsk_b200_rx_read(rx.h, capi.OUT_INDEX, p(out), 33 * 4 * 4) == ERR_STATE
    assert L.qpsk_b200_rx_process_host(rx.h, p(pcm), 5, p(out)) == ERR_ARG
    assert L.qpsk_b200_rx_process_host(rx.h, p(pcm), 0, p(out)) == ERR_ARG
    assert L.qpsk_b200_rx_process_host(rx.h, None, 2, p(out)) == ERR_ARG
    assert L.qpsk_b200_rx_process_device(rx.h, None, 2, None) == ERR_ARG
    assert L.qpsk_b200_rx_process_device(rx.h, C.c_void_p(0x1002), 2, None) == ERR_ARG      # misaligned device pointer
    assert b"aligned" in L.qpsk_b200_last_error()
    # a good call, then reads with the wrong size / of outputs that were not kept
    assert L.qpsk_b200_rx_process_host(rx.h, p(pcm), 2, p(out)) == 0
    idx = np.zeros((33, 2), np.int32)
    assert L.qpsk_b200_rx_read(rx.h, capi.OUT_INDEX, p(idx), idx.nbytes) == 0
    assert L.qpsk_b200_rx_read(rx.h, capi.OUT_INDEX, p(idx), idx.nbytes - 4) == ERR_ARG
    assert L.qpsk_b200_rx_read(rx.h, 77, p(idx), idx.nbytes) == ERR_ARG
    for what in (capi.OUT_SYMBOLS, capi.OUT_FIR, capi.OUT_FRAMES, capi.OUT_CRC_OK, capi.OUT_ROTATION):
        big = np.zeros(33 * 2 * 512 * 8, np.uint8)
        need = {capi.OUT_SYMBOLS: 33 * 2 * 128 * 8, capi.OUT_FIR: 33 * 2 * 512 * 8, capi.OUT_FRAMES: 33 * 2 * 32,
                capi.OUT_CRC_OK: 33 * 2, capi.OUT_ROTATION: 33 * 2}[what]
        assert L.qpsk_b200_rx_read(rx.h, what, p(big), need) == ERR_STATE, what
    n = C.c_ulonglong()
    assert L.qpsk_b200_rx_crc_counters(rx.h, C.byref(n), C.byref(n)) == ERR_STATE
    # the context is still usable after all of that
    assert L.qpsk_b200_rx_process_host(rx.h, p(pcm), 4, p(out)) == 0
    rx.close()


def test_fir_fft_bits_tx_argument_checks():
    import qpsk_b200
    capi, L = _lib()
    h = C.c_void_p()
    taps = qpsk_b200.rrc_make(127, 9600.0, 2400.0, 0.35)
    tp = taps.ctypes.data_as(C.c_void_p)
    assert L.qpsk_b200_fir_create(tp, 128, 4, 0, 0, C.byref(h)) == ERR_ARG          # only 127 and 256 taps are built
    assert L.qpsk_b200_fir_create(tp, 127, 0, 0, 0, C.byref(h)) == ERR_ARG
    assert L.qpsk_b200_fir_create(tp, 127, 4, 5, 0, C.byref(h)) == ERR_ARG
    assert L.qpsk_b200_fir_create(tp, 127, 4, 0, 64, C.byref(h)) == ERR_CUDA
    f = qpsk_b200.Fir(taps, 4)
    x = np.zeros((4, 16), np.complex64)
    assert L.qpsk_b200_fir_process_host(f.h, x.ctypes.data_as(C.c_void_p), 0) == ERR_ARG
    assert L.qpsk_b200_fir_process_host(f.h, None, 16) == ERR_ARG
    ms = C.c_float()
    assert L.qpsk_b200_fir_last_kernel_ms(f.h, C.byref(ms)) == ERR_STATE            # nothing timed yet
    f.close()
    # the reference's fft silently mis-computes lengths that are not powers of two (fft.c:38-96); here they are refused
    for n in (0, 1, 3, 100, 1000, 16384):
        assert L.qpsk_b200_fft_create(n, 0, C.byref(h)) == ERR_ARG, n
    buf = np.zeros((2, 44), np.uint8)
    bp = buf.ctypes.data_as(C.c_void_p)
    assert L.qpsk_b200_bits_interleave(bp, 44, 2, 2, 0) == ERR_ARG                   # dir is 0 or 1
    assert L.qpsk_b200_bits_interleave(bp, 8192, 2, 0, 0) == ERR_ARG                 # uint16_t bit count wraps (interleave.c:49)
    assert L.qpsk_b200_bits_interleave(None, 44, 2, 0, 0) == ERR_ARG
    assert L.qpsk_b200_frames_encode(bp, 22, 2, 2, bp, 0) == ERR_ARG                 # frames are 16 or 32 bytes
    assert L.qpsk_b200_frames_decode_rotated(bp, 32, 1, 1, bp, bp, None, 0) == ERR_ARG
    crc = np.zeros(2, np.uint16)
    assert L.qpsk_b200_bits_crc16(bp, 44, 0, crc.ctypes.data_as(C.c_void_p), 0) == ERR_ARG
    # transmit: any positive length is fine (tests/test_tx_gpu.py::test_any_length); zero is not
    tx = qpsk_b200.Transmitter(np.full(3, 1500.0, np.float32))
    assert tx.modulate(np.zeros((3, 33), np.uint8)).shape == (3, 132)
    with pytest.raises(qpsk_b200.QpskB200Error):
        tx.modulate(np.zeros((3, 0), np.uint8))
    tx.close()


def test_last_error_is_per_call_and_readable():
    capi, L = _lib()
    h = C.c_void_p()
    assert L.qpsk_b200_fft_create(12, 0, C.byref(h)) == ERR_ARG
    msg = L.qpsk_b200_last_error()
    assert b"12" in msg and b"powers of two" in msg


def test_contexts_release_their_device_memory():
    """create/destroy cycles of every context type leave the device's free memory where it was."""
    import torch
    import qpsk_b200
    taps = qpsk_b200.rrc_make(127, 9600.0, 2400.0, 0.35)
    pcm = np.zeros((64, 4 * 512), np.int16)

    def cycle():
        rx = qpsk_b200.Receiver(64, 4, decode_frames=True, resolve_rotation=True, estimate_offset=True, estimate_timing=True,
                                keep_fir=True, keep_symbols=True)
        rx.rx_frames(pcm)
        rx.estimate_offset(8)
        rx.close()
        f = qpsk_b200.Fir(taps, 64)
        f.filter(np.zeros((64, 4096), np.complex64))
        f.close()
        t = qpsk_b200.Fft(1024)
        t.argmax(np.zeros((16, 1024), np.complex64))
        t.close()
        tx = qpsk_b200.Transmitter(np.full(64, 1500.0, np.float32))
        tx.modulate(np.zeros((64, 256), np.uint8))
        tx.close()

    cycle()
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(20):
        cycle()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < (8 << 20), (free0, free1)
