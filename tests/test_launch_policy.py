"""The receiver's launch policy as host logic (no GPU): qpsk_b200_debug_plan runs the same planning functions
qpsk_b200_rx_process_device uses (rx_plan_chunks, rx_relay_blocks, rx_frame_blocks, rx_chase_plan in csrc/qpsk_b200.cu) for a
device it is told about.  Pins the decisions for BASELINE.json's shapes and the strong-scaling shares on a 148-SM B200; the
gpu tests assert the same plans through qpsk_b200_rx_last_plan on the device."""
import ctypes as C

import pytest


def plan(nchan, nframes, rs=2400.0, nsm=148):
    from qpsk_b200 import capi
    L = capi.lib()
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    capi.check(L.qpsk_b200_debug_plan(nsm, 0, nchan, nframes, rs, C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def test_baseline_shapes_on_a_b200():
    from qpsk_b200 import capi
    # configs[2], the headline: 2,048 channel groups = 6.92 waves of whole-stream CTAs, the loop in their spare warp
    assert plan(65536, 64) == (1, 1, capi.LOOP_FUSED)
    # its shares over 2 / 4 / 8 GPUs: frame blocks, the loop relayed from CTA to CTA
    assert plan(32768, 64) == (1, 4, capi.LOOP_RELAYED)
    assert plan(16384, 64) == (1, 8, capi.LOOP_RELAYED)
    assert plan(8192, 64) == (1, 16, capi.LOOP_RELAYED)
    # configs[1]: few channels x many frames = 32 frame chunks, one loop kernel on SMs of its own chasing them
    assert plan(1024, 256, rs=1200.0) == (32, 8, capi.LOOP_CHASING)
    # configs[0]: the reference's one channel, one frame per call
    assert plan(1, 1) == (1, 1, capi.LOOP_FUSED)


def test_policy_edges():
    from qpsk_b200 import capi
    # exact multiples of a wave stay whole-stream (nothing to even out)
    assert plan(9472, 64) == (1, 1, capi.LOOP_FUSED)
    assert plan(18944, 16) == (1, 1, capi.LOOP_FUSED)
    # a chunked call whose loop CTAs would need more than an eighth of the machine is not chased (the deadlock of round 2):
    # 4,096 channels = 32 loop CTAs > 148 / 8
    chunks, _, mode = plan(4096, 256, rs=1200.0)
    assert mode != capi.LOOP_CHASING
    # short calls are never chunked
    assert plan(300, 20)[0] == 1 and plan(64, 12)[0] == 1
    # a smaller device scales the waves: 16,384 channels on 74 SMs are 3.46 waves, relayed like 32,768 on 148
    assert plan(16384, 64, nsm=74) == (1, 4, capi.LOOP_RELAYED)


def test_bad_arguments():
    import qpsk_b200
    with pytest.raises(qpsk_b200.QpskB200Error):
        plan(0, 1)
    with pytest.raises(qpsk_b200.QpskB200Error):
        plan(16, 4, rs=300.0)
