"""The C-ABI library loads and exports every function include/*.h declares (no compute calls)."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for path in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        text = open(path).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        text = re.sub(r"^typedef[^;]*\(\s*\*[^;]*;", "", text, flags=re.M)   # function-pointer typedefs
        for m in re.finditer(r"^[A-Za-z_][\w\s\*]*?\b(\w+)\s*\([^;{]*\)\s*;", text, flags=re.M):
            if m.group(1) not in ("defined",):
                names.append((os.path.basename(path), m.group(1)))
    return names


def test_headers_declare_something():
    names = [n for _, n in declared_functions()]
    assert "qpsk_b200_rx_create" in names and "qpsk_b200_rx_process_host" in names


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from qpsk_b200 import capi
    L = ctypes.CDLL(capi.LIB_PATH)
    missing = [(h, n) for h, n in declared_functions() if not hasattr(L, n)]
    assert not missing, "declared in include/ but not exported by libqpsk_b200.so: %s" % missing


def test_no_gpu_means_loud_failure_not_fallback():
    """Without a usable sm_100 device the create call must fail with ERR_CUDA (there is no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from qpsk_b200 import capi
    L = capi.lib()
    cfg = capi.RxConfig()
    L.qpsk_b200_rx_default_config(ctypes.byref(cfg))
    h = ctypes.c_void_p()
    assert L.qpsk_b200_rx_create(ctypes.byref(cfg), 8, 4, ctypes.byref(h)) == -2
    assert L.qpsk_b200_last_error()
    assert L.qpsk_b200_device_count() == 0
