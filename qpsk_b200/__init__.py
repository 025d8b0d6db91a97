"""qpsk_b200 -- B200-native (sm_100a) batched QPSK receiver behind the MonsieurETM/QPSK API.

The product is libqpsk_b200.so (qpsk_b200/csrc + qpsk_b200/host, C-ABI in include/); this
package is a thin ctypes mirror used by tests and bench.py.
"""
from . import _capi as capi  # noqa: F401
from ._capi import QpskB200Error  # noqa: F401
from .receiver import Receiver, unpack_dibits  # noqa: F401
from .fir import Fir, rrc_make  # noqa: F401
from .fft import Fft  # noqa: F401
from . import bits, shard  # noqa: F401
from .stream import receive_files  # noqa: F401
from .transmitter import Transmitter, awgn_device, bits_to_symbols  # noqa: F401
