"""ctypes binding of libb200replay.so: the reference's rtl_sensor.h interface over a replayed capture
(include/rtl_sensor_replay.h).  Host code only -- no GPU needed to load or drive it."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libb200replay.so")
BUFFER_BYTES = 262144
BUFFERS = 15
EXPORTED_SYMBOLS = [
    "rtl_init", "rtl_set_frequency", "rtl_set_sample_rate", "rtl_set_gain", "rtl_freq", "rtl_sample_rate", "rtl_gain",
    "rtl_read_async", "rtl_cancel", "rtl_close",
    "b200_replay_set_capture", "b200_replay_gate", "b200_replay_delivered_bytes",
]
READ_CALLBACK = C.CFUNCTYPE(None, C.POINTER(C.c_ubyte), C.c_uint32, C.c_void_p)

_lib = None


def lib(mode: int = C.DEFAULT_MODE) -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH, mode=mode)
    vp = C.c_void_p
    L.rtl_init.argtypes = [C.POINTER(vp), C.c_int]
    L.rtl_set_frequency.argtypes = [vp, C.c_uint32]
    L.rtl_set_sample_rate.argtypes = [vp, C.c_uint32]
    L.rtl_set_gain.argtypes = [vp, C.c_double]
    L.rtl_freq.restype = C.c_uint32
    L.rtl_freq.argtypes = [vp]
    L.rtl_sample_rate.restype = C.c_uint32
    L.rtl_sample_rate.argtypes = [vp]
    L.rtl_gain.restype = C.c_double
    L.rtl_gain.argtypes = [vp]
    L.rtl_read_async.argtypes = [vp, READ_CALLBACK, vp]
    L.rtl_cancel.restype = None
    L.rtl_cancel.argtypes = [vp]
    L.rtl_close.restype = None
    L.rtl_close.argtypes = [vp]
    L.b200_replay_set_capture.argtypes = [C.c_int, vp, C.c_int64, C.c_int, C.c_int]
    L.b200_replay_gate.argtypes = [C.c_int, C.c_int]
    L.b200_replay_delivered_bytes.restype = C.c_int64
    L.b200_replay_delivered_bytes.argtypes = [C.c_int]
    _lib = L
    return L


class VirtualDongle:
    """One virtual dongle: a capture (numpy uint8, kept alive here) behind the rtl_* calls."""

    def __init__(self, index: int, iq: np.ndarray, loops: int = 1, realtime: bool = False):
        self.index = index
        self.iq = np.ascontiguousarray(iq, dtype=np.uint8).reshape(-1)
        L = lib()
        if L.b200_replay_set_capture(index, self.iq.ctypes.data, self.iq.size, loops, int(realtime)) != 0:
            raise ValueError("b200_replay_set_capture rejected the capture")
        dev = C.c_void_p()
        if L.rtl_init(C.byref(dev), index) != 0:
            raise RuntimeError("rtl_init failed")
        self.dev = dev

    def read_async(self, on_buffer) -> int:
        """Blocks until the capture is exhausted or cancel() is called; on_buffer(bytes ndarray view)."""
        def cb(buf, n, user):
            on_buffer(np.ctypeslib.as_array(buf, shape=(n,)))
        self._cb = READ_CALLBACK(cb)
        return lib().rtl_read_async(self.dev, self._cb, None)

    def cancel(self):
        lib().rtl_cancel(self.dev)

    def delivered_bytes(self) -> int:
        return int(lib().b200_replay_delivered_bytes(self.index))

    def close(self):
        if getattr(self, "dev", None):
            lib().rtl_close(self.dev)
            self.dev = None
            lib().b200_replay_set_capture(self.index, None, 0, 1, 0)
