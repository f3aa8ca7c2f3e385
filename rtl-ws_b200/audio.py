"""ctypes binding of libb200audio.so: the reference's audio_main.h interface (include/rtlws_audio_compat.h)
over the GPU demodulator of libb200sdr.so.  Loading it pulls in libb200sdr.so; calling it needs a GPU."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200audio.so")
EXPORTED_SYMBOLS = [
    "audio_init", "audio_new_audio_available", "audio_get_audio_payload", "audio_fm_demodulator", "audio_close",
    "b200_audio_take_buffer", "b200_audio_buffer_len", "b200_audio_dropped_blocks", "b200_audio_reset_stream",
]

_lib = None


def lib(mode: int = C.DEFAULT_MODE) -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH, mode=mode)
    L.audio_init.restype = None
    L.audio_close.restype = None
    L.audio_get_audio_payload.argtypes = [C.c_void_p, C.c_int]
    L.audio_fm_demodulator.restype = None
    L.audio_fm_demodulator.argtypes = [C.c_void_p, C.c_int]
    L.b200_audio_take_buffer.argtypes = [C.c_void_p, C.c_int]
    L.b200_audio_dropped_blocks.restype = C.c_longlong
    L.b200_audio_reset_stream.restype = None
    _lib = L
    return L


def demodulate(signal: np.ndarray) -> None:
    """audio_fm_demodulator(const cmplx_s32*, int): one decimator block into the pool."""
    signal = np.ascontiguousarray(signal, dtype=np.int32).reshape(-1, 2)
    lib().audio_fm_demodulator(signal.ctypes.data, len(signal))


def get_payload(n_bytes: int) -> np.ndarray:
    """audio_get_audio_payload(buf, n_bytes) -> the float32 samples it copied."""
    buf = np.zeros(n_bytes // 4, dtype=np.float32)
    got = lib().audio_get_audio_payload(buf.ctypes.data, n_bytes)
    return buf[:got // 4]


def take_buffer() -> np.ndarray | None:
    n = lib().b200_audio_buffer_len()
    if n <= 0:
        return None
    buf = np.empty(n, dtype=np.float32)
    got = lib().b200_audio_take_buffer(buf.ctypes.data, n)
    return buf if got == n else None
