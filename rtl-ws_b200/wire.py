"""Websocket wire formats (main.c:74-111) over the b200_wire_* entry points of libb200sdr.so.

Host helpers (header, fragment table, drain map) need no GPU; the batched emitters run on the
device and return tensors whose rows are ready for lws_write."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import binding as _b

SPECTRUM_HEADER_MAX = 48
AUDIO_FRAGMENTS = 8
AUDIO_FRAGMENT_BYTES = 2048
AUDIO_MESSAGE_BYTES = 8 + AUDIO_FRAGMENTS * AUDIO_FRAGMENT_BYTES
AUDIO_MESSAGE_FLOATS = AUDIO_FRAGMENTS * AUDIO_FRAGMENT_BYTES // 4
BINARY, CONTINUATION, NO_FIN = 1, 2, 0x40
REFERENCE_DRAIN = 1
POOL_BUFFER_LEN = 5120          # audio_main.c:90,100 at 2.048 MS/s, R = 10


def spectrum_header(freq_hz: int, sample_rate_hz: int, gain_db: int) -> bytes:
    buf = C.create_string_buffer(SPECTRUM_HEADER_MAX)
    n = _b.lib().b200_wire_spectrum_header(buf, SPECTRUM_HEADER_MAX, freq_hz, sample_rate_hz, gain_db)
    if n < 0:
        raise _b.B200Error(f"b200_wire_spectrum_header -> {n}: {_b.last_error()}")
    return buf.raw[:n]


def spectrum_message(freq_hz: int, sample_rate_hz: int, gain_db: int, payload: np.ndarray) -> bytes:
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    dst = np.zeros(SPECTRUM_HEADER_MAX + payload.size, dtype=np.uint8)
    n = _b.lib().b200_wire_spectrum_message(dst.ctypes.data, dst.size, freq_hz, sample_rate_hz, gain_db,
                                            payload.ctypes.data, payload.size)
    if n < 0:
        raise _b.B200Error(f"b200_wire_spectrum_message -> {n}: {_b.last_error()}")
    return dst[:n].tobytes()


def audio_fragments():
    """[(offset, length, flags)] of the eight lws_write calls that make one audio message."""
    out = []
    for i in range(AUDIO_FRAGMENTS):
        off, ln, fl = C.c_int32(), C.c_int32(), C.c_int32()
        _b._check(_b.lib().b200_wire_audio_fragment(i, C.byref(off), C.byref(ln), C.byref(fl)), "b200_wire_audio_fragment")
        out.append((off.value, ln.value, fl.value))
    return out


def reference_drain_index(wire_sample: int, buffer_len: int = POOL_BUFFER_LEN) -> int:
    return int(_b.lib().b200_wire_reference_drain_index(wire_sample, buffer_len))


def spectrum_messages(payload, freq_hz, sample_rate_hz, gain_db, stream=None):
    """payload: cuda uint8 [n_streams, n_bins].  -> (cuda uint8 [n_streams, stride], lengths int32 ndarray)."""
    torch = _b._torch()
    assert payload.is_cuda and payload.dtype == torch.uint8 and payload.dim() == 2 and payload.stride(1) == 1
    n_streams, n_bins = payload.shape
    f = np.ascontiguousarray(np.broadcast_to(np.asarray(freq_hz, dtype=np.uint32), (n_streams,)))
    r = np.ascontiguousarray(np.broadcast_to(np.asarray(sample_rate_hz, dtype=np.uint32), (n_streams,)))
    g = np.ascontiguousarray(np.broadcast_to(np.asarray(gain_db, dtype=np.int32), (n_streams,)))
    stride = (SPECTRUM_HEADER_MAX + n_bins + 15) // 16 * 16
    msgs = torch.empty((n_streams, stride), dtype=torch.uint8, device=payload.device)
    lens = np.zeros(n_streams, dtype=np.int32)
    _b._check(_b.lib().b200_wire_spectrum_messages(payload.data_ptr(), payload.stride(0), n_streams, n_bins,
                                                   f.ctypes.data, r.ctypes.data, g.ctypes.data, msgs.data_ptr(), stride,
                                                   lens.ctypes.data, _b._stream_ptr(stream)),
              "b200_wire_spectrum_messages")
    return msgs, lens


def audio_messages(audio, first_wire_sample: int, n_messages: int, flags: int = 0, buffer_len: int = POOL_BUFFER_LEN,
                   stream=None):
    """audio: cuda float32 [n_streams, n].  -> cuda uint8 [n_streams, n_messages, AUDIO_MESSAGE_BYTES]."""
    torch = _b._torch()
    assert audio.is_cuda and audio.dtype == torch.float32 and audio.dim() == 2 and audio.stride(1) == 1
    n_streams = audio.shape[0]
    msgs = torch.empty((n_streams, n_messages, AUDIO_MESSAGE_BYTES), dtype=torch.uint8, device=audio.device)
    _b._check(_b.lib().b200_wire_audio_messages(audio.data_ptr(), audio.stride(0), n_streams, first_wire_sample,
                                                n_messages, flags, buffer_len, msgs.data_ptr(),
                                                n_messages * AUDIO_MESSAGE_BYTES, _b._stream_ptr(stream)),
              "b200_wire_audio_messages")
    return msgs
