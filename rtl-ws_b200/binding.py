"""ctypes binding of libb200sdr.so (include/b200sdr.h, include/rtlws_compat.h).

PyTorch is used for device memory and streams only; every computation happens inside the
library's CUDA kernels.  There is no fallback: if the shared library is missing or the
machine has no CUDA device, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200sdr.so")

WINDOW_RECT = 0
WINDOW_HANN = 1

CHAIN_TILE = 5120          # samples: 5 frames of 1024 = 128 audio samples at R = 10
HALF_BAND_N = 11

# every symbol include/b200sdr.h and include/rtlws_compat.h declare
EXPORTED_SYMBOLS = [
    "b200_init", "b200_last_error", "b200_launch_count", "b200_sm_count",
    "b200_spectrum_plan_create", "b200_spectrum_plan_destroy", "b200_spectrum_plan_rows", "b200_spectrum_exec",
    "b200_spectrum_exec_cs32", "b200_spectrum_exec_rf32",
    "b200_fm_history_samples", "b200_fm_history_reset", "b200_fm_history_carry", "b200_fm_exec",
    "b200_chain_exec", "b200_chain_exec_r", "b200_chain_tile_samples",
    "b200_session_create", "b200_session_create_r", "b200_session_destroy", "b200_session_reset", "b200_session_chain", "b200_session_products",
    "b200_stream_create", "b200_stream_create_r", "b200_stream_destroy", "b200_stream_set_sinks", "b200_stream_set_payload_sink", "b200_stream_push", "b200_stream_poll",
    "b200_stream_flush", "b200_stream_pending_samples",
    "b200_wire_spectrum_header", "b200_wire_spectrum_message", "b200_wire_spectrum_messages",
    "b200_wire_audio_messages", "b200_wire_audio_fragment", "b200_wire_reference_drain_index",
    "b200_host_alloc", "b200_host_free",
    "b200_fm_exec_cs32", "b200_fm_demod_create", "b200_fm_demod_destroy", "b200_fm_demod_reset", "b200_fm_demod_block",
    "b200_debug_atan2", "b200_audio_post", "b200_audio_post_out_samples", "b200_audio_resample_taps",
    "b200_shard_count", "b200_shard_stream", "b200_comm_unique_id", "b200_comm_create", "b200_comm_create_all",
    "b200_comm_destroy", "b200_comm_world", "b200_comm_rank", "b200_comm_nccl_version", "b200_comm_gather_rows",
    "b200_comm_gather_rows_all", "b200_multi_create", "b200_multi_destroy", "b200_multi_chain",
    "spectrum_alloc", "spectrum_add_cmplx_u8", "spectrum_add_cmplx_s32", "spectrum_add_real_f32", "spectrum_free",
    "cic_decimate", "halfband_decimate",
    "rf_decimator_alloc", "rf_decimator_add_callback", "rf_decimator_set_parameters",
    "rf_decimator_decimate_cmplx_u8", "rf_decimator_remove_callbacks", "rf_decimator_free",
]


class B200Error(RuntimeError):
    pass


class CmplxS32(C.Structure):
    _fields_ = [("re", C.c_int32), ("im", C.c_int32)]


class CicDelayLine(C.Structure):
    """struct cic_delay_line (resample.h:8-12)."""
    _fields_ = [("integrator_prev_out", CmplxS32), ("comb_prev_in", CmplxS32)]


RF_CALLBACK = C.CFUNCTYPE(None, C.POINTER(CmplxS32), C.c_int)
SPECTRUM_SINK = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_float))
AUDIO_SINK = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_float))
PAYLOAD_SINK = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_ubyte))

_lib = None


def lib() -> C.CDLL:
    """Load libb200sdr.so (built by __graft_entry__.build()); raises if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, f32p = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_void_p
    L.b200_init.argtypes = [i32]
    L.b200_last_error.restype = C.c_char_p
    L.b200_launch_count.restype = u64
    L.b200_sm_count.restype = i32
    L.b200_spectrum_plan_create.restype = vp
    L.b200_spectrum_plan_create.argtypes = [i32, i32, i32, i64, i32, i32]
    L.b200_spectrum_plan_destroy.argtypes = [vp]
    L.b200_spectrum_plan_rows.restype = i64
    L.b200_spectrum_plan_rows.argtypes = [vp, i64]
    for name in ("b200_spectrum_exec", "b200_spectrum_exec_cs32", "b200_spectrum_exec_rf32"):
        getattr(L, name).argtypes = [vp, vp, i64, i32, i64, f32p, f32p, vp, vp]
    L.b200_fm_history_samples.argtypes = [i32]
    L.b200_fm_history_reset.argtypes = [vp, i64, i32, i32, vp]
    L.b200_fm_history_carry.argtypes = [vp, i64, i32, i64, i32, vp]
    L.b200_fm_exec.argtypes = [vp, i64, i32, i64, i32, f32p, i64, vp, i64, vp]
    L.b200_chain_exec.argtypes = [vp, i64, i32, i64, i32, f32p, f32p, i64, vp, i32, vp]
    L.b200_chain_exec_r.argtypes = [vp, i64, i32, i64, i32, i32, f32p, f32p, i64, vp, i32, vp]
    L.b200_chain_tile_samples.restype = i64
    L.b200_chain_tile_samples.argtypes = [i32]
    L.b200_session_create.restype = vp
    L.b200_session_create.argtypes = [i32, i64]
    L.b200_session_create_r.restype = vp
    L.b200_session_create_r.argtypes = [i32, i64, i32]
    L.b200_stream_create_r.restype = vp
    L.b200_stream_create_r.argtypes = [i32, i64, i32, i32]
    L.b200_session_destroy.argtypes = [vp]
    L.b200_session_reset.argtypes = [vp]
    L.b200_session_chain.argtypes = [vp, vp, i64, i32, vp, vp]
    L.b200_session_products.argtypes = [vp, vp, i64, i32, i32, vp, vp]
    L.b200_stream_create.restype = vp
    L.b200_stream_create.argtypes = [i32, i64, i32]
    L.b200_stream_destroy.argtypes = [vp]
    L.b200_stream_set_sinks.restype = None
    L.b200_stream_set_sinks.argtypes = [vp, SPECTRUM_SINK, AUDIO_SINK, vp]
    L.b200_stream_set_payload_sink.argtypes = [vp, i32, PAYLOAD_SINK]
    L.b200_stream_push.argtypes = [vp, i32, vp, i32]
    L.b200_stream_poll.argtypes = [vp]
    L.b200_stream_flush.argtypes = [vp]
    L.b200_stream_pending_samples.restype = i64
    L.b200_stream_pending_samples.argtypes = [vp, i32]
    L.b200_wire_spectrum_header.argtypes = [C.c_char_p, i32, C.c_uint32, C.c_uint32, i32]
    L.b200_wire_spectrum_message.argtypes = [vp, i32, C.c_uint32, C.c_uint32, i32, vp, i32]
    L.b200_wire_spectrum_messages.argtypes = [vp, i64, i32, i32, vp, vp, vp, vp, i64, vp, vp]
    L.b200_wire_audio_messages.argtypes = [vp, i64, i32, i64, i32, i32, i32, vp, i64, vp]
    L.b200_wire_audio_fragment.argtypes = [i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.b200_wire_reference_drain_index.restype = i64
    L.b200_wire_reference_drain_index.argtypes = [i64, i32]
    L.b200_host_alloc.restype = vp
    L.b200_host_alloc.argtypes = [u64]
    L.b200_host_free.argtypes = [vp]
    L.b200_fm_exec_cs32.argtypes = [vp, i64, i32, i64, vp, vp, i64, vp, i64, vp, i64, i32, vp]
    L.b200_fm_demod_create.restype = vp
    L.b200_fm_demod_destroy.restype = None
    L.b200_fm_demod_destroy.argtypes = [vp]
    L.b200_fm_demod_reset.argtypes = [vp]
    L.b200_fm_demod_block.argtypes = [vp, vp, i32, vp, vp]
    L.b200_debug_atan2.argtypes = [vp, i32, vp, i32, vp]
    L.b200_audio_post.argtypes = [vp, i64, i32, i64, C.c_double, i32, vp, vp, i64, vp]
    L.b200_audio_post_out_samples.restype = i64
    L.b200_audio_post_out_samples.argtypes = [i64, i32]
    L.b200_audio_resample_taps.argtypes = [vp]
    L.b200_shard_count.argtypes = [i32, i32, i32]
    L.b200_shard_stream.argtypes = [i32, i32, i32, i32]
    L.b200_comm_unique_id.argtypes = [vp]
    L.b200_comm_create.restype = vp
    L.b200_comm_create.argtypes = [vp, i32, i32]
    L.b200_comm_create_all.argtypes = [i32, C.POINTER(vp)]
    L.b200_comm_destroy.restype = None
    L.b200_comm_destroy.argtypes = [vp]
    L.b200_comm_world.argtypes = [vp]
    L.b200_comm_rank.argtypes = [vp]
    L.b200_comm_gather_rows.argtypes = [vp, vp, i32, i32, vp, i32, vp]
    L.b200_comm_gather_rows_all.argtypes = [C.POINTER(vp), i32, C.POINTER(vp), i32, i32, vp, i32, C.POINTER(vp)]
    L.b200_multi_create.restype = vp
    L.b200_multi_create.argtypes = [i32, i32, i64, i32, i32]
    L.b200_multi_destroy.restype = None
    L.b200_multi_destroy.argtypes = [vp]
    L.b200_multi_chain.argtypes = [vp, vp, i64, vp, vp, vp]
    # reference-named interface
    L.spectrum_alloc.restype = vp
    L.spectrum_alloc.argtypes = [i32]
    L.spectrum_free.argtypes = [vp]
    for name in ("spectrum_add_cmplx_u8", "spectrum_add_cmplx_s32", "spectrum_add_real_f32"):
        getattr(L, name).argtypes = [vp, vp, vp, i32]
    L.cic_decimate.argtypes = [i32, vp, i32, vp, i32, C.POINTER(CicDelayLine)]
    L.halfband_decimate.restype = None
    L.halfband_decimate.argtypes = [vp, vp, i32, vp]
    L.rf_decimator_alloc.restype = vp
    L.rf_decimator_add_callback.argtypes = [vp, RF_CALLBACK]
    L.rf_decimator_set_parameters.argtypes = [vp, C.c_double, i32]
    L.rf_decimator_decimate_cmplx_u8.argtypes = [vp, vp, i32]
    L.rf_decimator_remove_callbacks.argtypes = [vp]
    L.rf_decimator_free.argtypes = [vp]
    _lib = L
    return L


def last_error() -> str:
    return lib().b200_last_error().decode()


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise B200Error(f"{what} -> {rc}: {last_error()}")


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise B200Error("no CUDA device: libb200sdr has no CPU fallback")
    return torch


def _stream_ptr(stream=None):
    torch = _torch()
    s = torch.cuda.current_stream() if stream is None else stream
    return C.c_void_p(s.cuda_stream)


def init(device: int = 0) -> None:
    _check(lib().b200_init(device), "b200_init")


def launch_count() -> int:
    return int(lib().b200_launch_count())


# --------------------------------------------------------------------------------------
# batched spectrum
# --------------------------------------------------------------------------------------

class SpectrumPlan:
    """b200_spectrum_plan_*: N-point frames, K accumulated per row (spectrum.c + cbb_main.c:48-59,112-128)."""

    def __init__(self, N: int = 1024, hop: int | None = None, K: int = 1, row_hop: int | None = None,
                 window: int = WINDOW_RECT, gain_db: int = 0):
        self.N, self.K = N, K
        self.hop = N if hop is None else hop
        self.row_hop = K * self.hop if row_hop is None else row_hop
        self.h = lib().b200_spectrum_plan_create(N, self.hop, K, self.row_hop, window, gain_db)
        if not self.h:
            raise B200Error(f"b200_spectrum_plan_create: {last_error()}")

    def close(self):
        if getattr(self, "h", None):
            lib().b200_spectrum_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter shutdown: the module globals may already be gone
            pass

    def rows(self, n_samples: int) -> int:
        return int(lib().b200_spectrum_plan_rows(self.h, n_samples))

    def exec(self, iq, n_rows: int | None = None, db: bool = True, power: bool = False, db_u8: bool = False,
             out: dict | None = None, stream=None) -> dict:
        """iq: cuda uint8 tensor [n_streams, n_samples, 2] (rows may be strided views).

        Returns a dict with the requested [n_streams, n_rows, N] tensors."""
        torch = _torch()
        assert iq.is_cuda and iq.dtype == torch.uint8 and iq.dim() == 3 and iq.shape[2] == 2
        assert iq.stride(2) == 1 and iq.stride(1) == 2
        n_streams, n_samples = iq.shape[0], iq.shape[1]
        if n_rows is None:
            n_rows = self.rows(n_samples)
        res = {} if out is None else out
        if db and "db" not in res:
            res["db"] = torch.empty((n_streams, n_rows, self.N), dtype=torch.float32, device=iq.device)
        if power and "power" not in res:
            res["power"] = torch.empty((n_streams, n_rows, self.N), dtype=torch.float32, device=iq.device)
        if db_u8 and "db_u8" not in res:
            res["db_u8"] = torch.empty((n_streams, n_rows, self.N), dtype=torch.uint8, device=iq.device)
        ptr = lambda k: C.c_void_p(res[k].data_ptr()) if k in res else None
        stride = iq.stride(0) if n_streams > 1 else 0
        _check(lib().b200_spectrum_exec(self.h, C.c_void_p(iq.data_ptr()), stride, n_streams, n_rows,
                                        ptr("db"), ptr("power"), ptr("db_u8"), _stream_ptr(stream)),
               "b200_spectrum_exec")
        return res

    def exec_cs32(self, x, n_rows: int | None = None, stream=None) -> dict:
        """x: cuda int32 tensor [n_streams, n_samples, 2] (spectrum_add_cmplx_s32 semantics)."""
        torch = _torch()
        assert x.is_cuda and x.dtype == torch.int32 and x.is_contiguous()
        n_streams, n_samples = x.shape[0], x.shape[1]
        n_rows = self.rows(n_samples) if n_rows is None else n_rows
        power = torch.empty((n_streams, n_rows, self.N), dtype=torch.float32, device=x.device)
        _check(lib().b200_spectrum_exec_cs32(self.h, C.c_void_p(x.data_ptr()), n_samples * 8, n_streams, n_rows,
                                             None, C.c_void_p(power.data_ptr()), None, _stream_ptr(stream)),
               "b200_spectrum_exec_cs32")
        return {"power": power}

    def exec_rf32(self, x, n_rows: int | None = None, stream=None) -> dict:
        """x: cuda float32 tensor [n_streams, n_samples] (spectrum_add_real_f32 semantics)."""
        torch = _torch()
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        n_streams, n_samples = x.shape[0], x.shape[1]
        n_rows = self.rows(n_samples) if n_rows is None else n_rows
        power = torch.empty((n_streams, n_rows, self.N), dtype=torch.float32, device=x.device)
        _check(lib().b200_spectrum_exec_rf32(self.h, C.c_void_p(x.data_ptr()), n_samples * 4, n_streams, n_rows,
                                             None, C.c_void_p(power.data_ptr()), None, _stream_ptr(stream)),
               "b200_spectrum_exec_rf32")
        return {"power": power}


# --------------------------------------------------------------------------------------
# device ring of independent streams: [history | batch] per stream
# --------------------------------------------------------------------------------------

class StreamRing:
    """Device-resident IQ of `n_streams` independent dongles laid out as the FM branch wants it:
    each row is [history (32*R samples) | batch (n_samples)], rows 16-byte aligned."""

    def __init__(self, n_streams: int, n_samples: int, R: int = 10, device="cuda"):
        torch = _torch()
        self.n_streams, self.n_samples, self.R = n_streams, n_samples, R
        self.hist = int(lib().b200_fm_history_samples(R))
        row = self.hist + n_samples
        row = (row + 7) // 8 * 8
        self.buf = torch.empty((n_streams, row, 2), dtype=torch.uint8, device=device)
        self.batch = self.buf[:, self.hist:self.hist + n_samples, :]
        self.stride_bytes = self.buf.stride(0)
        self.reset()

    def reset(self, stream=None) -> None:
        """Every stream back to its start (history = 128: the reference's zeroed delay lines)."""
        _check(lib().b200_fm_history_reset(C.c_void_p(self.batch.data_ptr()), self.stride_bytes, self.n_streams,
                                           self.R, _stream_ptr(stream)), "b200_fm_history_reset")

    def carry(self, stream=None) -> None:
        """History <- tail of the batch (call after a batch has been processed)."""
        _check(lib().b200_fm_history_carry(C.c_void_p(self.batch.data_ptr()), self.stride_bytes, self.n_streams,
                                           self.n_samples, self.R, _stream_ptr(stream)), "b200_fm_history_carry")

    def load(self, iq) -> None:
        """Copy a [n_streams, n_samples, 2] uint8 array/tensor into the batch area."""
        torch = _torch()
        t = torch.as_tensor(iq)
        self.batch.copy_(t.reshape(self.n_streams, self.n_samples, 2), non_blocking=True)


def fm_exec(ring: StreamRing, audio=None, decimated: bool = False, stream=None):
    """b200_fm_exec over a StreamRing -> (audio [n_streams, n/(4R)] f32, decimated [n_streams, n/R, 2] i32 | None)."""
    torch = _torch()
    n_audio = ring.n_samples // (4 * ring.R)
    if audio is None:
        audio = torch.empty((ring.n_streams, n_audio), dtype=torch.float32, device=ring.buf.device)
    dec = None
    if decimated:
        dec = torch.empty((ring.n_streams, ring.n_samples // ring.R, 2), dtype=torch.int32, device=ring.buf.device)
    _check(lib().b200_fm_exec(C.c_void_p(ring.batch.data_ptr()), ring.stride_bytes, ring.n_streams, ring.n_samples,
                              ring.R, C.c_void_p(audio.data_ptr()), audio.stride(0),
                              C.c_void_p(dec.data_ptr()) if dec is not None else None,
                              dec.stride(0) // 2 if dec is not None else 0, _stream_ptr(stream)), "b200_fm_exec")
    return audio, dec


def chain_exec(ring: StreamRing, gain_db: int = 0, db=None, audio=None, avg_u8=None, K_avg: int = 6, stream=None):
    """b200_chain_exec: per-frame dB spectra (N = 1024) + FM audio from one pass over the ring's batch."""
    torch = _torch()
    n_audio = ring.n_samples // (4 * ring.R)
    if db is None:
        db = torch.empty((ring.n_streams, ring.n_samples // 1024, 1024), dtype=torch.float32, device=ring.buf.device)
    if audio is None:
        audio = torch.empty((ring.n_streams, n_audio), dtype=torch.float32, device=ring.buf.device)
    _check(lib().b200_chain_exec_r(C.c_void_p(ring.batch.data_ptr()), ring.stride_bytes, ring.n_streams,
                                   ring.n_samples, ring.R, gain_db, C.c_void_p(db.data_ptr()), C.c_void_p(audio.data_ptr()),
                                   audio.stride(0), C.c_void_p(avg_u8.data_ptr()) if avg_u8 is not None else None,
                                   K_avg, _stream_ptr(stream)), "b200_chain_exec_r")
    return db, audio


FM_STATE_FLOATS = 48
FM_SKIP_STAGE2 = 1
AUDIO_DEEMPH_50US, AUDIO_DEEMPH_75US, AUDIO_RESAMPLE_48K = 1, 2, 4
AUDIO_POST_STATE_FLOATS = 64
COMM_ID_BYTES = 128


def fm_exec_cs32(dec, state, audio=None, demod: bool = False, phase: bool = False, flags: int = 0, stream=None):
    """b200_fm_exec_cs32: decimated cmplx_s32 [n_streams, n, 2] (cuda int32) + carried state
    [n_streams, FM_STATE_FLOATS] -> dict(audio, demod?, phase?).  audio_main.c:110-139."""
    torch = _torch()
    assert dec.is_cuda and dec.dtype == torch.int32 and dec.dim() == 3 and dec.shape[2] == 2 and dec.stride(1) == 2
    assert state.is_cuda and state.dtype == torch.float32 and state.shape == (dec.shape[0], FM_STATE_FLOATS)
    n_streams, n = dec.shape[0], dec.shape[1]
    res = {}
    skip = bool(flags & FM_SKIP_STAGE2)
    if not skip:
        res["audio"] = audio if audio is not None else torch.empty((n_streams, n // 4), dtype=torch.float32, device=dec.device)
    if demod:
        res["demod"] = torch.empty((n_streams, n), dtype=torch.float32, device=dec.device)
    if phase:
        res["phase"] = torch.empty((n_streams, n), dtype=torch.float32, device=dec.device)
    ptr = lambda k: C.c_void_p(res[k].data_ptr()) if k in res else None
    stride = lambda k: res[k].stride(0) if k in res else 0
    _check(lib().b200_fm_exec_cs32(C.c_void_p(dec.data_ptr()), dec.stride(0) // 2, n_streams, n,
                                   C.c_void_p(state.data_ptr()), ptr("audio"), stride("audio"), ptr("demod"),
                                   stride("demod"), ptr("phase"), stride("phase"), flags, _stream_ptr(stream)),
           "b200_fm_exec_cs32")
    return res


def debug_atan2(y, x, which: int):
    """atan2_approx of integer pairs as kernel family `which` evaluates it (b200_debug_atan2)."""
    torch = _torch()
    yx = torch.stack([torch.as_tensor(y, dtype=torch.int32), torch.as_tensor(x, dtype=torch.int32)], dim=-1).contiguous().cuda()
    out = torch.empty(yx.shape[:-1], dtype=torch.float32, device="cuda")
    _check(lib().b200_debug_atan2(C.c_void_p(yx.data_ptr()), out.numel(), C.c_void_p(out.data_ptr()), which,
                                  _stream_ptr()), "b200_debug_atan2")
    return out.cpu().numpy()


class FmDemod:
    """b200_fm_demod_*: the body of an rf_decimator_callback -- host cmplx_s32 blocks in, host audio out."""

    def __init__(self):
        _torch()
        self.h = lib().b200_fm_demod_create()
        if not self.h:
            raise B200Error(f"b200_fm_demod_create: {last_error()}")

    def block(self, signal: np.ndarray, want_audio: bool = True, want_demod: bool = False):
        signal = np.ascontiguousarray(signal, dtype=np.int32).reshape(-1, 2)
        n = len(signal)
        audio = np.empty(n // 4, dtype=np.float32) if want_audio else None
        demod = np.empty(n, dtype=np.float32) if want_demod else None
        _check(lib().b200_fm_demod_block(self.h, signal.ctypes.data, n, audio.ctypes.data if want_audio else None,
                                         demod.ctypes.data if want_demod else None), "b200_fm_demod_block")
        return audio, demod

    def reset(self):
        _check(lib().b200_fm_demod_reset(self.h), "b200_fm_demod_reset")

    def close(self):
        if getattr(self, "h", None):
            lib().b200_fm_demod_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def audio_post(audio, state, flags: int, rate_hz: float = 51200.0, out=None, stream=None):
    """b200_audio_post: opt-in de-emphasis / 15:16 resampling of [n_streams, n] f32 audio rows (extensions,
    not in the reference).  state: cuda float32 [n_streams, AUDIO_POST_STATE_FLOATS], zero = stream start."""
    torch = _torch()
    assert audio.is_cuda and audio.dtype == torch.float32 and audio.dim() == 2 and audio.stride(1) == 1
    n_streams, n = audio.shape
    n_out = int(lib().b200_audio_post_out_samples(n, flags))
    if n_out < 0:
        raise B200Error("audio_post: the resampler takes multiples of 16 samples")
    if out is None:
        out = torch.empty((n_streams, n_out), dtype=torch.float32, device=audio.device)
    _check(lib().b200_audio_post(C.c_void_p(audio.data_ptr()), audio.stride(0), n_streams, n, rate_hz, flags,
                                 C.c_void_p(state.data_ptr()), C.c_void_p(out.data_ptr()), out.stride(0),
                                 _stream_ptr(stream)), "b200_audio_post")
    return out


def resample_taps() -> np.ndarray:
    h = np.empty(240, dtype=np.float32)
    assert lib().b200_audio_resample_taps(h.ctypes.data) == 240
    return h


class Comm:
    """b200_comm_*: the NCCL communicator behind the one exchange of the path (gather of per-stream rows).
    One process per GPU: `Comm.unique_id()` on rank 0, distribute the 128 bytes, `Comm(id, world, rank)`."""

    def __init__(self, id128: bytes, world: int, rank: int):
        _torch()
        assert len(id128) == COMM_ID_BYTES
        self.world, self.rank = world, rank
        buf = (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(id128)
        self.h = lib().b200_comm_create(buf, world, rank)
        if not self.h:
            raise B200Error(f"b200_comm_create: {last_error()}")

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_ubyte * COMM_ID_BYTES)()
        _check(lib().b200_comm_unique_id(buf), "b200_comm_unique_id")
        return bytes(buf)

    def gather_rows(self, send, n_streams_total: int, recv=None, root: int = 0, stream=None):
        """send: cuda tensor [n_local, row...] (contiguous); recv (root): [n_streams_total, row...]."""
        row_bytes = int(send[0].numel() * send.element_size()) if send.shape[0] else int(recv[0].numel() * recv.element_size())
        _check(lib().b200_comm_gather_rows(self.h, C.c_void_p(send.data_ptr()), n_streams_total, row_bytes,
                                           C.c_void_p(recv.data_ptr()) if recv is not None else None, root,
                                           _stream_ptr(stream)), "b200_comm_gather_rows")

    def close(self):
        if getattr(self, "h", None):
            lib().b200_comm_destroy(self.h)
            self.h = None


class Multi:
    """b200_multi_*: one process, G GPUs, host arrays in global stream order (stream s on device s mod G)."""

    def __init__(self, n_gpus: int, n_streams: int, max_samples: int, gain_db: int = 0, K_avg: int = 6):
        _torch()
        self.n_gpus, self.n_streams = n_gpus, n_streams
        self.h = lib().b200_multi_create(n_gpus, n_streams, max_samples, gain_db, K_avg)
        if not self.h:
            raise B200Error(f"b200_multi_create: {last_error()}")

    def chain(self, h_iq: np.ndarray, n_samples: int, want_db: bool = True, want_audio: bool = True, want_avg: bool = True):
        h_iq = np.ascontiguousarray(h_iq, dtype=np.uint8)
        db = np.empty((self.n_streams, n_samples // 1024, 1024), np.float32) if want_db else None
        audio = np.empty((self.n_streams, n_samples // 40), np.float32) if want_audio else None
        avg = np.empty((self.n_streams, 1024), np.uint8) if want_avg else None
        p = lambda a: a.ctypes.data if a is not None else None
        _check(lib().b200_multi_chain(self.h, h_iq.ctypes.data, n_samples, p(db), p(audio), p(avg)), "b200_multi_chain")
        return db, audio, avg

    def close(self):
        if getattr(self, "h", None):
            lib().b200_multi_destroy(self.h)
            self.h = None


def shard_count(n_streams: int, world: int, rank: int) -> int:
    return int(lib().b200_shard_count(n_streams, world, rank))


def shard_stream(n_streams: int, world: int, rank: int, i: int) -> int:
    return int(lib().b200_shard_stream(n_streams, world, rank, i))


class Session:
    """b200_session_*: the host-buffer entry point (PCIe copies inside)."""

    def __init__(self, n_streams: int, max_samples: int, R: int = 10):
        _torch()
        self.n_streams, self.max_samples, self.R = n_streams, max_samples, R
        self.h = lib().b200_session_create_r(n_streams, max_samples, R)
        if not self.h:
            raise B200Error(f"b200_session_create_r: {last_error()}")

    def close(self):
        if getattr(self, "h", None):
            lib().b200_session_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter shutdown: the module globals may already be gone
            pass

    def reset(self):
        lib().b200_session_reset(self.h)

    def chain(self, h_iq, n_samples: int, h_db=None, h_audio=None, gain_db: int = 0) -> None:
        """h_iq / h_db / h_audio: host tensors or numpy arrays (pinned for asynchronous copies)."""
        def hp(x):
            if x is None:
                return None
            return C.c_void_p(x.data_ptr() if hasattr(x, "data_ptr") else x.ctypes.data)
        _check(lib().b200_session_chain(self.h, hp(h_iq), n_samples, gain_db, hp(h_db), hp(h_audio)),
               "b200_session_chain")


    def products(self, h_iq, n_samples: int, h_audio, h_avg_u8, gain_db: int = 0, K_avg: int = 6) -> None:
        """b200_session_products: audio + the K-frame averaged payload bytes per stream (the reference's own products)."""
        def hp(x):
            if x is None:
                return None
            return C.c_void_p(x.data_ptr() if hasattr(x, "data_ptr") else x.ctypes.data)
        _check(lib().b200_session_products(self.h, hp(h_iq), n_samples, gain_db, K_avg, hp(h_audio), hp(h_avg_u8)),
               "b200_session_products")


class PushStream:
    """b200_stream_*: signal_source-style pushes in, spectra and audio out through sinks."""

    def __init__(self, n_streams: int, batch_samples: int, gain_db: int = 0, R: int = 10, payload_K: int = 0,
                 frames: bool = True):
        """payload_K > 0: also collect the K-frame averaged u8 payload of every batch (self.payloads);
        frames=False: no per-frame dB sink (the rows stay on the device)."""
        _torch()
        self.n_streams, self.batch, self.R = n_streams, batch_samples, R
        self.h = lib().b200_stream_create_r(n_streams, batch_samples, gain_db, R)
        if not self.h:
            raise B200Error(f"b200_stream_create_r: {last_error()}")
        self.payloads = [[] for _ in range(n_streams)]     # (first_frame, K, [1024] u8 copy)
        self._frames = frames
        self._payload_K = payload_K
        self.spectra = [[] for _ in range(n_streams)]      # (first_frame, [n_frames, 1024] copy)
        self.audio = [[] for _ in range(n_streams)]        # (first_sample, [n] copy)

        def on_spectrum(user, stream, first, n, ptr):
            self.spectra[stream].append((first, np.ctypeslib.as_array(ptr, shape=(n, 1024)).copy()))

        def on_audio(user, stream, first, n, ptr):
            self.audio[stream].append((first, np.ctypeslib.as_array(ptr, shape=(n,)).copy()))

        def on_payload(user, stream, first, k, ptr):
            self.payloads[stream].append((first, k, np.ctypeslib.as_array(ptr, shape=(1024,)).copy()))

        self._on_spectrum = SPECTRUM_SINK(on_spectrum)
        self._cb = (self._on_spectrum if self._frames else SPECTRUM_SINK(), AUDIO_SINK(on_audio), PAYLOAD_SINK(on_payload))
        lib().b200_stream_set_sinks(self.h, self._cb[0], self._cb[1], None)
        if self._payload_K > 0:
            _check(lib().b200_stream_set_payload_sink(self.h, self._payload_K, self._cb[2]), "b200_stream_set_payload_sink")

    def enable_frames(self) -> None:
        """Install the per-frame dB sink on a stream that was created without one (frames=False): the library
        allocates the row buffers now; batches submitted before this call deliver no rows."""
        self._frames = True
        self._cb = (self._on_spectrum, self._cb[1], self._cb[2])
        lib().b200_stream_set_sinks(self.h, self._cb[0], self._cb[1], None)

    def push(self, stream: int, samples: np.ndarray) -> None:
        samples = np.ascontiguousarray(samples, dtype=np.uint8)
        _check(lib().b200_stream_push(self.h, stream, samples.ctypes.data, samples.size // 2), "b200_stream_push")

    def poll(self) -> None:
        _check(lib().b200_stream_poll(self.h), "b200_stream_poll")

    def flush(self) -> None:
        _check(lib().b200_stream_flush(self.h), "b200_stream_flush")

    def pending(self, stream: int) -> int:
        return int(lib().b200_stream_pending_samples(self.h, stream))

    def close(self):
        if getattr(self, "h", None):
            lib().b200_stream_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter shutdown: the module globals may already be gone
            pass


# --------------------------------------------------------------------------------------
# the reference's interface, same names (numpy in / numpy out, synchronous)
# --------------------------------------------------------------------------------------

def spectrum_alloc(N: int):
    h = lib().spectrum_alloc(N)
    if not h:
        raise B200Error(f"spectrum_alloc({N}) failed")
    return h


def spectrum_free(s) -> None:
    lib().spectrum_free(s)


def spectrum_add_cmplx_u8(s, src: np.ndarray, power_spectrum: np.ndarray, length: int) -> int:
    src = np.ascontiguousarray(src, dtype=np.uint8)
    assert power_spectrum.dtype == np.float64 and power_spectrum.flags.c_contiguous
    return lib().spectrum_add_cmplx_u8(s, src.ctypes.data, power_spectrum.ctypes.data, length)


def spectrum_add_cmplx_s32(s, src: np.ndarray, power_spectrum: np.ndarray, length: int) -> int:
    src = np.ascontiguousarray(src, dtype=np.int32)
    assert power_spectrum.dtype == np.float64 and power_spectrum.flags.c_contiguous
    return lib().spectrum_add_cmplx_s32(s, src.ctypes.data, power_spectrum.ctypes.data, length)


def spectrum_add_real_f32(s, src: np.ndarray, power_spectrum: np.ndarray, length: int) -> int:
    src = np.ascontiguousarray(src, dtype=np.float32)
    assert power_spectrum.dtype == np.float64 and power_spectrum.flags.c_contiguous
    return lib().spectrum_add_real_f32(s, src.ctypes.data, power_spectrum.ctypes.data, length)


def cic_decimate(R: int, src: np.ndarray, src_len: int, dst: np.ndarray, dst_len: int, delay: CicDelayLine) -> int:
    src = np.ascontiguousarray(src, dtype=np.uint8)
    assert dst.dtype == np.int32 and dst.flags.c_contiguous
    return lib().cic_decimate(R, src.ctypes.data, src_len, dst.ctypes.data, dst_len, C.byref(delay))


def halfband_decimate(inp: np.ndarray, out: np.ndarray, output_len: int, delay: np.ndarray) -> None:
    inp = np.ascontiguousarray(inp, dtype=np.float32)
    assert out.dtype == np.float32 and delay.dtype == np.float32 and len(delay) == HALF_BAND_N - 1
    lib().halfband_decimate(inp.ctypes.data, out.ctypes.data, output_len, delay.ctypes.data)


class RfDecimator:
    """rf_decimator_* (rf_decimator.h:11-21): re-blocking to 100 ms, CIC on the GPU, host callbacks."""

    def __init__(self):
        self.h = lib().rf_decimator_alloc()
        self._keep = []

    def add_callback(self, fn) -> None:
        def tramp(ptr, n):
            arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int32)), shape=(n, 2))
            fn(arr, n)
        cb = RF_CALLBACK(tramp)
        self._keep.append(cb)
        lib().rf_decimator_add_callback(self.h, cb)

    def set_parameters(self, sample_rate: float, down_factor: int) -> int:
        return lib().rf_decimator_set_parameters(self.h, sample_rate, down_factor)

    def decimate_cmplx_u8(self, signal: np.ndarray, length: int | None = None) -> int:
        signal = np.ascontiguousarray(signal, dtype=np.uint8)
        n = signal.size // 2 if length is None else length
        return lib().rf_decimator_decimate_cmplx_u8(self.h, signal.ctypes.data, n)

    def remove_callbacks(self) -> None:
        lib().rf_decimator_remove_callbacks(self.h)
        self._keep.clear()

    def free(self) -> None:
        if self.h:
            lib().rf_decimator_free(self.h)
            self.h = None
