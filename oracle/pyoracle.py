"""ctypes bindings for the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Two libraries live behind this module:

* ``liboracle.so``  -- the explicit-state C restatement (oracle/oracle.c, "port").
* ``_ref/libref_rtlws.so`` -- the UNMODIFIED reference sources compiled in place
  (oracle/Makefile, "reference"); present wherever the build ran with /root/reference
  mounted, and shipped prebuilt to the GPU box.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_rtlws.so")
REF_BENCH = os.path.join(HERE, "_ref", "ref_bench")
# the reference's own glue (cbb_main.c, audio_main.c, signal_source.c) linked against the PRODUCT
# library instead of the reference's spectrum.o / resample.o / rf_decimator.o
DROPIN_SO = os.path.join(HERE, "_ref", "libdropin_rtlws.so")
DROPIN_AUDIO_SO = os.path.join(HERE, "_ref", "libdropin_audio_rtlws.so")   # ... and audio_main.o replaced by libb200audio.so
REPLAY_SO = os.path.join(HERE, "_ref", "libreplay_rtlws.so")     # reference signal_source.c over libb200replay.so
PRODUCT_SO = os.path.join(os.path.dirname(HERE), "rtl-ws_b200", "libb200sdr.so")

HALF_BAND_N = 11

_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(target: str = "all") -> None:
    """Run oracle/Makefile (port always; ref only where /root/reference is mounted)."""
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


# --------------------------------------------------------------------------------------
# port
# --------------------------------------------------------------------------------------

class CicState(C.Structure):
    _fields_ = [("integrator_prev_out", C.c_int32 * 2), ("comb_prev_in", C.c_int32 * 2)]


class FmState(C.Structure):
    _fields_ = [("prev_sample", C.c_float),
                ("delay_line_1", C.c_float * (HALF_BAND_N - 1)),
                ("delay_line_2", C.c_float * (HALF_BAND_N - 1))]


_port = None


def port():
    global _port
    if _port is not None:
        return _port
    if not os.path.exists(PORT_SO):
        build("port")
    lib = C.CDLL(PORT_SO)
    lib.orc_fft_plan_create.restype = C.c_void_p
    lib.orc_fft_plan_create.argtypes = [C.c_int]
    lib.orc_fft_plan_destroy.argtypes = [C.c_void_p]
    lib.orc_fft_execute.argtypes = [C.c_void_p, _f64p, _f64p]
    lib.orc_dft_naive.argtypes = [C.c_int, _f64p, _f64p]
    lib.orc_spectrum_alloc.restype = C.c_void_p
    lib.orc_spectrum_alloc.argtypes = [C.c_int]
    lib.orc_spectrum_free.argtypes = [C.c_void_p]
    lib.orc_spectrum_set_window.argtypes = [C.c_void_p, C.c_void_p]
    lib.orc_spectrum_add_cmplx_u8.argtypes = [C.c_void_p, _u8p, _f64p, C.c_int]
    lib.orc_spectrum_add_cmplx_s32.argtypes = [C.c_void_p, _i32p, _f64p, C.c_int]
    lib.orc_spectrum_add_real_f32.argtypes = [C.c_void_p, _f32p, _f64p, C.c_int]
    lib.orc_spectrum_rows_cmplx_u8.argtypes = [C.c_void_p, _u8p, C.c_int64, C.c_int, C.c_int,
                                               C.c_int64, _f64p, C.c_int64]
    lib.orc_db_payload.argtypes = [_f64p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.orc_cic_decimate.argtypes = [C.c_int, _u8p, C.c_int, _i32p, C.c_int, C.POINTER(CicState)]
    lib.orc_halfband_decimate.argtypes = [_f32p, _f32p, C.c_int, _f32p]
    lib.orc_atan2_approx.restype = C.c_float
    lib.orc_atan2_approx.argtypes = [C.c_float, C.c_float]
    lib.orc_fm_demodulate.argtypes = [_i32p, C.c_int, C.POINTER(FmState), _f32p, _f32p, _f32p]
    lib.orc_chain_create.restype = C.c_void_p
    lib.orc_chain_create.argtypes = [C.c_double, C.c_int]
    lib.orc_chain_free.argtypes = [C.c_void_p]
    lib.orc_chain_block_in.argtypes = [C.c_void_p]
    lib.orc_chain_block_out.argtypes = [C.c_void_p]
    lib.orc_chain_push.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_int64,
                                   C.c_void_p, C.POINTER(C.c_int64), C.c_int64]
    lib.orc_deemphasis.argtypes = [_f32p, C.c_int, C.c_float, _f32p, _f32p]
    lib.orc_deemphasis_alpha.restype = C.c_float
    lib.orc_deemphasis_alpha.argtypes = [C.c_double, C.c_double]
    lib.orc_resample_taps.argtypes = [_f32p]
    lib.orc_resample_15_16.argtypes = [_f32p, C.c_int, _f32p, _f32p]
    _port = lib
    return lib


def fft(x: np.ndarray) -> np.ndarray:
    """Unnormalised forward DFT of a complex128 vector (power-of-two length)."""
    x = np.ascontiguousarray(x, dtype=np.complex128)
    lib = port()
    plan = lib.orc_fft_plan_create(len(x))
    if not plan:
        raise ValueError("length must be a power of two")
    out = np.empty(len(x), dtype=np.complex128)
    lib.orc_fft_execute(plan, x.view(np.float64), out.view(np.float64))
    lib.orc_fft_plan_destroy(plan)
    return out


def dft_naive(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.complex128)
    out = np.empty(len(x), dtype=np.complex128)
    port().orc_dft_naive(len(x), x.view(np.float64), out.view(np.float64))
    return out


class Spectrum:
    """orc_spectrum_* (spectrum.c restated)."""

    def __init__(self, N: int, window: np.ndarray | None = None):
        self.lib = port()
        self.N = N
        self.h = self.lib.orc_spectrum_alloc(N)
        if not self.h:
            raise ValueError("bad N")
        if window is not None:
            w = np.ascontiguousarray(window, dtype=np.float64)
            assert len(w) == N
            self.lib.orc_spectrum_set_window(self.h, w.ctypes.data_as(C.c_void_p))

    def close(self):
        if self.h:
            self.lib.orc_spectrum_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def add_cmplx_u8(self, iq: np.ndarray, ps: np.ndarray, length: int | None = None) -> int:
        iq = np.ascontiguousarray(iq, dtype=np.uint8).reshape(-1)
        return self.lib.orc_spectrum_add_cmplx_u8(self.h, iq, ps, len(iq) // 2 if length is None else length)

    def add_cmplx_s32(self, iq: np.ndarray, ps: np.ndarray, length: int | None = None) -> int:
        iq = np.ascontiguousarray(iq, dtype=np.int32).reshape(-1)
        return self.lib.orc_spectrum_add_cmplx_s32(self.h, iq, ps, len(iq) // 2 if length is None else length)

    def add_real_f32(self, x: np.ndarray, ps: np.ndarray, length: int | None = None) -> int:
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
        return self.lib.orc_spectrum_add_real_f32(self.h, x, ps, len(x) if length is None else length)

    def rows(self, iq: np.ndarray, hop: int | None = None, K: int = 1, row_hop: int | None = None,
             n_rows: int | None = None) -> np.ndarray:
        """[n_rows, N] float64: each row = K frames accumulated into a zeroed row."""
        iq = np.ascontiguousarray(iq, dtype=np.uint8).reshape(-1)
        n = len(iq) // 2
        hop = self.N if hop is None else hop
        row_hop = K * hop if row_hop is None else row_hop
        if n_rows is None:
            span = (K - 1) * hop + self.N
            n_rows = 0 if n < span else (n - span) // row_hop + 1
        out = np.zeros((n_rows, self.N), dtype=np.float64)
        r = self.lib.orc_spectrum_rows_cmplx_u8(self.h, iq, n, hop, K, row_hop, out.reshape(-1), n_rows)
        if r:
            raise ValueError("rows: capture too short")
        return out


def db_payload(ps: np.ndarray, count: int, gain_db: int):
    """cbb_main.c:106-135 -> (uint8 payload, float64 dB before truncation)."""
    ps = np.ascontiguousarray(ps, dtype=np.float64).reshape(-1)
    out = np.empty(len(ps), dtype=np.uint8)
    dbf = np.empty(len(ps), dtype=np.float64)
    port().orc_db_payload(ps, len(ps), count, gain_db, out.ctypes.data_as(C.c_void_p),
                          dbf.ctypes.data_as(C.c_void_p))
    return out, dbf


def cic_decimate(R: int, iq: np.ndarray, state: CicState | None = None, dst_len: int | None = None):
    """-> (return code, int32 [dst_len, 2], state)."""
    iq = np.ascontiguousarray(iq, dtype=np.uint8).reshape(-1)
    n = len(iq) // 2
    dst_len = n // R if dst_len is None else dst_len
    dst = np.zeros((max(dst_len, 1), 2), dtype=np.int32)
    st = state if state is not None else CicState()
    r = port().orc_cic_decimate(R, iq, n, dst.reshape(-1), dst_len, C.byref(st))
    return r, dst[:dst_len], st


def halfband_decimate(x: np.ndarray, delay: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    assert delay.dtype == np.float32 and len(delay) == HALF_BAND_N - 1
    out = np.empty(len(x) // 2, dtype=np.float32)
    port().orc_halfband_decimate(x, out, len(out), delay)
    return out


def atan2_approx(y, x) -> np.ndarray:
    y = np.asarray(y, dtype=np.float32)
    x = np.asarray(x, dtype=np.float32)
    f = port().orc_atan2_approx
    out = np.empty(np.broadcast(y, x).shape, dtype=np.float32)
    yb, xb = np.broadcast_arrays(y, x)
    flat = out.reshape(-1)
    for i, (a, b) in enumerate(zip(yb.reshape(-1), xb.reshape(-1))):
        flat[i] = f(float(a), float(b))
    return out


def fm_demodulate(signal: np.ndarray, state: FmState | None = None):
    """audio_main.c:110-139 on one block -> (demod, work, audio, state)."""
    signal = np.ascontiguousarray(signal, dtype=np.int32).reshape(-1)
    n = len(signal) // 2
    st = state if state is not None else FmState()
    demod = np.empty(n, dtype=np.float32)
    work = np.empty(n // 2, dtype=np.float32)
    audio = np.empty(n // 4, dtype=np.float32)
    port().orc_fm_demodulate(signal, n, C.byref(st), demod, work, audio)
    return demod, work, audio, st


def deemphasis(x: np.ndarray, rate_hz: float, tau_s: float, state: np.ndarray | None = None):
    """Definition of the product's opt-in de-emphasis (not in the reference) -> (y, state[1])."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    st = np.zeros(1, np.float32) if state is None else state
    y = np.empty_like(x)
    port().orc_deemphasis(x, len(x), port().orc_deemphasis_alpha(rate_hz, tau_s), st, y)
    return y, st


def resample_taps() -> np.ndarray:
    h = np.empty(240, dtype=np.float32)
    port().orc_resample_taps(h)
    return h


def resample_15_16(x: np.ndarray, hist: np.ndarray | None = None):
    """Definition of the product's opt-in 51.2 -> 48 kHz resampler (not in the reference) -> (y, hist[15])."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    h = np.zeros(15, np.float32) if hist is None else hist
    y = np.empty(len(x) // 16 * 15, dtype=np.float32)
    n = port().orc_resample_15_16(x, len(x), h, y)
    if n < 0:
        raise ValueError("resample_15_16 takes multiples of 16 samples")
    return y, h


class Chain:
    """One explicit-state stream: rf_decimator re-blocking -> CIC -> FM demodulation."""

    def __init__(self, sample_rate: float = 2048000.0, down_factor: int = 10):
        self.lib = port()
        self.h = self.lib.orc_chain_create(sample_rate, down_factor)
        if not self.h:
            raise ValueError("bad parameters")
        self.block_in = self.lib.orc_chain_block_in(self.h)
        self.block_out = self.lib.orc_chain_block_out(self.h)
        self.R = down_factor

    def close(self):
        if self.h:
            self.lib.orc_chain_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def push(self, iq: np.ndarray):
        """-> (decimated int32 [m, 2], audio float32 [m/4]) produced by this push."""
        iq = np.ascontiguousarray(iq, dtype=np.uint8).reshape(-1)
        n = len(iq) // 2
        max_blocks = n // self.block_in + 2
        dec = np.empty((max_blocks * self.block_out, 2), dtype=np.int32)
        audio = np.empty(max_blocks * (self.block_out // 4), dtype=np.float32)
        n_dec = C.c_int64(0)
        n_audio = C.c_int64(0)
        r = self.lib.orc_chain_push(self.h, iq, n, dec.ctypes.data_as(C.c_void_p), C.byref(n_dec), len(dec),
                                    audio.ctypes.data_as(C.c_void_p), C.byref(n_audio), len(audio))
        if r:
            raise RuntimeError(f"orc_chain_push -> {r}")
        return dec[:n_dec.value].copy(), audio[:n_audio.value].copy()


def chain_run(iq: np.ndarray, chunk: int = 131072, sample_rate: float = 2048000.0, down_factor: int = 10):
    """Whole capture through a fresh Chain in `chunk`-sample pushes."""
    iq = np.ascontiguousarray(iq, dtype=np.uint8).reshape(-1, 2)
    ch = Chain(sample_rate, down_factor)
    decs, auds = [], []
    for pos in range(0, len(iq), chunk):
        d, a = ch.push(iq[pos:pos + chunk])
        decs.append(d)
        auds.append(a)
    ch.close()
    return np.concatenate(decs) if decs else np.zeros((0, 2), np.int32), \
        np.concatenate(auds) if auds else np.zeros(0, np.float32)


# --------------------------------------------------------------------------------------
# unmodified reference (fresh statics per instance)
# --------------------------------------------------------------------------------------

class Ref:
    """A private copy of libref_rtlws.so: audio_main.c / cbb_main.c statics start from zero."""

    SO = REF_SO

    def __init__(self):
        if not os.path.exists(self.SO):
            raise FileNotFoundError(self.SO)
        self._preload()
        fd, self._tmp = tempfile.mkstemp(prefix="libref_rtlws_", suffix=".so")
        os.close(fd)
        shutil.copyfile(self.SO, self._tmp)
        lib = C.CDLL(self._tmp)
        os.unlink(self._tmp)
        self.lib = lib
        lib.spectrum_alloc.restype = C.c_void_p
        lib.spectrum_alloc.argtypes = [C.c_int]
        lib.spectrum_free.argtypes = [C.c_void_p]
        lib.spectrum_add_cmplx_u8.argtypes = [C.c_void_p, _u8p, _f64p, C.c_int]
        lib.spectrum_add_cmplx_s32.argtypes = [C.c_void_p, _i32p, _f64p, C.c_int]
        lib.spectrum_add_real_f32.argtypes = [C.c_void_p, _f32p, _f64p, C.c_int]
        lib.cic_decimate.argtypes = [C.c_int, _u8p, C.c_int, _i32p, C.c_int, C.POINTER(CicState)]
        lib.halfband_decimate.argtypes = [_f32p, _f32p, C.c_int, _f32p]
        lib.ref_atan2_approx.restype = C.c_float
        lib.ref_atan2_approx.argtypes = [C.c_float, C.c_float]
        lib.ref_fm_open.argtypes = [C.c_double, C.c_int]
        lib.ref_fm_set_outputs.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        lib.ref_fm_push.argtypes = [_u8p, C.c_int64, C.c_int]
        lib.ref_fm_n_decimated.restype = C.c_int64
        lib.ref_fm_n_audio.restype = C.c_int64
        lib.ref_fm_demodulate_block.argtypes = [_i32p, C.c_int, _f32p]
        lib.ref_fm_copy_demod.argtypes = [_f32p, C.c_int]
        lib.ref_cbb_run.argtypes = [_u8p, C.c_int64, C.c_int, _u8p, _f64p, _i32p, C.c_int,
                                    C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        lib.ref_ws_set_log.argtypes = [_u8p, C.c_int64, _i32p, C.c_int]
        lib.ref_ws_n_bytes.restype = C.c_int64
        lib.ref_ws_command.argtypes = [C.c_char_p]
        lib.ref_ws_queue_command.argtypes = [C.c_char_p]
        lib.ref_ws_pump.argtypes = [C.c_int]
        lib.ref_ws_attach.argtypes = [C.c_int]
        lib.rf_decimator_alloc.restype = C.c_void_p
        lib.rf_decimator_set_parameters.argtypes = [C.c_void_p, C.c_double, C.c_int]
        lib.rf_decimator_decimate_cmplx_u8.argtypes = [C.c_void_p, _u8p, C.c_int]
        lib.rf_decimator_free.argtypes = [C.c_void_p]

    def _preload(self):
        pass

    # spectrum.c ------------------------------------------------------------------
    def spectrum_rows(self, iq: np.ndarray, N: int, hop: int | None = None, K: int = 1,
                      row_hop: int | None = None, n_rows: int | None = None) -> np.ndarray:
        iq = np.ascontiguousarray(iq, dtype=np.uint8).reshape(-1)
        n = len(iq) // 2
        hop = N if hop is None else hop
        row_hop = K * hop if row_hop is None else row_hop
        if n_rows is None:
            span = (K - 1) * hop + N
            n_rows = 0 if n < span else (n - span) // row_hop + 1
        s = self.lib.spectrum_alloc(N)
        out = np.zeros((n_rows, N), dtype=np.float64)
        for r in range(n_rows):
            for j in range(K):
                st = r * row_hop + j * hop
                rc = self.lib.spectrum_add_cmplx_u8(s, iq[2 * st:2 * (st + N)], out[r], N)
                assert rc == 0
        self.lib.spectrum_free(s)
        return out

    # resample.c ------------------------------------------------------------------
    def cic_decimate(self, R: int, iq: np.ndarray, state: CicState | None = None, dst_len: int | None = None):
        iq = np.ascontiguousarray(iq, dtype=np.uint8).reshape(-1)
        n = len(iq) // 2
        dst_len = n // R if dst_len is None else dst_len
        dst = np.zeros((max(dst_len, 1), 2), dtype=np.int32)
        st = state if state is not None else CicState()
        r = self.lib.cic_decimate(R, iq, n, dst.reshape(-1), dst_len, C.byref(st))
        return r, dst[:dst_len], st

    def halfband_decimate(self, x: np.ndarray, delay: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty(len(x) // 2, dtype=np.float32)
        self.lib.halfband_decimate(x, out, len(out), delay)
        return out

    def atan2_grid(self, ys: np.ndarray, xs: np.ndarray) -> np.ndarray:
        f = self.lib.ref_atan2_approx
        out = np.empty((len(ys), len(xs)), dtype=np.float32)
        for i, y in enumerate(ys):
            for j, x in enumerate(xs):
                out[i, j] = f(float(y), float(x))
        return out

    # rf_decimator.c + audio_main.c -------------------------------------------------
    def fm_chain(self, iq: np.ndarray, chunk: int = 131072, sample_rate: float = 2048000.0, down_factor: int = 10):
        """-> (decimated int32 [m, 2], audio float32).  One call per Ref instance."""
        iq = np.ascontiguousarray(iq, dtype=np.uint8).reshape(-1)
        n = len(iq) // 2
        rc = self.lib.ref_fm_open(sample_rate, down_factor)
        if rc:
            raise RuntimeError(f"ref_fm_open -> {rc}")
        dec = np.zeros((n // down_factor + 16, 2), dtype=np.int32)
        audio = np.zeros(n // (4 * down_factor) + 16, dtype=np.float32)
        self.lib.ref_fm_set_outputs(dec.ctypes.data_as(C.c_void_p), len(dec),
                                    audio.ctypes.data_as(C.c_void_p), len(audio))
        rc = self.lib.ref_fm_push(iq, n, chunk)
        if rc:
            raise RuntimeError(f"ref_fm_push -> {rc}")
        nd, na = self.lib.ref_fm_n_decimated(), self.lib.ref_fm_n_audio()
        self.lib.ref_fm_close()
        return dec[:nd].copy(), audio[:na].copy()

    def fm_demodulate_block(self, signal: np.ndarray):
        """audio_fm_demodulator on one block -> (demod after limiter, audio)."""
        signal = np.ascontiguousarray(signal, dtype=np.int32).reshape(-1)
        n = len(signal) // 2
        audio = np.empty(n // 4, dtype=np.float32)
        got = self.lib.ref_fm_demodulate_block(signal, n, audio)
        assert got == n // 4
        demod = np.empty(n, dtype=np.float32)
        self.lib.ref_fm_copy_demod(demod, n)
        return demod, audio

    # the whole driver (cbb_main.c + signal_source.c + synthetic sensor) ----------------
    def cbb_run(self, iq: np.ndarray, gain_db: int = 0, max_spectra: int = 256):
        """-> dict(payload u8 [k,1024], power f64 [k,1024], count [k], decimated, audio)."""
        iq = np.ascontiguousarray(iq, dtype=np.uint8).reshape(-1)
        n = len(iq) // 2
        payload = np.zeros((max_spectra, 1024), dtype=np.uint8)
        power = np.zeros((max_spectra, 1024), dtype=np.float64)
        count = np.zeros(max_spectra, dtype=np.int32)
        dec = np.zeros((n // 10 + 16, 2), dtype=np.int32)
        audio = np.zeros(n // 40 + 16, dtype=np.float32)
        k = self.lib.ref_cbb_run(iq, len(iq), gain_db, payload.reshape(-1), power.reshape(-1), count, max_spectra,
                                 dec.ctypes.data_as(C.c_void_p), len(dec),
                                 audio.ctypes.data_as(C.c_void_p), len(audio))
        nd, na = self.lib.ref_fm_n_decimated(), self.lib.ref_fm_n_audio()
        return dict(payload=payload[:k].copy(), power=power[:k].copy(), count=count[:k].copy(),
                    decimated=dec[:nd].copy(), audio=audio[:na].copy())


    # main.c's websocket callback on top of the whole driver -------------------------------
    def ws_run(self, iq: np.ndarray, commands=("start",), max_bytes: int | None = None):
        """Replay `iq` through the unmodified driver with the unmodified main.c callback as the
        consumer (pumped after every USB buffer).  The client connects at the first poll and sends
        `commands` as text messages (main.c:139-176).  -> list of (write_mode, bytes) in
        the order main.c handed them to lws_write."""
        iq = np.ascontiguousarray(iq, dtype=np.uint8).reshape(-1)
        n = len(iq) // 2
        cap = max_bytes if max_bytes is not None else n // 40 * 4 * 2 + (n // 100000 + 16) * 2048 + (1 << 16)
        log = np.zeros(cap, dtype=np.uint8)
        max_rec = cap // 512 + 64
        rec = np.zeros((max_rec, 3), dtype=np.int32)
        self.lib.ref_ws_set_log(log, cap, rec.reshape(-1), max_rec)
        self.lib.ref_ws_attach(1)
        for c in commands:
            if self.lib.ref_ws_queue_command(c.encode()):
                raise ValueError(f"cannot queue command {c!r}")
        dummy_u8 = np.zeros(1024, dtype=np.uint8)
        dummy_f64 = np.zeros(1024, dtype=np.float64)
        dummy_i32 = np.zeros(1, dtype=np.int32)
        self.lib.ref_cbb_run(iq, len(iq), 0, dummy_u8, dummy_f64, dummy_i32, 0, None, 0, None, 0)
        self.lib.ref_ws_attach(0)
        if self.lib.ref_ws_overflowed():
            raise RuntimeError("ws log overflow")
        k = self.lib.ref_ws_n_records()
        out = []
        for off, ln, mode in rec[:k]:
            out.append((int(mode), log[off:off + ln].tobytes()))
        return out


class DropIn(Ref):
    """The drop-in proof: the SAME harness and the reference's unmodified glue, but spectrum_*,
    rf_decimator_*, cic_decimate and halfband_decimate resolve to libb200sdr.so (the GPU).
    Only cbb_run / fm_chain / spectrum_rows make sense here (ref_atan2_approx etc. still exist)."""

    SO = DROPIN_SO

    def _preload(self):
        # the private copy lives in a temp dir, so its $ORIGIN rpath no longer finds the product
        # library: load it first, by path; the copy's DT_NEEDED entry then matches it by soname.
        # NOT into the global namespace: the product exports the reference's own symbol names, and a
        # global copy would interpose them into every Ref() loaded afterwards
        C.CDLL(PRODUCT_SO)


class DropInAudio(DropIn):
    """As DropIn, with audio_main.o replaced too: the unmodified main.c / cbb_main.c register and drain the
    GPU demodulator of libb200audio.so (audio_main.h interface) -- atan2, limiter and both half-bands on the GPU."""

    SO = DROPIN_AUDIO_SO

    def _preload(self):
        C.CDLL(PRODUCT_SO)
        C.CDLL(os.path.join(os.path.dirname(PRODUCT_SO), "libb200audio.so"))


def have_dropin_audio() -> bool:
    return os.path.exists(DROPIN_AUDIO_SO) and os.path.exists(PRODUCT_SO)


def have_dropin() -> bool:
    return os.path.exists(DROPIN_SO) and os.path.exists(PRODUCT_SO)
