"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container, where /root/reference is mounted:

    make -C oracle ref && python tests/golden/make_golden.py

Every output below comes from oracle/_ref/libref_rtlws.so -- the reference's own
spectrum.c, resample.c, rf_decimator.c, audio_main.c, cbb_main.c, signal_source.c compiled
in place with the reference's flags (oracle/Makefile) -- fed with the seeded synthetic
captures of rtl-ws_b200/synth.py.  The one non-reference ingredient is the FFT behind
spectrum.c: FFTW3 is absent from the image, so the transform is oracle/fft_f64.c (checked
against numpy.fft to 4e-16 in tests/test_oracle.py).

Inputs that are small are stored next to their outputs; large inputs are regenerated from
their seed at test time and pinned by a SHA-256 stored in the fixture.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

synth = graft.load_package().synth


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def save(name: str, **arrays) -> None:
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB  " + ", ".join(f"{k}{tuple(np.shape(v))}" for k, v in arrays.items()))


def main() -> None:
    assert po.have_ref(), "build oracle/_ref first (make -C oracle ref)"

    # ---- spectrum.c: tones + noise, N = 1024 (the reference's FFT_POINTS) --------------
    iq = synth.s2_tones(1024 * 12, N=1024, seed=1)
    ref = po.Ref()
    save("spectrum_n1024.npz", iq=iq,
         rows_k1=ref.spectrum_rows(iq, 1024, K=1)[:3],
         rows_k6=ref.spectrum_rows(iq, 1024, K=6))
    # other frame lengths the configs name
    iq = synth.s2_tones(4096 * 3, N=4096, seed=11)
    save("spectrum_n4096.npz", iq=iq, rows_k1=ref.spectrum_rows(iq, 4096, K=1)[:1],
         rows_k3=ref.spectrum_rows(iq, 4096, K=3))
    iq = synth.s2_tones(65536, N=65536, seed=12)
    save("spectrum_n65536.npz", iq_sha=sha(iq), seed=12, rows_k1=ref.spectrum_rows(iq, 65536, K=1))
    # known answers: DC-only input, full-scale tone on a bin, the DC-position patch
    iq_dc = np.full((1024, 2), 128, dtype=np.uint8)
    iq_tone = synth.tone(2048, 100, 1024)
    save("spectrum_kat.npz", iq_dc=iq_dc, rows_dc=ref.spectrum_rows(iq_dc, 1024),
         iq_tone=iq_tone, rows_tone_k2=ref.spectrum_rows(iq_tone, 1024, K=2))

    # ---- resample.c: cic_decimate on random bytes, several R, chained state ---------------
    iq = synth.s1_noise(6000, seed=0)
    cic = {}
    for R in (1, 2, 5, 10, 12, 16):
        n = (len(iq) // R) * R
        st = po.CicState()
        r1, d1, st = ref.cic_decimate(R, iq[:n // 2 // R * R], st)
        r2, d2, st = ref.cic_decimate(R, iq[n // 2 // R * R:n], st)
        assert r1 == 0 and r2 == 0
        cic[f"R{R}"] = np.concatenate([d1, d2])
        cic[f"R{R}_state"] = np.array(list(st.integrator_prev_out) + list(st.comb_prev_in), dtype=np.int32)
    save("cic.npz", iq=iq, **cic)

    # ---- resample.c: halfband_decimate impulse response and a chained random run -----------
    imp = np.zeros(64, dtype=np.float32)
    imp[0] = 1.0
    delay = np.zeros(10, dtype=np.float32)
    imp_out = ref.halfband_decimate(imp, delay)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(4096).astype(np.float32)
    delay = np.zeros(10, dtype=np.float32)
    y = np.concatenate([ref.halfband_decimate(x[:1000], delay), ref.halfband_decimate(x[1000:], delay)])
    save("halfband.npz", impulse_out=imp_out, x=x, y=y, delay_after=delay)

    # ---- common_sp.h: atan2_approx on the integer grid the CIC output lives on -------------
    ys = np.arange(-1280, 1271, 17, dtype=np.int32)
    xs = np.arange(-1280, 1271, 3, dtype=np.int32)
    grid = ref.atan2_grid(ys, xs)
    diag = np.arange(-1280, 1271, dtype=np.int32)
    f = ref.lib.ref_atan2_approx
    diag_pp = np.array([f(float(v), float(v)) for v in diag], dtype=np.float32)
    diag_pm = np.array([f(float(v), float(-v)) for v in diag], dtype=np.float32)
    axis_y0 = np.array([f(0.0, float(v)) for v in diag], dtype=np.float32)
    axis_x0 = np.array([f(float(v), 0.0) for v in diag], dtype=np.float32)
    save("atan2.npz", ys=ys, xs=xs, grid=grid, diag=diag, diag_pp=diag_pp, diag_pm=diag_pm,
         axis_y0=axis_y0, axis_x0=axis_x0)

    # ---- rf_decimator.c + audio_main.c: the FM branch, default parameters ------------------
    n = 204800 + 3 * 5120
    for name, dev, seed in (("fm_chain_25k.npz", 25_000.0, 2), ("fm_chain_75k.npz", 75_000.0, 3)):
        iq = synth.s3_fm(2 * 204800 + 777, deviation=dev, seed=seed)
        dec, audio = po.Ref().fm_chain(iq, chunk=131072)
        save(name, iq_sha=sha(iq), seed=seed, deviation=dev, n=len(iq),
             dec_head=dec[:4096], dec_sha=sha(dec), audio=audio)
    # a short custom-rate stream stored whole (fs = 204.8 kHz -> 2048-sample blocks), pushed in odd chunks
    iq = synth.s3_fm(5 * 20480 + 99, fs=204_800.0, deviation=8_000.0, tones=(300.0, 1100.0), seed=4)
    dec, audio = po.Ref().fm_chain(iq, chunk=7001, sample_rate=204_800.0, down_factor=10)
    save("fm_chain_small.npz", iq=iq, dec=dec, audio=audio)
    # one block straight into audio_fm_demodulator: discriminator output after the limiter
    rng = np.random.default_rng(6)
    sig = rng.integers(-1280, 1271, size=(2048, 2), dtype=np.int32)
    sig[:8] = [[0, 0], [5, 0], [-5, 0], [0, 5], [0, -5], [7, 7], [-7, 7], [7, -7]]
    demod, audio = po.Ref().fm_demodulate_block(sig)
    save("fm_demod_block.npz", signal=sig, demod=demod, audio=audio)

    # ---- the whole driver: signal_source -> cbb_main (250 ms cadence, K = 6, dB payload) ----
    iq = synth.s2_tones(131072 * 9, N=1024, seed=7)
    for gain in (0, 17, 30):
        out = po.Ref().cbb_run(iq, gain_db=gain)
        save(f"cbb_gain{gain}.npz", iq_sha=sha(iq), seed=7, n=len(iq), payload=out["payload"], power=out["power"],
             count=out["count"], audio_sha=sha(out["audio"]), audio_head=out["audio"][:2048])

    # ---- the websocket callback of the unmodified main.c on top of the driver (main.c:74-111) ----
    iq = synth.s3_fm(131072 * 7, seed=79)
    freq, rate, gain = 99_900_000, 2_048_000, 17
    records = po.Ref().ws_run(iq, commands=(f"spectrumgain {gain}", f"freq {freq // 1000}", "start"))
    drv = po.Ref().cbb_run(iq, gain_db=gain)
    blob = np.frombuffer(b"".join(b for _, b in records), dtype=np.uint8)
    offs = np.cumsum([0] + [len(b) for _, b in records[:-1]])
    rec = np.array([[o, len(b), m] for o, (m, b) in zip(offs, records)], dtype=np.int64)
    save("ws_stream.npz", iq_sha=sha(iq), seed=79, n=len(iq), freq=freq, rate=rate, gain=gain, records=rec, bytes=blob,
         payload=drv["payload"], audio=drv["audio"])


if __name__ == "__main__":
    main()
