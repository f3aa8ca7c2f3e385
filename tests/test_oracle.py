"""CPU tests of the oracle port (oracle/oracle.c) against the golden vectors that the
UNMODIFIED reference produced (tests/golden/, made by tests/golden/make_golden.py), the
known-answer cases of SURVEY.md section 8c, and numpy's FFT for the one third-party routine.
"""
import hashlib
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---- the FFTW stand-in ----------------------------------------------------------------

@pytest.mark.parametrize("N", [2, 4, 8, 16, 32, 64, 512, 1024, 2048, 4096, 65536])
def test_fft_matches_numpy(po, N):
    rng = np.random.default_rng(N)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    want = np.fft.fft(x)
    got = po.fft(x)
    assert np.abs(got - want).max() <= 2e-15 * np.log2(N) * np.abs(want).max()


def test_fft_matches_naive_dft(po):
    rng = np.random.default_rng(3)
    x = rng.standard_normal(128) + 1j * rng.standard_normal(128)
    assert np.abs(po.fft(x) - po.dft_naive(x)).max() < 1e-12


def test_fft_impulse_and_tone(po):
    x = np.zeros(1024, dtype=np.complex128)
    x[1] = 1.0
    want = np.exp(-2j * np.pi * np.arange(1024) / 1024)          # forward sign convention
    assert np.abs(po.fft(x) - want).max() < 1e-14
    t = np.exp(2j * np.pi * 37 * np.arange(1024) / 1024)
    X = po.fft(t)
    assert abs(X[37] - 1024) < 1e-9 and np.abs(np.delete(X, 37)).max() < 1e-9


# ---- spectrum.c ---------------------------------------------------------------------------

def test_spectrum_golden_n1024(po):
    g = gold("spectrum_n1024.npz")
    s = po.Spectrum(1024)
    np.testing.assert_allclose(s.rows(g["iq"], K=1)[:3], g["rows_k1"], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(s.rows(g["iq"], K=6), g["rows_k6"], rtol=1e-12, atol=1e-9)


def test_spectrum_golden_n4096_n65536(po, synth):
    g = gold("spectrum_n4096.npz")
    s = po.Spectrum(4096)
    np.testing.assert_allclose(s.rows(g["iq"], K=1)[:1], g["rows_k1"], rtol=1e-12, atol=1e-8)
    np.testing.assert_allclose(s.rows(g["iq"], K=3), g["rows_k3"], rtol=1e-12, atol=1e-8)
    g = gold("spectrum_n65536.npz")
    iq = synth.s2_tones(65536, N=65536, seed=int(g["seed"]))
    assert sha(iq) == str(g["iq_sha"]), "synthetic generator drifted: regenerate tests/golden"
    np.testing.assert_allclose(po.Spectrum(65536).rows(iq), g["rows_k1"], rtol=1e-11, atol=1e-6)


def test_spectrum_known_answers(po):
    g = gold("spectrum_kat.npz")
    s = po.Spectrum(1024)
    # (i) DC-only input -> all-zero spectrum, payload bytes all 0
    rows = s.rows(g["iq_dc"])
    assert np.array_equal(rows, g["rows_dc"]) and not rows.any()
    payload, _ = po.db_payload(rows[0], 1, 0)
    assert not payload.any()
    # (ii) tone on bin 100 lands at display index 612 with power (N*A/128)^2 per frame; K = 2 doubles it
    rows = s.rows(g["iq_tone"], K=2)
    np.testing.assert_allclose(rows, g["rows_tone_k2"], rtol=1e-12, atol=1e-9)
    assert rows[0].argmax() == 612
    assert abs(rows[0, 612] / (2 * (1024 * 127 / 128) ** 2) - 1) < 2e-2     # u8 rounding of the tone
    # the DC position (index 512) holds the cumulative sum of its left neighbour: after two adds
    # into a zeroed row that is 2*P1 + 1*P2 of bin N-1 (spectrum.c:30-33)
    one = po.Spectrum(1024)
    p1 = one.rows(g["iq_tone"][:1024])[0]
    p2 = one.rows(g["iq_tone"][1024:])[0]
    assert np.isclose(rows[0, 512], 2 * p1[511] + p2[511], rtol=1e-12)
    assert np.isclose(rows[0, 511], p1[511] + p2[511], rtol=1e-12)


def test_spectrum_wrong_length_is_minus_one(po):
    s = po.Spectrum(1024)
    ps = np.zeros(1024)
    assert s.add_cmplx_u8(np.zeros((512, 2), np.uint8), ps, 512) == -1
    assert s.add_cmplx_s32(np.zeros((512, 2), np.int32), ps, 512) == -1
    assert s.add_real_f32(np.zeros(512, np.float32), ps, 512) == -1
    assert not ps.any()


def test_spectrum_other_input_types(po):
    rng = np.random.default_rng(8)
    u8 = rng.integers(0, 256, size=(1024, 2), dtype=np.uint8)
    s = po.Spectrum(1024)
    a = np.zeros(1024)
    b = np.zeros(1024)
    s.add_cmplx_u8(u8, a)
    s.add_cmplx_s32(u8.astype(np.int32) - 128, b)       # same values after each path's own scaling
    np.testing.assert_allclose(a, b, rtol=1e-13)
    x = rng.standard_normal(1024).astype(np.float32)
    c = np.zeros(1024)
    s.add_real_f32(x, c)
    want = np.abs(np.fft.fftshift(np.fft.fft(x.astype(np.float64)))) ** 2
    want[512] = want[511]
    np.testing.assert_allclose(c, want, rtol=1e-10, atol=1e-9)


def test_window_reduces_to_reference_when_rectangular(po, synth):
    iq = synth.s2_tones(2048)
    a = po.Spectrum(1024).rows(iq)
    b = po.Spectrum(1024, window=np.ones(1024)).rows(iq)
    assert np.array_equal(a, b)
    c = po.Spectrum(1024, window=synth.hann(1024)).rows(iq)
    x = (iq[:1024].astype(np.float64) - 128) / 128
    want = np.abs(np.fft.fftshift(np.fft.fft((x[:, 0] + 1j * x[:, 1]) * synth.hann(1024)))) ** 2
    want[512] = want[511]
    np.testing.assert_allclose(c[0], want, rtol=1e-10, atol=1e-12)


# ---- cbb_main.c: the whole driver ------------------------------------------------------------

@pytest.mark.parametrize("gain", [0, 17, 30])
def test_cbb_driver_golden(po, synth, gain):
    g = gold(f"cbb_gain{gain}.npz")
    iq = synth.s2_tones(int(g["n"]), N=1024, seed=int(g["seed"]))
    assert sha(iq) == str(g["iq_sha"]), "synthetic generator drifted: regenerate tests/golden"
    # 250 ms gate on 64 ms buffers -> every 4th source buffer, first 6 frames (cbb_main.c:44-59)
    rows = po.Spectrum(1024).rows(iq, K=6, row_hop=4 * 131072)
    assert len(rows) == len(g["power"]) and (g["count"] == 6).all()
    np.testing.assert_allclose(rows, g["power"], rtol=1e-12, atol=1e-9)
    for r in range(len(rows)):
        payload, _ = po.db_payload(rows[r], 6, gain)
        assert np.array_equal(payload, g["payload"][r])
    # gain 17 behaves as gain 10: integer division (cbb_main.c:112)
    if gain == 17:
        p10, _ = po.db_payload(rows[0], 6, 10)
        assert np.array_equal(p10, g["payload"][0])
    dec, audio = po.chain_run(iq)
    assert sha(audio) == str(g["audio_sha"]) or np.abs(audio[:2048] - g["audio_head"]).max() < 1e-6


# ---- resample.c ----------------------------------------------------------------------------

@pytest.mark.parametrize("R", [1, 2, 5, 10, 12, 16])
def test_cic_golden_bit_exact(po, R):
    g = gold("cic.npz")
    iq = g["iq"]
    n = (len(iq) // R) * R
    cut = n // 2 // R * R
    st = po.CicState()
    r1, d1, st = po.cic_decimate(R, iq[:cut], st)
    r2, d2, st = po.cic_decimate(R, iq[cut:n], st)
    assert r1 == 0 and r2 == 0
    assert np.array_equal(np.concatenate([d1, d2]), g[f"R{R}"])
    assert list(st.integrator_prev_out) + list(st.comb_prev_in) == list(g[f"R{R}_state"])


def test_cic_known_answers_and_errors(po):
    # (iv) constant 129 -> every output (R, R)
    _, d, _ = po.cic_decimate(10, np.full((200, 2), 129, np.uint8))
    assert (d == 10).all()
    # boxcar identity on random bytes
    rng = np.random.default_rng(1)
    iq = rng.integers(0, 256, size=(1000, 2), dtype=np.uint8)
    _, d, st = po.cic_decimate(10, iq)
    want = (iq.astype(np.int32) - 128).reshape(100, 10, 2).sum(axis=1)
    assert np.array_equal(d, want)
    assert list(st.integrator_prev_out) == list(want.sum(axis=0)) == list(st.comb_prev_in)
    # size mismatch -> -1 (resample.c:18-19)
    r, _, _ = po.cic_decimate(10, iq, dst_len=99)
    assert r == -1


def test_halfband_golden(po):
    g = gold("halfband.npz")
    imp = np.zeros(64, dtype=np.float32)
    imp[0] = 1.0
    out = po.halfband_decimate(imp, np.zeros(10, np.float32))
    # (vii) impulse response = every other tap of h, starting with h[0]
    assert np.array_equal(out, g["impulse_out"])
    np.testing.assert_allclose(out[:6], [0.01824, -0.11614, 0.34790, 0.34790, -0.11614, 0.01824], rtol=1e-6)
    delay = np.zeros(10, np.float32)
    y = np.concatenate([po.halfband_decimate(g["x"][:1000], delay), po.halfband_decimate(g["x"][1000:], delay)])
    np.testing.assert_allclose(y, g["y"], rtol=0, atol=5e-7)       # reference is -ffast-math
    assert np.array_equal(delay, g["delay_after"])


# ---- common_sp.h ------------------------------------------------------------------------------

def test_atan2_golden(po):
    g = gold("atan2.npz")
    got = po.atan2_approx(g["ys"][:, None].astype(np.float32), g["xs"][None, :].astype(np.float32))
    assert np.abs(got - g["grid"]).max() <= 2.4e-7          # 1 ulp at pi (reference is -ffast-math)
    d = g["diag"].astype(np.float32)
    for name, y, x in (("diag_pp", d, d), ("diag_pm", d, -d), ("axis_y0", np.zeros_like(d), d),
                       ("axis_x0", d, np.zeros_like(d))):
        assert np.abs(po.atan2_approx(y, x) - g[name]).max() <= 2.4e-7, name
    # the branch discontinuity at |y| == |x| sits on the second branch: pi/2 - 1/1.28
    assert abs(float(po.atan2_approx(5.0, 5.0)) - (np.pi / 2 - 1 / 1.28)) < 1e-6
    assert float(po.atan2_approx(0.0, 0.0)) == 0.0


# ---- audio_main.c + rf_decimator.c ---------------------------------------------------------------

def test_fm_demod_block_golden(po):
    g = gold("fm_demod_block.npz")
    demod, work, audio, _ = po.fm_demodulate(g["signal"])
    assert np.abs(demod - g["demod"]).max() <= 5e-7
    assert np.abs(audio - g["audio"]).max() <= 5e-7
    assert demod.max() <= 1.0 and demod.min() >= -1.0       # the hard limiter


def test_fm_carrier_known_answer(po):
    # (vi) an unmodulated carrier at +f gives a constant discriminator output ~ 2*pi*f/fs_dec
    f, fs_dec = 10_000.0, 204_800.0
    n = np.arange(4096)
    sig = np.stack([np.rint(1000 * np.cos(2 * np.pi * f * n / fs_dec)),
                    np.rint(1000 * np.sin(2 * np.pi * f * n / fs_dec))], axis=1).astype(np.int32)
    demod, _, audio, _ = po.fm_demodulate(sig)
    want = 2 * np.pi * f / fs_dec
    wraps = np.abs(demod[1:] - want) > 0.02            # every +-pi wrap is clipped to -1, not unwrapped
    assert wraps.mean() < 0.06 and np.all(demod[1:][wraps] == -1.0)
    assert abs(np.median(demod[1:]) - want) < 0.01     # atan2_approx is good to ~5e-3 rad


@pytest.mark.parametrize("name", ["fm_chain_25k.npz", "fm_chain_75k.npz"])
def test_fm_chain_golden(po, synth, name):
    g = gold(name)
    iq = synth.s3_fm(int(g["n"]), deviation=float(g["deviation"]), seed=int(g["seed"]))
    assert sha(iq) == str(g["iq_sha"]), "synthetic generator drifted: regenerate tests/golden"
    dec, audio = po.chain_run(iq, chunk=131072)
    assert sha(dec) == str(g["dec_sha"]) and np.array_equal(dec[:4096], g["dec_head"])    # CIC: bit-exact
    assert audio.shape == g["audio"].shape
    assert np.abs(audio - g["audio"]).max() <= 1e-6
    # re-blocking makes the result independent of how the stream is chunked (rf_decimator.c:88-115)
    dec2, audio2 = po.chain_run(iq, chunk=4099)
    assert np.array_equal(dec, dec2) and np.array_equal(audio, audio2)


def test_fm_chain_small_golden_custom_rate(po):
    g = gold("fm_chain_small.npz")
    dec, audio = po.chain_run(g["iq"], chunk=7001, sample_rate=204_800.0, down_factor=10)
    assert np.array_equal(dec, g["dec"])
    assert np.abs(audio - g["audio"]).max() <= 1e-6


def test_chain_rejects_bad_parameters(po):
    with pytest.raises(ValueError):
        po.Chain(0.0, 10)
    with pytest.raises(ValueError):
        po.Chain(2048000.0, 0)


# ---- definitions of the product's opt-in extensions (no reference counterpart) ------------------------

def test_extension_definitions_deemphasis_and_resampler(po):
    """The reference has neither (audio_main.c:133-139 ends at 51.2 kS/s, no de-emphasis); these are the
    definitions csrc/audio_post.cu is checked against, so pin their basic properties."""
    h = po.resample_taps()
    assert h.shape == (240,) and np.allclose(h, h[::-1], atol=1e-9)            # linear phase
    assert abs(h.sum() - 15.0) < 1e-4
    assert all(abs(h[r::15].sum() - 1.0) < 5e-5 for r in range(15))            # unity DC gain in every phase
    x = np.ones(16 * 40, np.float32)
    y, hist = po.resample_15_16(x)
    assert len(y) == 15 * 40 and np.abs(y[30:] - 1.0).max() < 5e-5 and np.array_equal(hist, np.ones(15, np.float32))
    # split input = same output (the history carries)
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, 16 * 64).astype(np.float32)
    whole, _ = po.resample_15_16(x)
    a, hist = po.resample_15_16(x[:16 * 10])
    b, _ = po.resample_15_16(x[16 * 10:], hist)
    assert np.array_equal(np.concatenate([a, b]), whole)
    with pytest.raises(ValueError):
        po.resample_15_16(x[:24])
    # de-emphasis: first-order low-pass with the textbook time constant, -3 dB at 1 / (2 pi tau)
    y, st = po.deemphasis(np.ones(4096, np.float32), 51200.0, 75e-6)
    alpha = 1 - np.exp(-1 / (51200.0 * 75e-6))
    assert abs(y[0] - alpha) < 1e-6 and abs(y[-1] - 1.0) < 1e-5 and st[0] == y[-1]
    t = np.arange(51200) / 51200.0
    f3 = 1 / (2 * np.pi * 75e-6)
    y, _ = po.deemphasis(np.sin(2 * np.pi * f3 * t).astype(np.float32), 51200.0, 75e-6)
    assert abs(20 * np.log10(np.abs(y[5000:]).max()) + 3.0) < 0.35             # bilinear-free one-pole: close to -3 dB
