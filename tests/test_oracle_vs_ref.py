"""The oracle port (oracle/oracle.c) against the UNMODIFIED reference compiled in place
(oracle/_ref/libref_rtlws.so).  Runs wherever that library exists: in the build container
(oracle/Makefile builds it from /root/reference) and on the GPU box (it travels prebuilt).
"""
import numpy as np
import pytest

from oracle import pyoracle as _po

pytestmark = pytest.mark.skipif(not _po.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


def test_spectrum_rows_match(po, synth):
    for N, K in ((1024, 1), (1024, 6), (2048, 2), (4096, 3)):
        iq = synth.s2_tones(N * K * 2, N=N, seed=20 + K)
        a = po.Ref().spectrum_rows(iq, N, K=K)
        b = po.Spectrum(N).rows(iq, K=K)
        np.testing.assert_allclose(b, a, rtol=1e-12, atol=1e-9)


def test_spectrum_accumulates_into_dirty_buffer(po, synth):
    # the caller's array is read-modify-write; the DC position adds its UPDATED left neighbour
    iq = synth.s1_noise(1024, seed=9)
    rng = np.random.default_rng(10)
    init = rng.uniform(0, 100, 1024)
    ref = po.Ref()
    s = ref.lib.spectrum_alloc(1024)
    a = init.copy()
    assert ref.lib.spectrum_add_cmplx_u8(s, iq.reshape(-1), a, 1024) == 0
    assert ref.lib.spectrum_add_cmplx_u8(s, iq.reshape(-1), a, 512) == -1
    ref.lib.spectrum_free(s)
    b = init.copy()
    assert po.Spectrum(1024).add_cmplx_u8(iq, b) == 0
    np.testing.assert_allclose(b, a, rtol=1e-13)
    assert np.isclose(a[512], init[512] + a[511])


def test_cic_bit_exact_random_R(po, synth):
    iq = synth.s1_noise(50_000, seed=3)
    for R in (1, 3, 7, 10, 16, 25):
        n = (len(iq) // R) * R
        r1, d1, s1 = po.Ref().cic_decimate(R, iq[:n])
        r2, d2, s2 = po.cic_decimate(R, iq[:n])
        assert r1 == r2 == 0 and np.array_equal(d1, d2)
        assert list(s1.integrator_prev_out) == list(s2.integrator_prev_out)
        assert list(s1.comb_prev_in) == list(s2.comb_prev_in)
    assert po.Ref().cic_decimate(10, iq[:1000], dst_len=99)[0] == -1


def test_atan2_exhaustive_band(po):
    # every (y, x) on the CIC grid for a band of y: within 1 ulp of pi of the -ffast-math build
    ys = np.arange(-1280, 1271, 97, dtype=np.int32)
    xs = np.arange(-1280, 1271, 1, dtype=np.int32)
    a = po.Ref().atan2_grid(ys, xs)
    b = po.atan2_approx(ys[:, None].astype(np.float32), xs[None, :].astype(np.float32))
    assert np.abs(a - b).max() <= 2.4e-7


def test_fm_chain_matches_over_chunkings(po, synth):
    iq = synth.s3_fm(3 * 204800 + 12345, seed=31)
    d_ref, a_ref = po.Ref().fm_chain(iq, chunk=131072)
    for chunk in (131072, 204800, 9999):
        d, a = po.chain_run(iq, chunk=chunk)
        assert np.array_equal(d, d_ref)
        assert np.abs(a - a_ref).max() <= 1e-6
    d_ref2, a_ref2 = po.Ref().fm_chain(iq, chunk=9999)
    assert np.array_equal(d_ref2, d_ref) and np.array_equal(a_ref2, a_ref)


def test_fm_chain_limiter_and_other_rates(po, synth):
    iq = synth.s3_fm(2 * 204800, deviation=75_000.0, seed=32)
    _, a_ref = po.Ref().fm_chain(iq)
    _, a = po.chain_run(iq)
    assert np.abs(a - a_ref).max() <= 1e-6
    # main.c:149-155: a "bw" command re-derives down_factor = fs / 192000
    for fs, R in ((1_024_000.0, 5), (2_400_000.0, 12), (3_200_000.0, 16)):
        iq = synth.s3_fm(int(fs * 0.25), fs=fs, seed=33)
        d_ref, a_ref = po.Ref().fm_chain(iq, sample_rate=fs, down_factor=R)
        d, a = po.chain_run(iq, sample_rate=fs, down_factor=R)
        assert np.array_equal(d, d_ref) and np.abs(a - a_ref).max() <= 1e-6


def test_whole_driver_matches(po, synth):
    iq = synth.s3_fm(131072 * 6, seed=34)
    out = po.Ref().cbb_run(iq, gain_db=20)
    rows = po.Spectrum(1024).rows(iq, K=6, row_hop=4 * 131072)
    assert len(out["power"]) == len(rows) == 2
    np.testing.assert_allclose(rows, out["power"], rtol=1e-12, atol=1e-9)
    for r in range(len(rows)):
        assert np.array_equal(po.db_payload(rows[r], 6, 20)[0], out["payload"][r])
    dec, audio = po.chain_run(iq)
    assert np.array_equal(dec, out["decimated"])
    assert np.abs(audio - out["audio"]).max() <= 1e-6
