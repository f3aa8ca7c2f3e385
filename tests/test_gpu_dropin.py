"""The drop-in claim end to end: the reference's UNMODIFIED glue -- signal_source.c worker
thread, cbb_main.c (250 ms cadence, 6-frame averaging, dB payload), audio_main.c (discriminator
on the host, its two halfband_decimate calls), main.c's init order -- linked against
libb200sdr.so in place of spectrum.o / resample.o / rf_decimator.o (oracle/Makefile target
`dropin`), fed by the synthetic sensor, compared with what the all-reference build produced
for the same capture (tests/golden/cbb_gain*.npz)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import pyoracle as _po

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not _po.have_dropin(), reason="oracle/_ref/libdropin_rtlws.so not built")]

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("gain", [0, 17])
def test_reference_driver_on_the_gpu_library(pkg, cuda, po, synth, gain):
    g = np.load(os.path.join(GOLD, f"cbb_gain{gain}.npz"))
    iq = synth.s2_tones(int(g["n"]), N=1024, seed=int(g["seed"]))
    assert hashlib.sha256(iq.tobytes()).hexdigest() == str(g["iq_sha"])
    before = pkg.launch_count()
    out = po.DropIn().cbb_run(iq, gain_db=gain)
    assert pkg.launch_count() - before >= 18 + 5      # 3 spectra x 6 frames, 5 decimator blocks: all on the GPU
    assert out["payload"].shape == g["payload"].shape and (out["count"] == 6).all()
    # power_spectrum_transfer (cbb_main.c:64-69): f32 transform vs f64
    mean = g["power"].mean(axis=1, keepdims=True)
    assert (np.abs(out["power"] - g["power"]) <= 1e-4 * g["power"] + 1e-4 * mean).all()
    # payload bytes (cbb_main.c:121-130)
    for r in range(len(g["payload"])):
        _, dbf = po.db_payload(g["power"][r], 6, gain)
        diff = out["payload"][r] != g["payload"][r]
        assert diff.mean() < 0.01 and (np.abs(dbf[diff] - np.rint(dbf[diff])) <= 0.01).all()
    # the FM branch: CIC on the GPU is bit-exact, so the host discriminator sees identical input;
    # the two half-bands run on the GPU
    dec_want, audio_want = po.chain_run(iq)
    assert np.array_equal(out["decimated"], dec_want)
    assert np.abs(audio_want[:2048] - g["audio_head"]).max() < 1e-6
    assert out["audio"].shape == audio_want.shape
    assert np.abs(out["audio"] - audio_want).max() <= 1e-4


def test_reference_fm_branch_on_the_gpu_library(pkg, cuda, po, synth):
    iq = synth.s3_fm(3 * 204800 + 4321, seed=77)
    dec, audio = po.DropIn().fm_chain(iq, chunk=131072)
    dec_want, audio_want = po.chain_run(iq)
    assert np.array_equal(dec, dec_want)
    assert np.abs(audio - audio_want).max() <= 1e-4
