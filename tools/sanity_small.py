"""Tiny run of every kernel (for compute-sanitizer): small shapes, no timing."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as g

pkg = g.load_package()
pkg.init(0)
n = 5120 * 3
iq = np.stack([pkg.synth.s3_fm(n, seed=s) for s in range(3)])
ring = pkg.StreamRing(3, n)
ring.load(iq)
db, audio = pkg.chain_exec(ring)
a2, dec = pkg.fm_exec(ring, decimated=True)
ring.carry()
for N, K, hop, win in ((1024, 1, None, 0), (1024, 6, None, 0), (1024, 2, 512, 1), (2048, 1, None, 0), (4096, 2, None, 1),
                       (8192, 1, None, 0), (512, 1, None, 0)):
    pkg.SpectrumPlan(N, hop=hop, K=K, window=win).exec(ring.batch, db=True, power=True, db_u8=True)
big = pkg.StreamRing(1, 65536 * 2)
big.load(pkg.synth.s2_tones(65536 * 2, N=65536)[None])
pkg.SpectrumPlan(65536, hop=32768, window=1).exec(big.batch, db=True)
pkg.SpectrumPlan(65536, K=2).exec(big.batch, power=True)
# round 2: the four-step kernel at 16 / 32 branches, the opt-in cluster kernel, the cs32 demodulator, the audio extensions
pkg.SpectrumPlan(16384, hop=8192, window=1).exec(big.batch, db=True, power=True, db_u8=True)
pkg.SpectrumPlan(32768, K=2).exec(big.batch, db=True)
os.environ["B200_S64K_CLUSTER"] = "1"
pkg.SpectrumPlan(65536, hop=32768, window=1).exec(big.batch, db=True, power=True, db_u8=True)
pkg.SpectrumPlan(65536).exec(big.batch, db=True)
del os.environ["B200_S64K_CLUSTER"]
dm = pkg.FmDemod()
dm.block(dec[0][:4096].cpu().numpy(), want_demod=True)
dm.block(dec[0][4096:8192].cpu().numpy(), want_audio=False)
dm.close()
st = torch.zeros((3, pkg.AUDIO_POST_STATE_FLOATS), dtype=torch.float32, device="cuda")
pkg.audio_post(audio[:, :368], st, pkg.AUDIO_DEEMPH_50US | pkg.AUDIO_RESAMPLE_48K)
r7 = pkg.StreamRing(2, 4 * 7 * 8 * 9, R=7)
r7.load(np.stack([pkg.synth.s1_noise(4 * 7 * 8 * 9, seed=s) for s in range(2)]))
pkg.fm_exec(r7, decimated=True)
sess = pkg.Session(3, n)
hdb = np.empty((3, n // 1024, 1024), np.float32)
hau = np.empty((3, n // 40), np.float32)
sess.chain(np.ascontiguousarray(iq), n, hdb, hau)
ps = pkg.PushStream(2, 5120 * 2)
for s in range(2):
    ps.push(s, iq[s])
ps.flush()
torch.cuda.synchronize()
print("sanity ok", float(db.abs().max()) > 0, float(audio.abs().max()) >= 0)
