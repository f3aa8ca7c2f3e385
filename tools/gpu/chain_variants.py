#!/usr/bin/env python
"""Round-2 tuning aid: time and cross-check the variants of the fused chain kernel (B200_CHAIN_VARIANT).

  python tools/gpu/chain_variants.py [--variants 0,1,2] [--streams 256] [--samples 8192000]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as graft  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--variants", default="0,1,2,3,4,5,6")
ap.add_argument("--streams", type=int, default=256)
ap.add_argument("--samples", type=int, default=5120 * 1600)
ap.add_argument("--long-reps", type=int, default=300)
args = ap.parse_args()

pkg = graft.load_package()
pkg.init(0)
S, L = args.streams, args.samples
ring = pkg.StreamRing(S, L)
g = torch.Generator(device="cuda").manual_seed(0)
ring.batch.copy_(torch.randint(0, 256, (S, L, 2), dtype=torch.uint8, device="cuda", generator=g))
# a few streams of strong FM-like structure so the discriminator is not all limiter
t = torch.arange(L, device="cuda", dtype=torch.float32)
for s in range(min(S, 4)):
    ph = 0.02 * (s + 1) * t + 3.0 * torch.sin(t * 0.001 * (s + 1))
    ring.batch[s, :, 0] = (127.5 + 100 * torch.cos(ph)).round().clamp(0, 255).to(torch.uint8)
    ring.batch[s, :, 1] = (127.5 + 100 * torch.sin(ph)).round().clamp(0, 255).to(torch.uint8)
db = torch.empty((S, L // 1024, 1024), dtype=torch.float32, device="cuda")
audio = torch.empty((S, L // 40), dtype=torch.float32, device="cuda")
ref_db = ref_audio = None


def timed(reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        pkg.chain_exec(ring, db=db, audio=audio)
    e1.record()
    torch.cuda.synchronize()
    return S * L * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


for v in [int(x) for x in args.variants.split(",")]:
    os.environ["B200_CHAIN_VARIANT"] = str(v)
    db.zero_()
    audio.zero_()
    for _ in range(3):
        pkg.chain_exec(ring, db=db, audio=audio)
    torch.cuda.synchronize()
    if ref_db is None:
        ref_db, ref_audio = db.clone(), audio.clone()
        check = "reference"
    else:
        fin = torch.isfinite(ref_db)
        same_inf = bool((torch.isfinite(db) == fin).all())
        ddb = float((db[fin] - ref_db[fin]).abs().max())
        da = float((audio - ref_audio).abs().max())
        check = f"max|ddB| {ddb:.3g} max|daudio| {da:.3g} inf-pattern-equal {same_inf}"
    burst = timed(10)
    long = timed(args.long_reps)
    print(f"variant {v}: burst(10) {burst:7.1f} Gsamples/s   sustained({args.long_reps}) {long:7.1f} Gsamples/s   {check}", flush=True)
