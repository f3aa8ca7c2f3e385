#!/usr/bin/env python
"""tools/kbench.py -- per-kernel timings (CUDA events) for the pieces of the chain, device-resident.

  python tools/kbench.py [--streams 64] [--samples 2048000] [--reps 10] [--only NAME]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as graft  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=64)
ap.add_argument("--samples", type=int, default=5120 * 400)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--only", default="")
args = ap.parse_args()

pkg = graft.load_package()
pkg.init(0)
S, L = args.streams, args.samples
ring = pkg.StreamRing(S, L)
g = torch.Generator(device="cuda").manual_seed(0)
ring.batch.copy_(torch.randint(0, 256, (S, L, 2), dtype=torch.uint8, device="cuda", generator=g))
db = torch.empty((S, L // 1024, 1024), dtype=torch.float32, device="cuda")
audio = torch.empty((S, L // 40), dtype=torch.float32, device="cuda")
plan = pkg.SpectrumPlan(1024)
plan6 = pkg.SpectrumPlan(1024, K=6)
out = {"db": db}
db6 = torch.empty((S, L // 6144, 1024), dtype=torch.float32, device="cuda")
plan4k = pkg.SpectrumPlan(4096)
db4k = torch.empty((S, L // 4096, 4096), dtype=torch.float32, device="cuda")

plan64k = pkg.SpectrumPlan(65536, hop=32768, window=pkg.WINDOW_HANN)
rows64k = plan64k.rows(L)
db64k = torch.empty((S, rows64k, 65536), dtype=torch.float32, device="cuda")
plan1k_h = pkg.SpectrumPlan(1024, window=pkg.WINDOW_HANN)
plan4k_h = pkg.SpectrumPlan(4096, window=pkg.WINDOW_HANN)
plan8k = pkg.SpectrumPlan(8192)
db8k = torch.empty((S, L // 8192, 8192), dtype=torch.float32, device="cuda")
plan16k = pkg.SpectrumPlan(16384)
db16k = torch.empty((S, L // 16384, 16384), dtype=torch.float32, device="cuda")
plan32k = pkg.SpectrumPlan(32768)
db32k = torch.empty((S, L // 32768, 32768), dtype=torch.float32, device="cuda")
plan2k = pkg.SpectrumPlan(2048)
db2k = torch.empty((S, L // 2048, 2048), dtype=torch.float32, device="cuda")

cases = {
    "spectrum1024_db": (lambda: plan.exec(ring.batch, db=True, out=out), 6.0),
    "spectrum1024_k6_db": (lambda: plan6.exec(ring.batch, db=True, out={"db": db6}), 2.0 + 4.0 / 6),
    "fm_chain": (lambda: pkg.fm_exec(ring, audio=audio), 2.1),
    "chain_fused": (lambda: pkg.chain_exec(ring, db=db, audio=audio), 6.1),
    "spectrum4096_db": (lambda: plan4k.exec(ring.batch, db=True, out={"db": db4k}), 6.0),
    "spectrum1024_hann_db": (lambda: plan1k_h.exec(ring.batch, db=True, out=out), 6.0),
    "spectrum4096_hann_db": (lambda: plan4k_h.exec(ring.batch, db=True, out={"db": db4k}), 6.0),
    "spectrum2048_db": (lambda: plan2k.exec(ring.batch, db=True, out={"db": db2k}), 6.0),
    "spectrum8192_db": (lambda: plan8k.exec(ring.batch, db=True, out={"db": db8k}), 6.0),
    "spectrum16384_db": (lambda: plan16k.exec(ring.batch, db=True, out={"db": db16k}), 6.0),
    "spectrum32768_db": (lambda: plan32k.exec(ring.batch, db=True, out={"db": db32k}), 6.0),
    "spectrum65536_hann_50pct": (lambda: plan64k.exec(ring.batch, db=True, out={"db": db64k}), 10.0),
}
for name, (fn, bps) in cases.items():
    if args.only and args.only != name:
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    gs = S * L / ms / 1e6
    print(f"{name:22s} {ms:8.3f} ms  {gs:9.1f} Gsamples/s  {gs * bps:8.1f} GB/s algorithmic ({bps:.2f} B/sample)", flush=True)
