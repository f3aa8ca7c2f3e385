"""b200_fm_exec throughput for the down factors the reference's `bw` command can produce
(R = fs / 192000, main.c:154): 5 (1.024 MS/s), 10 (2.048), 12 (2.4), 16 (3.2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as graft
pkg = graft.load_package(); pkg.init(0)
S = 256
for R in (5, 7, 10, 12, 13, 15, 16, 20):
    L = (2048000 // (4 * R * 8)) * (4 * R * 8)
    ring = pkg.StreamRing(S, L, R=R)
    g = torch.Generator(device="cuda").manual_seed(0)
    ring.batch.copy_(torch.randint(0, 256, (S, L, 2), dtype=torch.uint8, device="cuda", generator=g))
    audio = torch.empty((S, L // (4 * R)), dtype=torch.float32, device="cuda")
    for _ in range(3): pkg.fm_exec(ring, audio=audio)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): pkg.fm_exec(ring, audio=audio)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    bps = 2 + 4 / (4 * R)
    print(f"R={R:3d}  {ms:7.3f} ms  {S*L/ms/1e6:8.1f} Gsamples/s  {S*L/ms/1e6*bps:7.1f} GB/s algorithmic ({bps:.2f} B/sample)")
