"""End-to-end (host buffers, PCIe copies inside) for the two product sets of the chain: per-frame float dB rows +
audio (what bench.py's e2e measures: 4.1 bytes per sample back) and audio only (0.1 bytes per sample back; the
reference's own spectrum product, 1 KB of payload bytes per batch, adds nothing measurable)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as graft
pkg = graft.load_package(); pkg.init(0)
S, L = 256, 5120 * 100
sess = pkg.Session(S, L)
h_iq = torch.randint(0, 256, (S, L, 2), dtype=torch.uint8).pin_memory()
h_db = torch.empty((S, L // 1024, 1024), dtype=torch.float32).pin_memory()
h_audio = torch.empty((S, L // 40), dtype=torch.float32).pin_memory()
for name, db in (("dB rows + audio", h_db), ("audio only", None)):
    for _ in range(2): sess.chain(h_iq, L, db, h_audio)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): sess.chain(h_iq, L, db, h_audio)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name:18s} {S*L/dt/1e9:6.2f} Gsamples/s   H2D {2*S*L/dt/1e9:5.1f} GB/s  D2H {((4*S*L if db is not None else 0) + 4*S*L/40)/dt/1e9:5.1f} GB/s")
