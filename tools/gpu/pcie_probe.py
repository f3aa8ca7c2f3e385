"""Aggregate pinned-memory PCIe bandwidth with one process per GPU (run under torchrun): what the box can
carry when all ranks copy at once.  Explains where the e2e number of bench.py saturates at N = 4 / 8."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(2 * n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=8):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return reps / t.item()


for name, a, b in (("H2D only", True, False), ("D2H only", False, True), ("both", True, True)):
    run(a, b, 2)
    r = run(a, b)
    if rank == 0:
        gb_in = r * n * world / 1e9 if a else 0.0
        gb_out = r * 2 * n * world / 1e9 if b else 0.0
        print(f"{world} ranks  {name:9s}  H2D {gb_in:7.1f} GB/s  D2H {gb_out:7.1f} GB/s aggregate", flush=True)
if rank == 0:
    os.system("nvidia-smi topo -m 2>/dev/null | head -14; nproc; numactl -H 2>/dev/null | head -6")
if world > 1:
    dist.destroy_process_group()
