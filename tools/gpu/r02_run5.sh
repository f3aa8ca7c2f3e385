# round 2, run 5: why the cluster kernel is slow -- diff magnitude vs the scratch kernel, one full ncu capture
mkdir -p gpurun_out/r02
timeout 200 python - <<'PY' 2>&1 | tail -12
import os, sys, torch
sys.path.insert(0, os.getcwd())
import __graft_entry__ as g
pkg = g.load_package(); pkg.init(0)
gen = torch.Generator(device="cuda").manual_seed(1)
hop, rows, S = 32768, 47, 3
n = 65536 + hop * (rows - 1)
iq = torch.randint(0, 256, (S, n, 2), dtype=torch.uint8, device="cuda", generator=gen)
plan = pkg.SpectrumPlan(65536, hop=hop, window=pkg.WINDOW_HANN)
a = plan.exec(iq, db=True, power=True)
torch.cuda.synchronize()
a2 = plan.exec(iq, db=True, power=True)
torch.cuda.synchronize()
os.environ["B200_S64K_SCRATCH"] = "1"
b = plan.exec(iq, db=True, power=True)
torch.cuda.synchronize()
print("cluster run-to-run equal:", torch.equal(a["power"], a2["power"]))
d = (a["power"] - b["power"]).abs() / b["power"].abs().clamp_min(1e-30)
print("rel diff max", d.max().item(), "count nonzero", (a["power"] != b["power"]).sum().item(), "of", d.numel())
idx = (a["power"] != b["power"]).nonzero()
print("first diffs", idx[:8].tolist())
print("cols mod 1024 of diffs (unique count)", torch.unique(idx[:, 2] % 1024).numel(), "cols // 1024 unique", torch.unique(idx[:, 2] // 1024).numel())
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spectrum64k_cluster -s 3 -c 1 -o gpurun_out/r02/s64k_cluster_v0 -f python tools/kbench.py --only spectrum65536_hann_50pct --streams 256 --reps 2 > gpurun_out/r02/ncu_s64k.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r02/ncu_s64k.log
