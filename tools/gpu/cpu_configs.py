"""The reference's own C (oracle/_ref, stand-in FFT) on the GPU box's host cores for the CPU-runnable configs of
SURVEY.md 8d: config 1 (20 000 frames of 1024 through spectrum_add_cmplx_u8), config 3 (FM chain), config 5 (both),
on one core and on all cores (one process per core, one stream of 20 480 000 samples per process)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
cores = os.cpu_count() or 1
per = 20_480_000
out = {"cores": cores, "samples_per_stream": per}
for mode in ("spectrum", "fm", "chain"):
    r1 = bench.run_ref_bench(1, per, 1, mode)
    rn = bench.run_ref_bench(cores, per, cores, mode)
    out[mode] = {"msamples_per_s_1core": r1 and r1["msamples_per_s"], "msamples_per_s_all_cores": rn and rn["msamples_per_s"]}
out["fft_only_brackets"] = bench.fft_brackets()
print(json.dumps(out))
