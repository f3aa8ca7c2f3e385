# round 2, run 1: the job-queue variants of the fused chain kernel -- cross-check against the round-1 kernel,
# burst and sustained timings, the chain parity tests per variant, then every kernel of tools/kbench.py
mkdir -p gpurun_out/r02
timeout 600 python tools/gpu/chain_variants.py > gpurun_out/r02/variants.log 2>&1; echo "variants rc=$?"
cat gpurun_out/r02/variants.log | tail -12
for V in 1 2 6 7; do
  B200_CHAIN_VARIANT=$V timeout 300 python -m pytest tests/test_gpu_fm.py -q -x -k "chain" 2>&1 | tail -2
done
timeout 300 compute-sanitizer --tool memcheck python tools/sanity_small.py 2>&1 | tail -4
timeout 600 python tools/kbench.py --streams 256 > gpurun_out/r02/kbench_run1.log 2>&1; echo "kbench rc=$?"
cat gpurun_out/r02/kbench_run1.log
