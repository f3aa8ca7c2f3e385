# round 2, run 4: the cluster form of the 65536-point kernel -- parity first (short timeouts: a hung cluster must not hold the box)
mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_spectrum.py -x -q -k "65536 or other_frame or full_size" > gpurun_out/r02/pytest_64k.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02/pytest_64k.log
timeout 200 python tools/kbench.py --streams 256 --only spectrum65536_hann_50pct > gpurun_out/r02/kbench_64k_cluster.log 2>&1; echo "rc=$?"
cat gpurun_out/r02/kbench_64k_cluster.log
B200_S64K_SCRATCH=1 timeout 200 python tools/kbench.py --streams 256 --only spectrum65536_hann_50pct 2>&1 | tail -1
