mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_spectrum.py -x -q -k "cluster" 2>&1 | tail -2
B200_S64K_CLUSTER=1 timeout 200 python tools/kbench.py --streams 256 --only spectrum65536_hann_50pct 2>&1 | tail -1
B200_S64K_CLUSTER=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:spectrum64k_cluster -s 3 -c 1 -o gpurun_out/r02/s64k_cluster_v6 -f python tools/kbench.py --only spectrum65536_hann_50pct --streams 256 --reps 2 > gpurun_out/r02/ncu_s64k.log 2>&1; echo "ncu rc=$?"
