mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_spectrum.py -x -q -k "65536 or other_frame or full_size" > gpurun_out/r02/pytest_64k.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02/pytest_64k.log
timeout 200 python tools/kbench.py --streams 256 --only spectrum65536_hann_50pct 2>&1 | tail -1
B200_S64K_8WARP=1 timeout 200 python tools/kbench.py --streams 256 --only spectrum65536_hann_50pct 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spectrum64k16 -s 3 -c 1 -o gpurun_out/r02/s64k16_v1 -f python tools/kbench.py --only spectrum65536_hann_50pct --streams 256 --reps 2 > gpurun_out/r02/ncu_s64k.log 2>&1; echo "ncu rc=$?"
