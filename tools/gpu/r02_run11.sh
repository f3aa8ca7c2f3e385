mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_spectrum.py -x -q -k "65536 or other_frame or full_size" > gpurun_out/r02/pytest_64k.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02/pytest_64k.log
for i in 1 2; do timeout 200 python tools/kbench.py --streams 256 --only spectrum65536_hann_50pct 2>&1 | tail -1; done
