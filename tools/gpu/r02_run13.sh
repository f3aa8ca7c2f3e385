mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_spectrum.py -x -q > gpurun_out/r02/pytest_spectrum.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02/pytest_spectrum.log
for k in spectrum16384_db; do timeout 200 python tools/kbench.py --streams 256 --only $k 2>&1 | tail -1; done
