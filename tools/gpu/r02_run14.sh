mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_spectrum.py -x -q -k "parseval" > gpurun_out/r02/pytest_parseval.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02/pytest_parseval.log
