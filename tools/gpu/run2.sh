mkdir -p gpurun_out
python -m pytest tests/test_gpu_spectrum.py -x -q 2>&1 | tail -3
