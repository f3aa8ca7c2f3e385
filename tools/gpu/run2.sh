mkdir -p gpurun_out
python -m pytest tests/test_gpu_spectrum.py -x -q 2>&1 | tail -3
python tools/kbench.py --only spectrum2048_db --streams 256
python tools/kbench.py --only spectrum2048_db --streams 256 --samples 8192000
