mkdir -p gpurun_out
python -m pytest tests/test_gpu_spectrum.py -x -q -k "65536" 2>&1 | tail -3
python tools/kbench.py --only spectrum65536_hann_50pct --streams 256
