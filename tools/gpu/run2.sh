mkdir -p gpurun_out
python -m pytest tests/test_gpu_spectrum.py -x -q -k "4096" 2>&1 | tail -3
python tools/kbench.py --only spectrum4096_hann_db --streams 256
python tools/kbench.py --only spectrum4096_db --streams 256
