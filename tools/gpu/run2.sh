mkdir -p gpurun_out
python -m pytest tests/test_gpu_spectrum.py tests/test_gpu_wire.py -x -q 2>&1 | tail -3
python tools/kbench.py --only spectrum1024_db --streams 256 --samples 8192000
python tools/kbench.py --only spectrum1024_hann_db --streams 256 --samples 8192000
python tools/kbench.py --only spectrum1024_db --streams 256
