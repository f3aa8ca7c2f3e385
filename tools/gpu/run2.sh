mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/kbench.py --only spectrum4096_db --streams 256
python tools/kbench.py --only spectrum2048_db --streams 256
