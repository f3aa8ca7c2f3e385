mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/kbench.py --streams 256 > gpurun_out/s5_kbench.log 2>&1; cat gpurun_out/s5_kbench.log
