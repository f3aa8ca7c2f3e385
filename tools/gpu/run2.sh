mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/kbench.py --streams 256 > gpurun_out/r01c_kbench.log 2>&1; cat gpurun_out/r01c_kbench.log
python tools/kbench.py --streams 256 --samples 8192000 --only spectrum1024_db
python tools/kbench.py --streams 256 --samples 8192000 --only spectrum4096_db
python tools/kbench.py --streams 256 --samples 8192000 --only chain_fused
