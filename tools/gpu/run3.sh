mkdir -p gpurun_out
python -m pytest tests/test_gpu_spectrum.py -x -q -k "parseval" 2>&1 | grep -E "assert|Error|rel|passed|failed" | head -20
