mkdir -p gpurun_out
python -m pytest tests/test_gpu_spectrum.py -x -q -k "65536" 2>&1 | tail -3
python tools/kbench.py --only spectrum65536_hann_50pct --streams 256
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:spectrum64k -s 3 -c 1 python tools/kbench.py --only spectrum65536_hann_50pct --streams 256 --reps 1 2>&1 | grep -E "dram__|gpu__time"
