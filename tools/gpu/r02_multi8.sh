# the multi-GPU tests that skip below 8 GPUs, and the C example over 8 devices
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_audio.py -x -q -k "comm or multi" -rs 2>&1 | tail -4
gcc -O2 -std=c99 -Iinclude examples/multi_gpu_gather.c -o /tmp/mgg -Lrtl-ws_b200 -lb200sdr -Wl,-rpath,$PWD/rtl-ws_b200 -lm && timeout 300 /tmp/mgg 2>&1 | tail -3
