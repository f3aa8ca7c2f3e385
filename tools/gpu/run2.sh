mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chain_fused|spectrum|fm_|wire" -c 400 --csv --log-file gpurun_out/r01b_launches.csv $CMD > gpurun_out/r01b_ncu_ll.log 2>&1
grep -c chain_fused gpurun_out/r01b_launches.csv
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_spectrum.py tests/test_gpu_wire.py -x -q -k "other_frame_lengths or goldens_n4096 or spectrum_messages or audio_messages or test_n1024_per_frame" > gpurun_out/s7_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/s7_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_spectrum.py -x -q -k "other_frame_lengths and (2048 or 4096-1 or 8192)" > gpurun_out/s7_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -8 gpurun_out/s7_racecheck.log
