set -x; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest.log
python tools/kbench.py > gpurun_out/s2_kbench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spectrum_mx1024 -s 3 -c 1 -o gpurun_out/s2_prof_4k -f python tools/kbench.py --only spectrum4096_db --reps 2 > gpurun_out/s2_ncu_4k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spectrum64k -s 3 -c 1 -o gpurun_out/s2_prof_64k -f python tools/kbench.py --only spectrum65536_hann_50pct --reps 2 > gpurun_out/s2_ncu_64k.log 2>&1
tail -3 gpurun_out/s2_pytest.log; cat gpurun_out/s2_kbench.log
