# Round profile: plain bench first (must exit 0), then the ncu launch list and one full capture of the
# dominant kernel for the SAME command, then timings and DRAM traffic of every kernel of tools/kbench.py
# and full captures of the 4096- and 65536-point kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/r01b_plain.json 2> gpurun_out/r01b_plain.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chain_fused|spectrum|fm_|wire" -c 400 --csv --log-file gpurun_out/r01b_launches.csv $CMD > gpurun_out/r01b_ncu_ll.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_fused -s 3 -c 1 -o gpurun_out/r01b_chain_fused -f $CMD > gpurun_out/r01b_ncu_full.log 2>&1
python tools/kbench.py --streams 256 > gpurun_out/r01b_kbench.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"chain_fused|spectrum|fm_" --csv --log-file gpurun_out/r01b_kbench_traffic.csv python tools/kbench.py --streams 256 --reps 1 > gpurun_out/r01b_kbench_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spectrum4096 -s 3 -c 1 -o gpurun_out/r01b_spectrum4096 -f python tools/kbench.py --only spectrum4096_db --streams 256 --reps 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:spectrum64k -s 3 -c 1 -o gpurun_out/r01b_spectrum64k -f python tools/kbench.py --only spectrum65536_hann_50pct --streams 256 --reps 2 > /dev/null 2>&1
python bench.py > gpurun_out/r01b_bench_n1.json 2> gpurun_out/r01b_bench_n1.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/r01b_bench_n1.json; cat gpurun_out/r01b_kbench.log
