# Round-2 profile: plain bench first (must exit 0), then the ncu launch list and one full capture of the dominant
# kernel for the SAME command; then timings + DRAM traffic of every kernel of tools/kbench.py and full captures of
# the config 2 / 3 kernels.
mkdir -p gpurun_out/r02
CMD="python bench.py --steps 2 --warmup 3 --launches-per-step 8 --no-e2e --no-cpu-baseline --no-configs"
$CMD > gpurun_out/r02/plain.json 2> gpurun_out/r02/plain.err || { echo "plain run failed"; tail -5 gpurun_out/r02/plain.err; exit 1; }
cut -c1-200 gpurun_out/r02/plain.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chain_|spectrum|fm_|wire|nccl|Gather|gather" -c 600 --csv --log-file gpurun_out/r02/launches.csv $CMD > gpurun_out/r02/ncu_ll.log 2>&1; echo "ll rc=$?"
ncu --set full --clock-control none --import-source on -k regex:chain_jobs -s 5 -c 1 -o gpurun_out/r02/chain_jobs_bench -f $CMD > gpurun_out/r02/ncu_full.log 2>&1; echo "full rc=$?"
python tools/kbench.py --streams 256 > gpurun_out/r02/kbench.log 2>&1; cat gpurun_out/r02/kbench.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"chain_|spectrum|fm_" --csv --log-file gpurun_out/r02/kbench_traffic.csv python tools/kbench.py --streams 256 --reps 1 > gpurun_out/r02/kbench_ncu.log 2>&1; echo "traffic rc=$?"
ncu --set full --clock-control none --import-source on -k regex:spectrum4096 -s 3 -c 1 -o gpurun_out/r02/spectrum4096 -f python tools/kbench.py --only spectrum4096_db --streams 256 --reps 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:spectrum1024_tiled -s 3 -c 1 -o gpurun_out/r02/spectrum1024_tiled -f python tools/kbench.py --only spectrum1024_db --streams 256 --reps 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fm_chain_warp -s 3 -c 1 -o gpurun_out/r02/fm_chain_warp -f python tools/kbench.py --only fm_chain --streams 256 --reps 2 > /dev/null 2>&1
ls -la gpurun_out/r02/
