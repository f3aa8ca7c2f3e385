# round 2: whole GPU suite on the current build, smoke, then the launch list of the bench command (kernels of this library only)
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02/pytest_all.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
CMD="python bench.py --steps 2 --warmup 3 --launches-per-step 8 --no-e2e --no-cpu-baseline --no-configs"
$CMD > gpurun_out/r02/plain.json 2> gpurun_out/r02/plain.err || { echo "plain run failed"; tail -5 gpurun_out/r02/plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chain_|spectrum|fm_|wire|nccl|ather" -c 600 --csv --log-file gpurun_out/r02/launches.csv $CMD > gpurun_out/r02/ncu_ll.log 2>&1; echo "ll rc=$?"
grep -c chain_jobs gpurun_out/r02/launches.csv
