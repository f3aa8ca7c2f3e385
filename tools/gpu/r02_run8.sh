mkdir -p gpurun_out/r02
for i in 1 2; do timeout 200 python tools/kbench.py --streams 256 --only spectrum65536_hann_50pct 2>&1 | tail -1; done
B200_S64K_PERSIST=1 timeout 200 python tools/kbench.py --streams 256 --only spectrum65536_hann_50pct 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:spectrum64k -s 3 -c 1 --csv --log-file gpurun_out/r02/s64k_traffic_nopersist.csv python tools/kbench.py --only spectrum65536_hann_50pct --streams 256 --reps 2 > /dev/null 2>&1; tail -3 gpurun_out/r02/s64k_traffic_nopersist.csv | cut -d, -f13-
