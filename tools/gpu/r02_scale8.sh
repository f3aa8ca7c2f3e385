# the driver's 8-GPU commands, once: reference arm (rank 0 only works), then the B200 arm with default steps
mkdir -p gpurun_out/r02
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r02/ref_n8.json 2> gpurun_out/r02/ref_n8.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/r02/ref_n8.json
S=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02/bench_n8.json 2> gpurun_out/r02/bench_n8.err; echo "bench rc=$? wall=$(( $(date +%s) - S )) s"
tail -3 gpurun_out/r02/bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_n8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches','timed_region_s') if k in d})
print('gather', d.get('gather_check'), 'roofline', d['roofline']['frac'], 'burst', d['roofline'].get('burst',{}).get('gsamples_per_s'))
print('e2e', d.get('e2e',{}).get('value'), 'e2e_products', d.get('e2e_products',{}).get('value'), 'clocks', d.get('clocks'))
PY
