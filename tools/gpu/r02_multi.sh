# round 2: the multi-GPU entry points of the C ABI on a 2-GPU box -- tests that skip on one GPU, then the bench under torchrun
mkdir -p gpurun_out/r02
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_audio.py -x -q -k "comm or multi" -rs > gpurun_out/r02/pytest_multi.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r02/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --launches-per-step 16 --e2e-seconds 0.5 --no-cpu-baseline > gpurun_out/r02/bench_n2.json 2> gpurun_out/r02/bench_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/r02/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches') if k in d})
print('gather', {k:v for k,v in d.items() if 'gather' in k})
print('roofline', d['roofline']['frac'], 'e2e', d.get('e2e',{}).get('value'), 'e2e_products', d.get('e2e_products',{}).get('value'))
PY
gcc -O2 -std=c99 -Iinclude examples/multi_gpu_gather.c -o /tmp/mgg -Lrtl-ws_b200 -lb200sdr -Wl,-rpath,$PWD/rtl-ws_b200 -lm && timeout 300 /tmp/mgg 2>&1 | tail -5
