# what the driver does at round end, in one go
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2>/dev/null; echo "ref rc=$?"; cut -c1-160 gpurun_out/final_ref.json
python bench.py > gpurun_out/final_bench.json 2>/dev/null; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
