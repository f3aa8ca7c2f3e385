# what the driver does at round end, in one go (1 GPU): GPU suite, smoke, reference arm, B200 arm
mkdir -p gpurun_out/r02
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02/final_ref.json 2>/dev/null; echo "ref rc=$?"; cut -c1-160 gpurun_out/r02/final_ref.json
S=$(date +%s); python bench.py > gpurun_out/r02/final_bench.json 2> gpurun_out/r02/final_bench.err; echo "bench rc=$? wall=$(( $(date +%s) - S )) s"; python -c "
import json; d=json.loads(open('gpurun_out/r02/final_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['roofline']['burst']['gsamples_per_s'], d['e2e']['value'], d['e2e_products']['value'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks']); print([(c['name'], round(c['value']/1e3,1), round(c['roofline']['frac'],3)) for c in d['configs']])"
