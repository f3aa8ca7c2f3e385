python tools/kbench.py --only chain_fused --streams 256 --samples 8192000 --reps 100
python tools/kbench.py --only chain_fused --streams 256 --samples 8192000 --reps 300
python bench.py --steps 200 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench200', d['value'], d['roofline']['kernel_ms'], d['clocks'])"
