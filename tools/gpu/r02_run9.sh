mkdir -p gpurun_out/r02
timeout 600 python tools/gpu/chain_variants.py --variants 0,6,7,8,9,6,7 > gpurun_out/r02/variants2.log 2>&1; echo "variants rc=$?"
tail -6 gpurun_out/r02/variants2.log
for V in 8 9; do
  B200_CHAIN_VARIANT=$V timeout 300 python -m pytest tests/test_gpu_fm.py -q -x -k "chain" 2>&1 | tail -1
done
