# usage: bash tools/gpu/scale.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/r01b_bench_n$N.json 2> gpurun_out/r01b_bench_n$N.err
echo "rc=$?"; cut -c1-260 gpurun_out/r01b_bench_n$N.json; tail -3 gpurun_out/r01b_bench_n$N.err
