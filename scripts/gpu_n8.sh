N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -n 3 gpurun_out/bench_n$N.err; python scripts/show_bench.py < gpurun_out/bench_n$N.json 2>/dev/null | head -1
