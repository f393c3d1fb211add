N=${1:-8}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --workload c4 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo "bench c4 rc=$?"
tail -n 4 gpurun_out/bench_c4_n$N.err; python scripts/show_bench.py < gpurun_out/bench_c4_n$N.json 2>/dev/null | head -8
