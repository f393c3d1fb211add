N=2
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -n 2 gpurun_out/bench_n$N.err; python scripts/show_bench.py < gpurun_out/bench_n$N.json 2>/dev/null | head -1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --workload c4 --no-cpu-baseline > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo "bench c4 rc=$?"
tail -n 2 gpurun_out/bench_c4_n$N.err; python scripts/show_bench.py < gpurun_out/bench_c4_n$N.json 2>/dev/null | head -1
