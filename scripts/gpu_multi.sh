# usage: gpurun --gpus N -- 'bash scripts/gpu_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1; echo "check rc=$?"
tail -n 6 gpurun_out/multi_check.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -n 3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
