# usage: gpurun --gpus N -- 'bash scripts/gpu_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1; echo "check rc=$?"
tail -n 3 gpurun_out/multi_check.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -n 3 gpurun_out/bench_n$N.err; head -c 300 gpurun_out/bench_n$N.json; echo; python scripts/show_bench.py < gpurun_out/bench_n$N.json | head -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --workload c4 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo "bench c4 rc=$?"
tail -n 4 gpurun_out/bench_c4_n$N.err; python scripts/show_bench.py < gpurun_out/bench_c4_n$N.json | head -6
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_n$N.json | head -c 400
