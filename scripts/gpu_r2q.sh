N=2
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1; echo "multi_gpu_check rc=$?"
tail -n 12 gpurun_out/multi_check.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err; echo "bench rc=$?"
tail -n 3 gpurun_out/bench_r02_n$N.err; python scripts/show_bench.py < gpurun_out/bench_r02_n$N.json 2>/dev/null | head -1
python - <<'P'
import json
for l in open('gpurun_out/bench_r02_n2.json'):
    if l.startswith('{'):
        j=json.loads(l); print(json.dumps(j.get('vocab_sharded'))[:1200])
P
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_r02_ref_n$N.json 2> gpurun_out/bench_r02_ref_n$N.err; echo "ref arm rc=$?"
tail -c 600 gpurun_out/bench_r02_ref_n$N.json
