N=2
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1; echo "multi_gpu_check rc=$?"
tail -n 3 gpurun_out/multi_check.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 50 --warmup 5 --no-long-seq > gpurun_out/bench_r02_n2_c.json 2> gpurun_out/bench_r02_n2_c.err; echo "bench rc=$?"
python scripts/show_bench.py < gpurun_out/bench_r02_n2_c.json 2>/dev/null | head -1
python - <<'P'
import json
for l in open('gpurun_out/bench_r02_n2_c.json'):
    if l.startswith('{'):
        j=json.loads(l); v=j.get('vocab_sharded'); print({k:v[k] for k in ('value','ms_per_step','eval')})
P
