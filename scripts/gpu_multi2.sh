# usage: gpurun --gpus N -- 'bash scripts/gpu_multi2.sh N'
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1; echo "check rc=$?"
tail -n 12 gpurun_out/multi_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -n 6 gpurun_out/bench_n$N.err; python scripts/show_bench.py < gpurun_out/bench_n$N.json | head -2
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
    print('vocab_sharded', json.dumps(j.get('vocab_sharded'))[:900])
except Exception as e:
    print('no json', e)
PY
