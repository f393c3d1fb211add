N=${1:-4}
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_r02_n${N}_full.json 2> gpurun_out/bench_r02_n${N}_full.err; echo "bench rc=$? wall=$(( $(date +%s) - S ))s"
tail -n 2 gpurun_out/bench_r02_n${N}_full.err; python scripts/show_bench.py < gpurun_out/bench_r02_n${N}_full.json 2>/dev/null | head -1
python - <<P
import json
for l in open('gpurun_out/bench_r02_n${N}_full.json'):
    if l.startswith('{'):
        j=json.loads(l)
        for k in ('vocab_sharded','long_sequence'):
            v=j.get(k); print(k, json.dumps({kk:v[kk] for kk in v if kk in ('value','ms_per_step','eval','error','n_gpus','parallelism','storage')}) if v else None)
P
