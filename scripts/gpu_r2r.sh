mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$? wall=$(( $(date +%s) - S ))s"
tail -n 3 gpurun_out/bench_default.err
python scripts/show_bench.py < gpurun_out/bench_default.json 2>/dev/null | head -3
python - <<'P'
import json
for l in open('gpurun_out/bench_default.json'):
    if l.startswith('{'):
        j=json.loads(l)
        for k in ('long_sequence','vocab_sharded','large_batch','cpu_baseline','eager_cuda_baseline','parity','e2e','e2e_device_resident','clocks'):
            print(k, json.dumps(j.get(k))[:700])
P
S=$(date +%s)
timeout 600 python bench.py --impl reference > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err; echo "ref default rc=$? wall=$(( $(date +%s) - S ))s"
tail -c 400 gpurun_out/bench_default_ref.json
