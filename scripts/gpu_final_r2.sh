mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo "bench default rc=$? wall=$(( $(date +%s) - S ))s"
tail -n 2 gpurun_out/bench_r02_final.err
python scripts/show_bench.py < gpurun_out/bench_r02_final.json 2>/dev/null | head -24
python - <<'P'
import json
for l in open('gpurun_out/bench_r02_final.json'):
    if l.startswith('{'):
        j=json.loads(l)
        for k in ('long_sequence','vocab_sharded','large_batch','cpu_baseline','cpu_baseline_eval','eager_cuda_baseline','parity','e2e','e2e_device_resident','roofline'):
            print(k, json.dumps(j.get(k))[:600])
P
for w in c1 c3 c3v c4; do
  timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-parity > gpurun_out/bench_r02_$w.json 2> gpurun_out/bench_r02_$w.err; echo "bench $w rc=$?"
  python scripts/show_bench.py < gpurun_out/bench_r02_$w.json 2>/dev/null | head -1
done
