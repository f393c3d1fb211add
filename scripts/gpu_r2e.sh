mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"
tail -n 6 gpurun_out/t_all.log
timeout 900 python bench.py --no-large-batch > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
tail -n 3 gpurun_out/bench_c2.err
python scripts/show_bench.py < gpurun_out/bench_c2.json 2>/dev/null | head -3
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1])
for k in ('e2e','e2e_device_resident','vocab_sharded'):
    print(k, json.dumps(j.get(k))[:500])
PY
