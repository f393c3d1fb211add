mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"
tail -n 4 gpurun_out/t_all.log
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
tail -n 3 gpurun_out/bench_c2.err
python scripts/show_bench.py < gpurun_out/bench_c2.json 2>/dev/null | head -24
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1])
for k in ('parity','large_batch','cpu_baseline','cpu_baseline_eval','eager_cuda_baseline','roofline'):
    print(k, json.dumps(j.get(k))[:700])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc=$?"; head -c 600 gpurun_out/bench_ref.json; echo
