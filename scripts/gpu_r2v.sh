mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_all_v.log 2>&1; echo "all rc=$?"
tail -n 6 gpurun_out/t_all_v.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-long-seq > gpurun_out/bench_c2_v.json 2> gpurun_out/bench_c2_v.err; echo "bench rc=$?"
tail -n 2 gpurun_out/bench_c2_v.err
python scripts/show_bench.py < gpurun_out/bench_c2_v.json 2>/dev/null | head -4
python - <<'P'
import json
for l in open('gpurun_out/bench_c2_v.json'):
    if l.startswith('{'):
        j=json.loads(l); print(j['e2e_device_resident'])
P
ACSR_FOLD_GATE=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-long-seq --no-parity > gpurun_out/bench_c2_v0.json 2> /dev/null; echo "bench nofold rc=$?"
python scripts/show_bench.py < gpurun_out/bench_c2_v0.json 2>/dev/null | head -1
