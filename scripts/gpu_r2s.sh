mkdir -p gpurun_out
python scripts/umma_rate.py 2>&1 | tail -10
timeout 300 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "device_resident" 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-parity --no-long-seq > gpurun_out/bench_c2_s.json 2> gpurun_out/bench_c2_s.err; echo "bench rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/bench_c2_s.json'):
    if l.startswith('{'):
        j=json.loads(l)
        print(j['value'], j['ms_per_step'], j['e2e'], j['e2e_device_resident']['value'])
P
