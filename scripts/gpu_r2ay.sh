mkdir -p gpurun_out
cd tests
timeout 1500 python -m pytest test_gpu_kernels.py test_gpu_model.py -x -q -m gpu 2>&1 | tail -3
cd ..
timeout 600 python bench.py --steps 50 --warmup 5 --no-long-seq --no-vocab-sharded --no-large-batch --no-cpu-baseline > gpurun_out/bench_ay.json 2> gpurun_out/bench_ay.err; echo "bench rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/bench_ay.json'):
    if l.startswith('{'):
        j = json.loads(l)
        print('train', j['value'], 'eval', j['eval']['value'])
        for k, v in j.get('sibling_models', {}).items():
            print(k, v if isinstance(v, str) else {a: v[a] for a in v if a not in ('config', 'cpu_baseline')})
P
