mkdir -p gpurun_out
cd tests
timeout 1200 python -m pytest test_gpu_model.py test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -5
cd ..
python scripts/sibling_prof.py ACTiSASRec 2>&1 | grep -v CUDAEvent | head -9 | tee gpurun_out/sibling_prof_ti.txt
python scripts/sibling_prof.py ACSSEPT 2>&1 | grep -v CUDAEvent | head -6 | tee gpurun_out/sibling_prof_ssept.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-long-seq --no-vocab-sharded --no-large-batch --no-cpu-baseline > gpurun_out/bench_aq.json 2> gpurun_out/bench_aq.err; echo "bench rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/bench_aq.json'):
    if l.startswith('{'):
        j = json.loads(l)
        print('train', j['value'], 'eval', j['eval']['value'])
        for k, v in j.get('sibling_models', {}).items():
            print(k, v if isinstance(v, str) else {a: v[a] for a in v if a not in ('config', 'cpu_baseline')})
P
