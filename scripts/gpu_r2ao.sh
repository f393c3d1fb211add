mkdir -p gpurun_out
cd tests
timeout 900 python -m pytest test_gpu_model.py -x -q -m gpu -k "sibling or ssept_trainer or ti_trainer or bert_trainer" 2>&1 | tail -25
cd ..
timeout 900 python bench.py --no-long-seq --no-vocab-sharded --no-large-batch > gpurun_out/bench_ao.json 2> gpurun_out/bench_ao.err; echo "bench rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/bench_ao.json'):
    if l.startswith('{'):
        j = json.loads(l)
        print('train', j['value'], 'eval', j['eval']['value'])
        for k, v in j.get('sibling_models', {}).items():
            print(k, v if isinstance(v, str) else {a: v[a] for a in v if a != 'config'})
P
tail -n 5 gpurun_out/bench_ao.err
