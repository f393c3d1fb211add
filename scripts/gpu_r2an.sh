# default bench with the sibling-model record
mkdir -p gpurun_out
timeout 900 python bench.py --no-long-seq --no-vocab-sharded > gpurun_out/bench_an.json 2> gpurun_out/bench_an.err; echo "bench rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/bench_an.json'):
    if l.startswith('{'):
        j = json.loads(l)
        print('train', j['value'], 'eval', j['eval']['value'])
        print(json.dumps(j.get('sibling_models'), indent=1))
P
tail -n 5 gpurun_out/bench_an.err
