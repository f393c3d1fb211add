# final validation of the round: whole GPU suite, smoke, default bench (+ reference arm)
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4; echo "tests wall=$(( $(date +%s) - S ))s"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo "bench default rc=$? wall=$(( $(date +%s) - S ))s"
tail -n 2 gpurun_out/bench_r02_final.err
python scripts/show_bench.py < gpurun_out/bench_r02_final.json 2>/dev/null | head -8
python - <<'P'
import json
for l in open('gpurun_out/bench_r02_final.json'):
    if l.startswith('{'):
        j=json.loads(l)
        for k in ('long_sequence','vocab_sharded','large_batch','cpu_baseline','parity','e2e','e2e_device_resident','sibling_models'):
            print(k, json.dumps(j.get(k))[:400])
P
timeout 300 python bench.py --impl reference > gpurun_out/bench_r02_final_ref.json 2> /dev/null; echo "ref rc=$?"; head -c 300 gpurun_out/bench_r02_final_ref.json
