mkdir -p gpurun_out
for w in c3 c3v c1; do
timeout 400 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-parity > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
python scripts/show_bench.py < gpurun_out/bench_$w.json 2>/dev/null | head -8
done
timeout 900 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"
tail -n 3 gpurun_out/bench_c5.err
python scripts/show_bench.py < gpurun_out/bench_c5.json 2>/dev/null | head -14
for nb in 2 4; do
ACSR_STEP_BRANCHES=$nb timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-large-batch --no-vocab-sharded > gpurun_out/bench_c2_nb$nb.json 2> gpurun_out/bench_c2_nb$nb.err; echo "bench nb=$nb rc=$?"
python scripts/show_bench.py < gpurun_out/bench_c2_nb$nb.json 2>/dev/null | head -1
done
