mkdir -p gpurun_out
ACSR_BENCH_CALLS=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vocab-sharded --no-large-batch --no-parity > gpurun_out/bench_c2_n.json 2> gpurun_out/bench_c2_n.err; echo "bench rc=$?"
grep CALL gpurun_out/bench_c2_n.err
for w in c1 c3 c3v c4 c5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-parity > gpurun_out/bench_r02b_$w.json 2> gpurun_out/bench_r02b_$w.err; echo "bench $w rc=$?"
  python scripts/show_bench.py < gpurun_out/bench_r02b_$w.json 2>/dev/null | head -8
done
