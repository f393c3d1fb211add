mkdir -p gpurun_out
for v in 1 2 4; do
  ACSR_STEP_BRANCHES=$v timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_br$v.json 2> gpurun_out/bench_br$v.err; echo "BRANCHES=$v rc=$?"
  python scripts/show_bench.py < gpurun_out/bench_br$v.json 2>/dev/null | head -1
done
