mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -x -k "ce_ or edge" > gpurun_out/t_ce.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/t_ce.log
for v in 1 0 1 0; do
  ACSR_ORDER_SIDE=$v timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_os$v.json 2> gpurun_out/bench_os$v.err; echo "ORDER_SIDE=$v rc=$?"
  python scripts/show_bench.py < gpurun_out/bench_os$v.json 2>/dev/null | head -1
done
