mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --tb=short -x > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/t_all.log | grep -v Warn
for v in 1 0; do
  ACSR_FOLD_ATTACK=$v timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_fold$v.json 2> gpurun_out/bench_fold$v.err; echo "FOLD=$v rc=$?"
  tail -n 2 gpurun_out/bench_fold$v.err; python scripts/show_bench.py < gpurun_out/bench_fold$v.json 2>/dev/null | head -1
done
