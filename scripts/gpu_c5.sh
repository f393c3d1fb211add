mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --tb=short -x > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"; tail -n 5 gpurun_out/t_all.log
timeout 600 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"
tail -n 3 gpurun_out/bench_c5.err; python scripts/show_bench.py < gpurun_out/bench_c5.json 2>/dev/null | head -9
timeout 900 python bench.py --workload c5 --full-len --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_full.json 2> gpurun_out/bench_c5_full.err; echo "bench c5 full rc=$?"
tail -n 3 gpurun_out/bench_c5_full.err; python scripts/show_bench.py < gpurun_out/bench_c5_full.json 2>/dev/null | head -9
