mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --tb=short -x > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"
tail -n 12 gpurun_out/t_all.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -n 3 gpurun_out/bench.err; python scripts/show_bench.py < gpurun_out/bench.json | head -8
timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
python scripts/show_bench.py < gpurun_out/bench_c4.json | head -3
