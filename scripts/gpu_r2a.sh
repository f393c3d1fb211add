mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?"
tail -n 4 gpurun_out/t_gemm.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"
tail -n 4 gpurun_out/t_kernels.log
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu > gpurun_out/t_model.log 2>&1; echo "model rc=$?"
tail -n 15 gpurun_out/t_model.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
python scripts/show_bench.py < gpurun_out/bench_c2.json 2>/dev/null | head -40
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --workload c3v > gpurun_out/bench_c3v.json 2> gpurun_out/bench_c3v.err; echo "bench c3v rc=$?"
python scripts/show_bench.py < gpurun_out/bench_c3v.json 2>/dev/null | head -30
