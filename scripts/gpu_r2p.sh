mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py -q -m gpu -x > gpurun_out/t_p1.log 2>&1; echo "gemm+kernel tests rc=$?"
tail -n 5 gpurun_out/t_p1.log
for w in c3v c5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-parity > gpurun_out/bench_r02c_$w.json 2> gpurun_out/bench_r02c_$w.err; echo "bench $w rc=$?"
  python scripts/show_bench.py < gpurun_out/bench_r02c_$w.json 2>/dev/null | head -4
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-parity > gpurun_out/bench_c2_p.json 2> gpurun_out/bench_c2_p.err; echo "bench rc=$?"
python scripts/show_bench.py < gpurun_out/bench_c2_p.json 2>/dev/null | head -5
