mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_headline.py -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"
tail -n 6 gpurun_out/t_all.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-large-batch --no-vocab-sharded > gpurun_out/bench_c2_xs.json 2> gpurun_out/bench_c2_xs.err; echo "bench rc=$?"
tail -n 3 gpurun_out/bench_c2_xs.err
python scripts/show_bench.py < gpurun_out/bench_c2_xs.json 2>/dev/null | head -8
ACSR_LINEAR_XS=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-parity > gpurun_out/bench_c2_noxs.json 2> /dev/null; echo "bench noxs rc=$?"
python scripts/show_bench.py < gpurun_out/bench_c2_noxs.json 2>/dev/null | head -3
