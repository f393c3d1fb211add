mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "tail_fused" > gpurun_out/t_m1.log 2>&1; echo "tail tests rc=$?"
tail -n 12 gpurun_out/t_m1.log
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/t_all_m.log 2>&1; echo "all rc=$?"
tail -n 6 gpurun_out/t_all_m.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-large-batch --no-vocab-sharded > gpurun_out/bench_c2_m.json 2> gpurun_out/bench_c2_m.err; echo "bench rc=$?"
tail -n 3 gpurun_out/bench_c2_m.err
python scripts/show_bench.py < gpurun_out/bench_c2_m.json 2>/dev/null | head -24
ACSR_TAIL_FUSED=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-parity > gpurun_out/bench_c2_m0.json 2> /dev/null; echo "bench old rc=$?"
python scripts/show_bench.py < gpurun_out/bench_c2_m0.json 2>/dev/null | head -1
