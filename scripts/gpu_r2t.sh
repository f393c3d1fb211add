mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x > gpurun_out/t_t1.log 2>&1; echo "kernel tests rc=$?"
tail -n 5 gpurun_out/t_t1.log
timeout 300 python scripts/ce_bwd_micro.py > gpurun_out/ce_micro.log 2>&1; echo "micro rc=$?"; cat gpurun_out/ce_micro.log | tail -8
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-long-seq > gpurun_out/bench_c2_t.json 2> gpurun_out/bench_c2_t.err; echo "bench rc=$?"
tail -n 2 gpurun_out/bench_c2_t.err
python scripts/show_bench.py < gpurun_out/bench_c2_t.json 2>/dev/null | head -3
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-parity > gpurun_out/bench_c4_t.json 2> gpurun_out/bench_c4_t.err; echo "bench c4 rc=$?"
python scripts/show_bench.py < gpurun_out/bench_c4_t.json 2>/dev/null | head -8
