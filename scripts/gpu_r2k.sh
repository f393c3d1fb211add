mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "ce_backward_without or logits_ce_forward" > gpurun_out/t_k1.log 2>&1; echo "ce tests rc=$?"
tail -n 15 gpurun_out/t_k1.log
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/t_all_k.log 2>&1; echo "all rc=$?"
tail -n 6 gpurun_out/t_all_k.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-large-batch --no-vocab-sharded > gpurun_out/bench_c2_k.json 2> gpurun_out/bench_c2_k.err; echo "bench rc=$?"
tail -n 3 gpurun_out/bench_c2_k.err
python scripts/show_bench.py < gpurun_out/bench_c2_k.json 2>/dev/null | head -12
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-parity > gpurun_out/bench_c4_k.json 2> gpurun_out/bench_c4_k.err; echo "bench c4 rc=$?"
python scripts/show_bench.py < gpurun_out/bench_c4_k.json 2>/dev/null | head -10
