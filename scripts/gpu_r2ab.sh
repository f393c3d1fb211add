mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "ce_backward_without or logits_ce" > gpurun_out/t_ab.log 2>&1; echo "ce tests rc=$?"
tail -n 3 gpurun_out/t_ab.log
timeout 300 python scripts/ce_bwd_micro.py > gpurun_out/ce_micro.log 2>&1; echo "micro rc=$?"; cat gpurun_out/ce_micro.log | tail -8
timeout 300 python scripts/ce_bwd_micro.py 12102 512 2>&1 | tail -4
timeout 300 python bench.py --workload c4 --steps 30 --warmup 5 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-parity > gpurun_out/bench_c4_ab.json 2> gpurun_out/bench_c4_ab.err; echo "bench c4 rc=$?"
python scripts/show_bench.py < gpurun_out/bench_c4_ab.json 2>/dev/null | head -6
