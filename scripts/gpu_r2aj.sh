mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_all_aj.log 2>&1; echo "all rc=$?"
tail -n 4 gpurun_out/t_all_aj.log
timeout 300 python scripts/topk_micro.py 2>&1 | tail -4
timeout 300 python bench.py --workload c4 --steps 30 --warmup 5 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-parity > gpurun_out/bench_c4_aj.json 2> gpurun_out/bench_c4_aj.err; echo "bench c4 rc=$?"
python scripts/show_bench.py < gpurun_out/bench_c4_aj.json 2>/dev/null | head -1
