mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_all_y.log 2>&1; echo "all rc=$?"
tail -n 4 gpurun_out/t_all_y.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-long-seq > gpurun_out/bench_c2_y.json 2> gpurun_out/bench_c2_y.err; echo "bench rc=$?"
tail -n 2 gpurun_out/bench_c2_y.err
python scripts/show_bench.py < gpurun_out/bench_c2_y.json 2>/dev/null | head -12
ACSR_FUSE_ACT_BWD=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-long-seq --no-parity > gpurun_out/bench_c2_y0.json 2> /dev/null; echo "bench nofuse rc=$?"
python scripts/show_bench.py < gpurun_out/bench_c2_y0.json 2>/dev/null | head -1
