mkdir -p gpurun_out
for i in 1 2; do
for v in 0 40 64 100; do
ACSR_WGRAD_CTAS=$v timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-long-seq --no-parity > gpurun_out/bench_c2_ag$v.json 2> /dev/null
echo "wgrad_ctas=$v $(python scripts/show_bench.py < gpurun_out/bench_c2_ag$v.json 2>/dev/null | head -1)"
done
done
