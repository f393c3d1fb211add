mkdir -p gpurun_out
for i in 1 2; do
for v in 0 1; do
ACSR_EVAL_PDL=$v timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-long-seq --no-parity > gpurun_out/bench_c2_aa$v.json 2> /dev/null
echo "eval_pdl=$v $(python scripts/show_bench.py < gpurun_out/bench_c2_aa$v.json 2>/dev/null | head -1)"
done
done
