mkdir -p gpurun_out
for i in 1 2; do
for v in 1 3; do
ACSR_SPLIT_WGRAD=$v timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-long-seq --no-parity > gpurun_out/bench_c2_z$v.json 2> /dev/null
echo "split=$v $(python scripts/show_bench.py < gpurun_out/bench_c2_z$v.json 2>/dev/null | head -1)"
done
done
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_headline.py -q -m gpu -x > gpurun_out/t_z.log 2>&1; echo "model tests rc=$?"
tail -n 3 gpurun_out/t_z.log
