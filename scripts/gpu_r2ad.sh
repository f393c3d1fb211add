mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "dense_fwd_fused or tail_fused" > gpurun_out/t_ad.log 2>&1; echo "dense tests rc=$?"
tail -n 8 gpurun_out/t_ad.log
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_all_ad.log 2>&1; echo "all rc=$?"
tail -n 8 gpurun_out/t_all_ad.log
for v in 1 0 1 0; do
ACSR_DENSE_FUSED=$v timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-long-seq --no-parity > gpurun_out/bench_c2_ad$v.json 2> gpurun_out/bench_c2_ad$v.err
echo "dense_fused=$v $(python scripts/show_bench.py < gpurun_out/bench_c2_ad$v.json 2>/dev/null | head -1)"
done
python scripts/show_bench.py < gpurun_out/bench_c2_ad1.json 2>/dev/null | grep -i "dense\|tail_fwd"
tail -n 3 gpurun_out/bench_c2_ad1.err
