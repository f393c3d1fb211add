mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x > gpurun_out/t_af.log 2>&1; echo "gemm tests rc=$?"
tail -n 5 gpurun_out/t_af.log
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_all_af.log 2>&1; echo "all rc=$?"
tail -n 5 gpurun_out/t_all_af.log
for i in 1 2; do
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-long-seq --no-parity > gpurun_out/bench_c2_af.json 2> gpurun_out/bench_c2_af.err
echo "$(python scripts/show_bench.py < gpurun_out/bench_c2_af.json 2>/dev/null | head -1)"
done
python scripts/show_bench.py < gpurun_out/bench_c2_af.json 2>/dev/null | grep gemm_batch
timeout 600 python bench.py --workload c5 --steps 5 --warmup 2 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-parity > gpurun_out/bench_c5_af.json 2> gpurun_out/bench_c5_af.err; echo "bench c5 rc=$?"
python scripts/show_bench.py < gpurun_out/bench_c5_af.json 2>/dev/null | head -3
timeout 600 python bench.py --workload c3v --steps 20 --warmup 3 --no-cpu-baseline --no-large-batch --no-vocab-sharded --no-parity > gpurun_out/bench_c3v_af.json 2> /dev/null; echo "bench c3v rc=$?"
python scripts/show_bench.py < gpurun_out/bench_c3v_af.json 2>/dev/null | head -2
