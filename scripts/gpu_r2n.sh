mkdir -p gpurun_out
ACSR_BENCH_CALLS=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vocab-sharded --no-large-batch --no-parity > gpurun_out/bench_c2_n.json 2> gpurun_out/bench_c2_n.err; echo "bench rc=$?"
grep CALL gpurun_out/bench_c2_n.err
