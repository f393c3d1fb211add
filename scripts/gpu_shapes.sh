# model-level parity on the config shapes + bench lines of the non-headline workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --tb=short -x > gpurun_out/t_model.log 2>&1; echo "model rc=$?"
tail -n 25 gpurun_out/t_model.log
for w in c3 c3v c4; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
  tail -n 3 gpurun_out/bench_$w.err; python scripts/show_bench.py < gpurun_out/bench_$w.json 2>/dev/null | head -12
done
timeout 600 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"
tail -n 5 gpurun_out/bench_c5.err; python scripts/show_bench.py < gpurun_out/bench_c5.json 2>/dev/null | head -30
nvidia-smi --query-gpu=memory.used --format=csv
