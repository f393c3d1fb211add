mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "all rc=$?"
tail -n 12 gpurun_out/t_all.log
for w in c3v c5; do
timeout 900 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
tail -n 2 gpurun_out/bench_$w.err
python scripts/show_bench.py < gpurun_out/bench_$w.json 2>/dev/null | head -8
done
