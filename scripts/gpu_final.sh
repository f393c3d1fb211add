mkdir -p gpurun_out
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python scripts/show_bench.py < gpurun_out/bench.json 2>/dev/null | head -3
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
for w in c3 c3v c4; do timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"; python scripts/show_bench.py < gpurun_out/bench_$w.json 2>/dev/null | head -1; done
bash scripts/gpu_traffic.sh
