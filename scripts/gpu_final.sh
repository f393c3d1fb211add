mkdir -p gpurun_out
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python scripts/show_bench.py < gpurun_out/bench.json 2>/dev/null | head -1
bash scripts/gpu_traffic.sh
