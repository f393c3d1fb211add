mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --tb=short -x > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"; tail -n 5 gpurun_out/t_all.log
for mt in 1 8; do echo "== ACSR_TOPK_MIN_TILES=$mt"; ACSR_TOPK_MIN_TILES=$mt timeout 200 python scripts/topk_micro.py 2>&1 | tail -4; done
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -n 3 gpurun_out/bench.err; python scripts/show_bench.py < gpurun_out/bench.json 2>/dev/null | head -2
