mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --tb=short -x -k "topk or select or million or golden_eval or config_shapes or full_size" > gpurun_out/t_sel.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/t_sel.log | grep -v Warn
timeout 100 python scripts/topk_micro.py 2>&1 | tail -4
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python scripts/show_bench.py < gpurun_out/bench.json 2>/dev/null | head -1
