mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --tb=short > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/t_all.log | grep -v Warn
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python scripts/show_bench.py < gpurun_out/bench.json 2>/dev/null | head -1
