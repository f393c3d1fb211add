mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "ce_backward_fused" > gpurun_out/t_cebwd.log 2>&1; echo "cebwd rc=$?"
tail -n 30 gpurun_out/t_cebwd.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_headline.py -q -m gpu -x > gpurun_out/t_model.log 2>&1; echo "model rc=$?"
tail -n 8 gpurun_out/t_model.log
timeout 900 python bench.py --no-large-batch > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
tail -n 3 gpurun_out/bench_c2.err
python scripts/show_bench.py < gpurun_out/bench_c2.json 2>/dev/null | head -26
timeout 600 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
tail -n 3 gpurun_out/bench_c4.err
python scripts/show_bench.py < gpurun_out/bench_c4.json 2>/dev/null | head -12
