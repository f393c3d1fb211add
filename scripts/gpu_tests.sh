mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu > gpurun_out/t_model.log 2>&1; echo "model rc=$?" >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 3 gpurun_out/t_kernels.log gpurun_out/t_model.log gpurun_out/smoke.log
tail -n 5 gpurun_out/bench.err
if [ -n "$ATTN_MICRO" ]; then timeout 200 python scripts/attn_micro.py 2>&1 | tail -3; fi
