mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?"
tail -n 25 gpurun_out/t_gemm.log
timeout 300 python scripts/gemm_ks_micro.py > gpurun_out/gemm_micro.log 2>&1; echo "micro rc=$?"
cat gpurun_out/gemm_micro.log | tail -n 30
