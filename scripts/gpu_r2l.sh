mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x > gpurun_out/t_k1.log 2>&1; echo "kernel tests rc=$?"
tail -n 8 gpurun_out/t_k1.log
timeout 300 python scripts/ce_bwd_micro.py > gpurun_out/ce_micro.log 2>&1; echo "micro rc=$?"; cat gpurun_out/ce_micro.log | tail -12
timeout 300 python scripts/ce_bwd_micro.py 12102 512 > gpurun_out/ce_micro2.log 2>&1; echo "micro rc=$?"; cat gpurun_out/ce_micro2.log | tail -12
