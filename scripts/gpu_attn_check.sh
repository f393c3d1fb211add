mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attn or philox" > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"
tail -n 3 gpurun_out/t_kernels.log
ATTN_ORDER=1 timeout 200 python scripts/attn_micro.py 2>&1 | tail -3
ATTN_ORDER=0 timeout 200 python scripts/attn_micro.py 2>&1 | tail -3
