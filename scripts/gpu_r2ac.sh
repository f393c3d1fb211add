mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "dense_fwd_fused" > gpurun_out/t_ac.log 2>&1; echo "dense tests rc=$?"
tail -n 15 gpurun_out/t_ac.log
