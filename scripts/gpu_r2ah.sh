mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "topk" > gpurun_out/t_ah.log 2>&1; echo "topk tests rc=$?"
tail -n 5 gpurun_out/t_ah.log
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "full_size or properties or eval" > gpurun_out/t_ah2.log 2>&1; echo "model eval tests rc=$?"
tail -n 4 gpurun_out/t_ah2.log
timeout 300 python scripts/topk_micro.py 2>&1 | tail -5
