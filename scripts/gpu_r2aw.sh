cd tests
for i in 1 2 3; do timeout 600 python -m pytest test_gpu_model.py -x -q -m gpu 2>&1 | grep -E "^E  |passed|failed" | head -6; done
timeout 900 python -m pytest test_gpu_kernels.py test_gpu_headline.py test_gpu_gemm.py -x -q -m gpu 2>&1 | tail -2
