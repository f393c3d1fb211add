# focused run of the newest kernels first (fail fast, short tracebacks), then the whole GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --tb=short -k "fp32_path or 200 or 80- or 130 or 65 or chunked or unsupported" > gpurun_out/t_new.log 2>&1; echo "new rc=$?"
tail -n 40 gpurun_out/t_new.log
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --tb=short > gpurun_out/t_all.log 2>&1; echo "all rc=$?"
tail -n 15 gpurun_out/t_all.log
