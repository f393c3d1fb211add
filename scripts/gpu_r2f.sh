mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "bidirectional or gather_rows" > gpurun_out/t_bidir.log 2>&1; echo "bidir rc=$?"
tail -n 25 gpurun_out/t_bidir.log
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -k "bert" > gpurun_out/t_bert.log 2>&1; echo "bert rc=$?"
tail -n 25 gpurun_out/t_bert.log
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "all rc=$?"
tail -n 4 gpurun_out/t_all.log
