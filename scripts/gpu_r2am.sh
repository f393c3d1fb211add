# ACTiSASRec: time-aware pair kernels + the attention entry points around them, model tests, headline regression check
mkdir -p gpurun_out
cd tests
timeout 900 python -m pytest test_gpu_kernels.py -x -q -m gpu -k "time_aware" 2>&1 | tail -15
timeout 900 python -m pytest test_gpu_model.py -x -q -m gpu -k "ti_" 2>&1 | tail -25
cd ..


