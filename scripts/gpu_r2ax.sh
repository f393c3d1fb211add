mkdir -p gpurun_out
cd tests
timeout 900 python -m pytest test_gpu_kernels.py test_gpu_model.py -x -q -m gpu -k "time_aware or ti_ or sibling or direct" 2>&1 | tail -3
cd ..
python scripts/sibling_prof.py ACTiSASRec 2>&1 | grep -v CUDAEvent | head -9 | tee gpurun_out/sibling_prof_ti.txt
python scripts/sibling_prof.py ACSSEPT 2>&1 | grep -v CUDAEvent | head -9 | tee gpurun_out/sibling_prof_ssept.txt
