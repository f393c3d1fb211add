mkdir -p gpurun_out
timeout 200 python scripts/dense_micro.py 2>&1 | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dense_fwd_kernel" -c 1 -o gpurun_out/r02_prof_dense python scripts/dense_micro.py > gpurun_out/ncu_dense.log 2>&1; echo "ncu rc=$?"
