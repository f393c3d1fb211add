mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -x > gpurun_out/t_model.log 2>&1; echo "model rc=$?" >> gpurun_out/summary.txt
for pdl in 1 0; do
ACSR_PDL=$pdl timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err; echo "pdl=$pdl rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
tail -n 4 gpurun_out/t_kernels.log gpurun_out/t_model.log
