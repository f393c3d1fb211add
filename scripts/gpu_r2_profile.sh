# ncu evidence for profiles/ (round 2): launch list of the step, DRAM traffic per launch, full captures of the top kernels.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "all rc=$?"
tail -n 4 gpurun_out/t_all.log
CMD="python bench.py --profile --steps 2 --warmup 1"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed.avg.per_cycle_elapsed --clock-control none -k regex:"acsr" -s 70 -c 80 --csv --log-file gpurun_out/r02_traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
echo "traffic rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"linear_tok_kernel" -s 25 -c 3 -o gpurun_out/r02_prof_tok $CMD > gpurun_out/ncu_full1.log 2>&1
echo "full tok rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"gemm_ks_kernel" -s 3 -c 3 -o gpurun_out/r02_prof_gemm $CMD > gpurun_out/ncu_full2.log 2>&1
echo "full gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_kernel" -s 2 -c 2 -o gpurun_out/r02_prof_attn $CMD > gpurun_out/ncu_full3.log 2>&1
echo "full attn rc=$?"
ls -la gpurun_out | grep r02
