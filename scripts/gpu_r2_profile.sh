# ncu evidence for profiles/ (round 2, final kernels): launch list of the step, DRAM traffic per launch, full captures of the top kernels.
mkdir -p gpurun_out
CMD="python bench.py --profile --steps 2 --warmup 1"
$CMD > gpurun_out/plain.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 400 --csv --log-file gpurun_out/r02_traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
echo "traffic rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"linear_tok_kernel" -s 25 -c 3 -o gpurun_out/r02_prof_tok $CMD > gpurun_out/ncu_full1.log 2>&1
echo "full tok rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"ce_dout_kernel|ce_dtable_kernel|tail_fwd_kernel|tail_bwd_kernel" -s 4 -c 4 -o gpurun_out/r02_prof_ce $CMD > gpurun_out/ncu_full2.log 2>&1
echo "full ce/tail rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"ce_dout_kernel|logits_tc_kernel" -c 3 -o gpurun_out/r02_prof_ce1m python scripts/ce_bwd_micro.py > gpurun_out/ncu_full3.log 2>&1
echo "full ce 1M rc=$?"
ls -la gpurun_out | grep r02_
