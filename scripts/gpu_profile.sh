# ncu evidence for profiles/: launch list of the step, DRAM traffic per launch, one full capture of the top kernels.
mkdir -p gpurun_out
CMD="python bench.py --profile --steps 2 --warmup 1"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed.avg.per_cycle_elapsed --clock-control none -k regex:"acsr" -s 80 -c 90 --csv --log-file gpurun_out/traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
echo "traffic rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_kernel|attn_fwd_kernel" -s 4 -c 4 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_full.log 2>&1
echo "full attn rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"linear_tok_kernel" -s 25 -c 3 -o gpurun_out/prof_tok $CMD > gpurun_out/ncu_full2.log 2>&1
echo "full tok rc=$?"
ls -la gpurun_out | head -30
