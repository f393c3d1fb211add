# ncu evidence for profiles/: launch list of the step + one full capture of the top kernels.
mkdir -p gpurun_out
CMD="python bench.py --profile --steps 2 --warmup 1"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"attn_(fwd|bwd)_kernel|linear_wgrad" -s 8 -c 10 -o gpurun_out/prof_top $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
ls -la gpurun_out
