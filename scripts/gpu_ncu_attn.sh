mkdir -p gpurun_out
CMD="python bench.py --profile --steps 1 --warmup 1"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"attn_(fwd|bwd)_kernel" -s 2 -c 4 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
