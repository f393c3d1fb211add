# --set full captures (after the plain run exited 0): the radix-select top-k and the tcgen05 token-tile GEMM of the current build
mkdir -p gpurun_out
python bench.py --profile --steps 1 --warmup 1 > gpurun_out/plain.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:topk_merge_kernel -c 1 -o gpurun_out/prof_select -f python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_full_select.log 2>&1; echo "ncu select rc=$?"
ncu --set full --import-source on --clock-control none -k regex:linear_tok_kernel -s 40 -c 3 -o gpurun_out/prof_tok2 -f python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_full_tok2.log 2>&1; echo "ncu tok rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_select.ncu-rep | head -30
ncu -i gpurun_out/prof_select.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_select_src.csv 2>/dev/null
python scripts/ncu_hot_lines.py gpurun_out/prof_select_src.csv 16 x
