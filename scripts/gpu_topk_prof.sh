mkdir -p gpurun_out
python scripts/topk_prof.py 1000001 > /dev/null 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:logits_tc_kernel -s 1 -c 1 -o gpurun_out/prof_topk -f python scripts/topk_prof.py 1000001 > gpurun_out/ncu_topk.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_topk.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_topk_src.csv 2>/dev/null
python scripts/ncu_hot_lines.py gpurun_out/prof_topk_src.csv 30 x
python scripts/ncu_summary.py gpurun_out/prof_topk.ncu-rep
timeout 200 python scripts/topk_micro.py 2>&1 | tail -4
