mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -x -k "select or topk" > gpurun_out/t_sel.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/t_sel.log
ncu --set full --import-source on --clock-control none -k regex:attn_long_bwd_rows -s 1 -c 1 -o gpurun_out/prof_long_rows -f python bench.py --workload c5 --batch 256 --profile --steps 1 --warmup 0 > gpurun_out/ncu_long.log 2>&1; echo "ncu rc=$?"
tail -n 3 gpurun_out/ncu_long.log
ncu -i gpurun_out/prof_long_rows.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_long_rows_src.csv 2>/dev/null
python scripts/ncu_hot_lines.py gpurun_out/prof_long_rows_src.csv 40 x
python scripts/ncu_summary.py gpurun_out/prof_long_rows.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_long -c 12 --csv --log-file gpurun_out/long_list.csv python bench.py --workload c5 --batch 256 --profile --steps 1 --warmup 0 > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/long_list.csv
