mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x > gpurun_out/t_model.log 2>&1; echo "model rc=$?"
tail -n 5 gpurun_out/t_model.log
for nb in 1 2 4 8; do
ACSR_STEP_BRANCHES=$nb timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nb$nb.json 2> gpurun_out/bench_nb$nb.err; echo "nb=$nb rc=$?"
python scripts/show_bench.py < gpurun_out/bench_nb$nb.json | head -1
done
