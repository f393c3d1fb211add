mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --tb=short -x > gpurun_out/t_model.log 2>&1; echo "model tests rc=$?"; tail -n 4 gpurun_out/t_model.log | grep -v Warn
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -n 3 gpurun_out/bench.err; python scripts/show_bench.py < gpurun_out/bench.json 2>/dev/null | head -1
timeout 600 python run_recbole.py --model=ACSASRec --dataset=ml-100k --config_files=config/ml-100k.yaml --epochs=1 --checkpoint_dir=/tmp/acsr_ml100k > gpurun_out/ml100k.log 2>&1; echo "run rc=$?"
grep -E "epoch [0-9]+ (training|evaluating)" gpurun_out/ml100k.log | tail -6
