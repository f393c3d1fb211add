mkdir -p gpurun_out
timeout 600 python run_recbole.py --model=ACSASRec --dataset=ml-100k --config_files=config/ml-100k.yaml --epochs=3 --checkpoint_dir=/tmp/acsr_ml100k > gpurun_out/ml100k.log 2>&1; echo "run rc=$?"
grep -E "training \[|evaluating|valid result|recall@10|best valid|test result" gpurun_out/ml100k.log | tail -30
tail -n 5 gpurun_out/ml100k.log
