mkdir -p gpurun_out
timeout 45 python run_recbole.py --model=AcBERT4Rec --dataset=ml-100k --config_files=config/ml-100k-acbert4rec.yaml --epochs=1 --checkpoint_dir=/tmp/acsr_bert > gpurun_out/ml100k_acbert4rec.log 2>&1; echo "AcBERT4Rec rc=$?"
grep -E "training \[|Error|error" gpurun_out/ml100k_acbert4rec.log | tail -3
tail -n 1 gpurun_out/ml100k_acbert4rec.log | cut -c1-400
