mkdir -p gpurun_out
timeout 40 python run_recbole.py --model=ACSASRec --dataset=ml-100k --config_files=config/ml-100k.yaml --epochs=1 --checkpoint_dir=/tmp/acsr_sas > gpurun_out/ml100k_acsasrec_1ep.log 2>&1; echo "ACSASRec rc=$?"
grep -E "training \[|Error|error|DeviceTrain" gpurun_out/ml100k_acsasrec_1ep.log | tail -3
tail -n 1 gpurun_out/ml100k_acsasrec_1ep.log | cut -c1-200
