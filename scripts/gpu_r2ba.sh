# sibling models end to end on ml-100k through run_recbole.py (1 epoch + validation + test)
mkdir -p gpurun_out
for M in ACTiSASRec:actisasrec ACSSEPT:acssept; do
  N=${M%%:*}; Y=${M##*:}
  timeout 55 python run_recbole.py --model=$N --dataset=ml-100k --config_files=config/ml-100k-$Y.yaml --epochs=1 --checkpoint_dir=/tmp/acsr_$Y > gpurun_out/ml100k_$Y.log 2>&1; echo "$N rc=$?"
  grep -E "training \[|valid result|test result|Error|error" gpurun_out/ml100k_$Y.log | tail -4
  tail -n 2 gpurun_out/ml100k_$Y.log | cut -c1-300
done
