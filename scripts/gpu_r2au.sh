# ncu launch lists of one eager training step of the sibling models (after the program has exited 0 without ncu)
mkdir -p gpurun_out
for M in ACTiSASRec ACSSEPT; do
  python scripts/sibling_prof.py $M > gpurun_out/sib_plain_$M.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sib_plain_$M.log; exit 1; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_$M.csv python scripts/sibling_prof.py $M > gpurun_out/sib_ncu_$M.log 2>&1
  echo "$M list rc=$?"
  python scripts/summarize_launches.py gpurun_out/r02_launches_$M.csv 16 | tee gpurun_out/r02_launches_${M}_summary.txt
done
