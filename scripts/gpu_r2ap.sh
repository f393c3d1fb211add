mkdir -p gpurun_out
python scripts/sibling_prof.py ACTiSASRec 2>&1 | tail -30 | tee gpurun_out/sibling_prof_ti.txt
python scripts/sibling_prof.py ACSSEPT 2>&1 | tail -30 | tee gpurun_out/sibling_prof_ssept.txt
