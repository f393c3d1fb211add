mkdir -p gpurun_out
timeout 300 python scripts/step_timeline.py > gpurun_out/timeline_j.log 2>&1; echo "timeline rc=$?"
tail -n 8 gpurun_out/timeline_j.log
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/t_all_j.log 2>&1; echo "all rc=$?"
tail -n 4 gpurun_out/t_all_j.log
