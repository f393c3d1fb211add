mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_all_u.log 2>&1; echo "all rc=$?"
tail -n 4 gpurun_out/t_all_u.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
