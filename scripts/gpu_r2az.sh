mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1; echo "multi_gpu_check rc=$?"
grep -v "^\[W\|Warning" gpurun_out/multi_check.log | tail -n 12
