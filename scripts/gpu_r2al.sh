# plain (transformer_layers.py) attention variant + ACSSEPT: kernel and model tests, then a headline regression check
mkdir -p gpurun_out
cd tests
timeout 900 python -m pytest test_gpu_kernels.py -x -q -m gpu -k "transformer_layers_variant or bidirectional or attn_calib_forward" 2>&1 | tail -15
timeout 900 python -m pytest test_gpu_model.py -x -q -m gpu -k "ssept" 2>&1 | tail -25
cd ..
timeout 600 python bench.py --steps 200 --warmup 20 --no-long-seq > gpurun_out/bench_al.json 2> gpurun_out/bench_al.err; echo "bench rc=$?"
python scripts/show_bench.py < gpurun_out/bench_al.json 2>/dev/null | head -8; tail -n 3 gpurun_out/bench_al.err
