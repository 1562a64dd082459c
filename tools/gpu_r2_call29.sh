mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" 2>&1 | tail -3
timeout 200 python tools/attn_sustained_bench.py --rounds 40 2>&1 | tee gpurun_out/r2c29_attn_sustained.txt
