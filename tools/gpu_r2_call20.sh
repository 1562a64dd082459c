mkdir -p gpurun_out
timeout 300 python tools/attn_split_bench.py --iters 7 2>&1 | tee gpurun_out/r2c20_attn_split_bench.txt
