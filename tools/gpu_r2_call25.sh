mkdir -p gpurun_out
timeout 300 python tools/kernel_bench.py gemm rows attn --iters 9 2>&1 | tee gpurun_out/r2c25_kernel_bench.txt
timeout 300 python tools/kernel_bench.py gemm --iters 9 --batch 1 --ntok 4444 2>&1 | tee gpurun_out/r2c25_kernel_bench_sp4.txt
