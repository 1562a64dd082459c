mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm" 2>&1 | tail -2
timeout 100 python tools/kernel_bench.py gemm --iters 7 > gpurun_out/r2c9_kernel_bench_gemm.txt 2>&1
cat gpurun_out/r2c9_kernel_bench_gemm.txt
timeout 100 python tools/kernel_bench.py gemm --iters 7 --ntok 4444 --batch 1 > gpurun_out/r2c9_kernel_bench_gemm_sp4.txt 2>&1
cat gpurun_out/r2c9_kernel_bench_gemm_sp4.txt
