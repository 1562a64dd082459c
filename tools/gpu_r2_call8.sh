mkdir -p gpurun_out
timeout 100 python tools/kernel_bench.py gemm --iters 7 > gpurun_out/r2c8_kernel_bench_gemm.txt 2>&1
cat gpurun_out/r2c8_kernel_bench_gemm.txt
LD_GEMM_1CTA=1 timeout 100 python tools/kernel_bench.py gemm --iters 7 > gpurun_out/r2c8_kernel_bench_gemm_1cta.txt 2>&1
cat gpurun_out/r2c8_kernel_bench_gemm_1cta.txt
timeout 100 python tools/kernel_bench.py gemm --iters 7 --ntok 4444 --batch 1 > gpurun_out/r2c8_kernel_bench_gemm_sp4.txt 2>&1
cat gpurun_out/r2c8_kernel_bench_gemm_sp4.txt
LD_GEMM_1CTA=1 timeout 100 python tools/kernel_bench.py gemm --iters 7 --ntok 4444 --batch 1 > gpurun_out/r2c8_kernel_bench_gemm_sp4_1cta.txt 2>&1
cat gpurun_out/r2c8_kernel_bench_gemm_sp4_1cta.txt
timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_network_gpu.py -m gpu -q -k "gemm or tiny or config1 or small_warp" > gpurun_out/r2c8_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r2c8_pytest.log | tail -5
