mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_network_gpu.py -m gpu -q -k "row_kernels or tiny or small_warp or config1_against" 2>&1 | tail -3
timeout 100 python tools/kernel_bench.py rows --iters 7 > gpurun_out/r2c11_kernel_bench_rows.txt 2>&1
cat gpurun_out/r2c11_kernel_bench_rows.txt
