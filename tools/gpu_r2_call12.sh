mkdir -p gpurun_out
timeout 100 python tools/kernel_bench.py rows --iters 9 > gpurun_out/r2c12_kernel_bench_rows.txt 2>&1
cat gpurun_out/r2c12_kernel_bench_rows.txt
