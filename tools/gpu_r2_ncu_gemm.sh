mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -c 3 -f -o gpurun_out/prof_gemm2_r2 python tools/kernel_bench.py gemm --iters 1 --warmup 0 > gpurun_out/ncu_gemm2.log 2>&1
tail -3 gpurun_out/ncu_gemm2.log
ls -la gpurun_out/prof_gemm2_r2.ncu-rep
