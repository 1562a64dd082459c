set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 5 -f -o gpurun_out/prof_gemm_r1 python tools/kernel_bench.py gemm --iters 1 --warmup 0 > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -c 1 -f -o gpurun_out/prof_attn_r1 python tools/kernel_bench.py attn --iters 1 --warmup 0 --batch 1 > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -5 gpurun_out/ncu_gemm.log gpurun_out/ncu_attn.log
