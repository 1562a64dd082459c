set -x
mkdir -p gpurun_out
# launch list of our kernels over whole sampler steps (one pass, cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_kernel|attn|ln_modulate|final_norm|patchify|small_linear|sampler_update|timestep" -c 1600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
# full captures: attention (B=1 to keep the replays short) and the four GEMM shapes
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn4_kernel -c 1 -f -o gpurun_out/prof_attn4_r1 python tools/kernel_bench.py attn --iters 1 --warmup 0 --batch 1 > gpurun_out/ncu_attn4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 5 -f -o gpurun_out/prof_gemm_r1b python tools/kernel_bench.py gemm --iters 1 --warmup 0 > gpurun_out/ncu_gemm2.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"ln_modulate|sampler_update" -c 2 -f -o gpurun_out/prof_rows_r1 python tools/kernel_bench.py rows --iters 1 --warmup 0 > gpurun_out/ncu_rows.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r1.csv
