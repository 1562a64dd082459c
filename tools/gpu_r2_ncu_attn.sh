mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"attn5_kernel|attn_split_merge" -c 2 -o gpurun_out/r2_attn5_final python tools/kernel_bench.py attn --iters 1 --warmup 1 --batch 1 > gpurun_out/r2_ncu_attn_final.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_attn5_final.ncu-rep > gpurun_out/r2_ncu_attn5_final.csv 2>&1; cut -c1-420 gpurun_out/r2_ncu_attn5_final.csv
ls -la gpurun_out/r2_attn5_final.ncu-rep
