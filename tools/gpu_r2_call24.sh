mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r2c24_smoke.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ln_modulate_bulk|final_norm|sampler_update" -c 3 -o gpurun_out/r2c24_rows python tools/kernel_bench.py rows --iters 1 --warmup 1 > gpurun_out/r2c24_ncu_rows.log 2>&1
python tools/ncu_summary.py gpurun_out/r2c24_rows.ncu-rep > gpurun_out/r2c24_ncu_rows.csv 2>&1; cat gpurun_out/r2c24_ncu_rows.csv | cut -c1-400
timeout 400 ncu --set full --clock-control none -k regex:"gemm_kernel" -c 6 -o gpurun_out/r2c24_conv python tools/kernel_bench.py semantic --iters 1 --warmup 1 > gpurun_out/r2c24_ncu_conv.log 2>&1
python tools/ncu_summary.py gpurun_out/r2c24_conv.ncu-rep > gpurun_out/r2c24_ncu_conv.csv 2>&1; cat gpurun_out/r2c24_ncu_conv.csv | cut -c1-300
rm -f gpurun_out/r2c24_conv.ncu-rep
