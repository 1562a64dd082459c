#!/bin/bash
# round 2, GPU call 6: full single-GPU test suite with the new kernels, kernel bench, bench.py, ncu captures
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_multirank_gpu.py -s > gpurun_out/r2c6_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c6_pytest.log
grep -E "^FAILED|passed|failed|rel-L2|PSNR" gpurun_out/r2c6_pytest.log | tail -14
timeout 200 python tools/kernel_bench.py attn rows gemm --iters 5 > gpurun_out/r2c6_kernel_bench.txt 2>&1
cat gpurun_out/r2c6_kernel_bench.txt
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err
cut -c1-1500 gpurun_out/r2c6_bench.json; tail -3 gpurun_out/r2c6_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn5_kernel -c 1 -f -o gpurun_out/prof_attn5_r2 python tools/kernel_bench.py attn --iters 1 --warmup 0 --batch 1 > gpurun_out/ncu_attn5.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:"ln_modulate|sampler_update" -c 2 -f -o gpurun_out/prof_rows_r2 python tools/kernel_bench.py rows --iters 1 --warmup 0 > gpurun_out/ncu_rows.log 2>&1
ls -la gpurun_out/*.ncu-rep
