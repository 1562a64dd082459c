#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" > gpurun_out/r2c4_pytest_attn.log 2>&1
echo "pytest attn exit $?" >> gpurun_out/r2c4_pytest_attn.log
grep -E "^FAILED|passed|failed" gpurun_out/r2c4_pytest_attn.log | tail -8
timeout 120 python tools/kernel_bench.py attn --iters 5 > gpurun_out/r2c4_kernel_bench_attn.txt 2>&1
cat gpurun_out/r2c4_kernel_bench_attn.txt
timeout 60 python tools/attn_phase_prof.py > gpurun_out/r2c4_attn_phase.txt 2>&1
cat gpurun_out/r2c4_attn_phase.txt
