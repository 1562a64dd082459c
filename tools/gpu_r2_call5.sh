#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_torch_ops.py -m gpu -q > gpurun_out/r2c5_pytest_kernels.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c5_pytest_kernels.log
grep -E "^FAILED|passed|failed|Error" gpurun_out/r2c5_pytest_kernels.log | tail -8
timeout 150 python tools/kernel_bench.py attn rows --iters 5 > gpurun_out/r2c5_kernel_bench.txt 2>&1
cat gpurun_out/r2c5_kernel_bench.txt
timeout 60 python tools/attn_phase_prof.py > gpurun_out/r2c5_attn_phase.txt 2>&1
cat gpurun_out/r2c5_attn_phase.txt
