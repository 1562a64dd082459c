#!/bin/bash
# round 2, GPU call 2: attn5 timing + remaining attention tests, MMA microbench, multi-rank tests on one GPU
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" > gpurun_out/r2c2_pytest_attn.log 2>&1
echo "pytest attn exit $?" >> gpurun_out/r2c2_pytest_attn.log
grep -E "^FAILED|passed|failed" gpurun_out/r2c2_pytest_attn.log | tail -5
timeout 300 python tools/kernel_bench.py attn --iters 5 > gpurun_out/r2c2_kernel_bench_attn.txt 2>&1
cat gpurun_out/r2c2_kernel_bench_attn.txt
timeout 120 tools/mma_bench > gpurun_out/r2c2_mma_bench.txt 2>&1
tail -4 gpurun_out/r2c2_mma_bench.txt | cut -c1-160
timeout 120 python tools/attn_phase_prof.py > gpurun_out/r2c2_attn_phase.txt 2>&1
cat gpurun_out/r2c2_attn_phase.txt
timeout 900 python -m pytest tests/test_multirank_gpu.py -m gpu -q -x -s > gpurun_out/r2c2_pytest_multirank.log 2>&1
echo "pytest multirank exit $?" >> gpurun_out/r2c2_pytest_multirank.log
grep -E "ring_worker|passed|failed|Error|error" gpurun_out/r2c2_pytest_multirank.log | tail -12
