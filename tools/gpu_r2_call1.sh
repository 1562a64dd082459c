#!/bin/bash
# round 2, GPU call 1: state check of the round-1 product + new parity tests, microbenchmarks for the attn5 design, B2 baseline
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_kernels_gpu.py::test_attention -s > gpurun_out/r2c1_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c1_pytest.log
timeout 120 tools/softmax_mix_bench > gpurun_out/r2c1_softmax_mix.txt 2>&1
timeout 120 tools/mma_bench > gpurun_out/r2c1_mma_bench.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
tail -3 gpurun_out/r2c1_pytest.log
tail -12 gpurun_out/r2c1_softmax_mix.txt
tail -5 gpurun_out/r2c1_mma_bench.txt
cat gpurun_out/r2c1_bench.json
