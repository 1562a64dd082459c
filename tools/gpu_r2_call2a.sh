#!/bin/bash
# first contact of the new attention kernel with the GPU: mbarrier hang-watchdog build, correctness only
export LD_EXTRA_NVCC_FLAGS=-DLD_HANG_CHECK
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" > gpurun_out/r2c2a_pytest_attn.log 2>&1
echo "pytest attn exit $?" >> gpurun_out/r2c2a_pytest_attn.log
tail -25 gpurun_out/r2c2a_pytest_attn.log
