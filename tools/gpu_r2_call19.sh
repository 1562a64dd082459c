mkdir -p gpurun_out
export LD_EXTRA_NVCC_FLAGS=-DLD_HANG_CHECK
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" 2>&1 | tail -15 | tee gpurun_out/r2c19_attn_tests_hangcheck.txt
