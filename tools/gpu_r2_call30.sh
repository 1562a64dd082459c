mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_multirank_gpu.py tests/test_torch_ops.py -m gpu -x -q -k "unpatchify or ranks or ops" -s 2>&1 | grep -v "^$" | tail -25 | tee gpurun_out/r2c30_tests.txt
