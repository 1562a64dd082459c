mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_semantic.py tests/test_torch_ops.py -m gpu -q -s 2>&1 | tail -25 | tee gpurun_out/r2c16_semantic_tests.txt
