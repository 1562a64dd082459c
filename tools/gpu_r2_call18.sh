mkdir -p gpurun_out
export LD_EXTRA_NVCC_FLAGS=-DLD_HANG_CHECK
timeout 240 python -m pytest tests/test_semantic.py -m gpu -x -q -s 2>&1 | tail -30 | tee gpurun_out/r2c18_semantic_tests.txt
