mkdir -p gpurun_out
export LD_WORKER_WATCHDOG_S=150
timeout 500 python -m pytest tests/test_multirank_gpu.py tests/test_torch_ops.py -m gpu -x -q -s --timeout 240 2>&1 | grep -v "^$" | tail -12 | tee gpurun_out/r2c32_tests.txt
