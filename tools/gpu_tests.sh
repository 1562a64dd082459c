mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA --timeout 600 2>&1 | tail -60 > gpurun_out/gpu_tests.log
tail -45 gpurun_out/gpu_tests.log
