mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_network_gpu.py::test_50_step_trajectory_psnr tests/test_streaming.py -m gpu -q --timeout 300 -s 2>&1 | tail -12
timeout 120 python tools/host_overhead.py 2>&1 | tail -3
