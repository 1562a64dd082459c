mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_network_gpu.py::test_full_shape_step_against_oracle_fp32_on_gpu -m gpu -q --timeout 280 -s 2>&1 | tail -6
timeout 150 python tools/attn_stress.py 2>&1 | tail -3
