mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 600 2>&1 | tail -8
timeout 600 python tools/kernel_bench.py gemm attn > gpurun_out/kbench2.log 2>&1; cat gpurun_out/kbench2.log
