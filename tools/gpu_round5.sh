mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 600 2>&1 | tail -15
timeout 300 python tools/attn_phase_prof.py > gpurun_out/attn_phase.log 2>&1; cat gpurun_out/attn_phase.log
timeout 600 python tools/kernel_bench.py gemm attn > gpurun_out/kbench4.log 2>&1; cat gpurun_out/kbench4.log
