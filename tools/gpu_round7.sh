mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 60 -k "attention and (32 or 33 or 34 or 35)" -x 2>&1 | tail -15
timeout 60 python -u tools/attn_phase_prof.py > gpurun_out/attn_phase3.log 2>&1; cat gpurun_out/attn_phase3.log
timeout 100 python tools/kernel_bench.py attn > gpurun_out/kbench5.log 2>&1; cat gpurun_out/kbench5.log
