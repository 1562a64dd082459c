mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 60 -k "attention" -x 2>&1 | tail -5
LD_ATTN_VARIANT=48 timeout 60 python -u tools/attn_phase_prof.py > gpurun_out/attn_phase4.log 2>&1; cat gpurun_out/attn_phase4.log
timeout 100 python tools/kernel_bench.py attn > gpurun_out/kbench6.log 2>&1; cat gpurun_out/kbench6.log
