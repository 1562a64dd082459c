mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 600 -k attention 2>&1 | tail -15
timeout 300 python tools/attn_phase_prof.py > gpurun_out/attn_phase.log 2>&1; cat gpurun_out/attn_phase.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn2_kernel -c 1 -f -o gpurun_out/prof_attn2_r1 python tools/kernel_bench.py attn --iters 1 --warmup 0 --batch 1 > gpurun_out/ncu_attn2.log 2>&1
tail -3 gpurun_out/ncu_attn2.log
