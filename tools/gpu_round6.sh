mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 600 -k attention 2>&1 | tail -5
for st in 0 300 450 540 650 800; do
echo "== stagger $st"
LD_ATTN_STAGGER=$st timeout 300 python tools/attn_phase_prof.py 2>&1 | tail -8
LD_ATTN_STAGGER=$st timeout 600 python tools/kernel_bench.py attn 2>&1 | grep "variant 1[67]"
done 2>&1 | tee gpurun_out/stagger.log
