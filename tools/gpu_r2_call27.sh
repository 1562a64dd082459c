mkdir -p gpurun_out
for kp in 2 3 5; do LD_ATTN_KP=$kp timeout 200 python tools/attn_sustained_bench.py --rounds 40 2>&1 | grep -v "^each\|^inter" ; done | tee gpurun_out/r2c27_attn_sustained_kp.txt
