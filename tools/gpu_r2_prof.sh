timeout 120 python tools/attn_phase_prof.py > gpurun_out/r2_attn_phase.txt 2>&1
cat gpurun_out/r2_attn_phase.txt
