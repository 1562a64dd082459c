mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/probe1_smi.txt 2>&1
for c in rows gemm attn0 attn1; do
  echo "=== $c" >> gpurun_out/probe1.log
  timeout 240 python tools/gpu_probe.py $c >> gpurun_out/probe1.log 2>&1
  echo "exit $?" >> gpurun_out/probe1.log
done
tail -c 6000 gpurun_out/probe1.log
