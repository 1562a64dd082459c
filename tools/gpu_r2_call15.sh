mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "row or layernorm or ln" 2>&1 | tail -2
for c in 15 115 0; do echo "LD_LN_CFG=$c"; LD_LN_CFG=$c timeout 100 python tools/kernel_bench.py rows --iters 20 2>&1 | grep layernorm; done | tee gpurun_out/r2c15_ln_cfgs.txt
