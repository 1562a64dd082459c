#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q -s > gpurun_out/r2c10_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c10_pytest.log
grep -E "^FAILED|passed|failed|rel-L2|PSNR|ring_worker" gpurun_out/r2c10_pytest.log | tail -16
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c10_bench.json 2> gpurun_out/r2c10_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c10_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','cuda_graph','shard_wait_timeouts','gpu_eager_baseline','clocks')}, d['roofline'])
PY
tail -3 gpurun_out/r2c10_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm|attn|ln_modulate|final_norm|patchify|small_linear|sampler_update|timestep" -s 1100 -c 400 --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --steps 1 --warmup 3 --no-eager --no-graph > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -c 5 -f -o gpurun_out/prof_gemm2_r2 python tools/kernel_bench.py gemm --iters 1 --warmup 0 > gpurun_out/ncu_gemm2.log 2>&1
ls -la gpurun_out/*.csv gpurun_out/prof_gemm2_r2.ncu-rep
