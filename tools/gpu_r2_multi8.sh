#!/bin/bash
# 8-GPU box: bench at N = 8 (denoise + stream3), rank-0 step breakdown
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/r2m8_bench_n8.err | grep "^{" > gpurun_out/r2m8_bench_n8.json
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2m8_bench_n8.json').read().strip().splitlines()[-1])
    print(8, {k:d.get(k) for k in ('value','ms_per_step','gpu_launches','shard_wait_timeouts')}, d['e2e']['value'], d['roofline']['launch_ms'], d['roofline']['frac'], d['clocks'])
except Exception as e:
    print('bench 8 failed', e)
PY
tail -2 gpurun_out/r2m8_bench_n8.err | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/multigpu_profile.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM\|warn" | tail -18 | tee gpurun_out/r2m8_step_breakdown_8gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --workload stream3 2> gpurun_out/r2m8_stream3_n8.err | grep "^{" > gpurun_out/r2m8_stream3_n8.json
cut -c1-900 gpurun_out/r2m8_stream3_n8.json; tail -2 gpurun_out/r2m8_stream3_n8.err | cut -c1-300
