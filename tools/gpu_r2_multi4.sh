#!/bin/bash
# 4-GPU box: multi-rank parity over NCCL (2 and 4 GPUs, both transports), benches at N = 2 and 4, rank-0 breakdown at N = 4
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -4
timeout 500 python -m pytest tests/test_multirank_gpu.py -m gpu -q -s -k "gpus" > gpurun_out/r2m4_pytest_multirank.log 2>&1
echo "pytest multirank exit $?" >> gpurun_out/r2m4_pytest_multirank.log
grep -E "ring_worker world|passed|failed|skipped" gpurun_out/r2m4_pytest_multirank.log | tail -12
for N in 2 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/r2m4_bench_n$N.err | grep "^{" > gpurun_out/r2m4_bench_n$N.json
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2m4_bench_n$N.json').read().strip().splitlines()[-1])
    print($N, {k:d.get(k) for k in ('value','ms_per_step','gpu_launches','shard_wait_timeouts')}, d['e2e']['value'], d['roofline']['launch_ms'], d['roofline']['frac'], d['clocks'])
except Exception as e:
    print('bench $N failed', e)
PY
tail -2 gpurun_out/r2m4_bench_n$N.err | cut -c1-300
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 tools/multigpu_profile.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -20 | tee gpurun_out/r2m4_step_breakdown_4gpu.txt
