#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_network_gpu.py tests/test_kernels_gpu.py -m gpu -q -k "graph or row_kernels or tiny" -s > gpurun_out/r2c7_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c7_pytest.log
grep -E "^FAILED|passed|failed|Error" gpurun_out/r2c7_pytest.log | tail -8
timeout 100 python tools/kernel_bench.py rows --iters 5 > gpurun_out/r2c7_kernel_bench_rows.txt 2>&1
cat gpurun_out/r2c7_kernel_bench_rows.txt
timeout 600 python -m pytest tests/test_multirank_gpu.py -m gpu -q -x -s > gpurun_out/r2c7_pytest_multirank.log 2>&1
echo "pytest multirank exit $?" >> gpurun_out/r2c7_pytest_multirank.log
grep -E "ring_worker world|passed|failed" gpurun_out/r2c7_pytest_multirank.log | tail -8
timeout 400 python bench.py --steps 10 --warmup 3 --no-eager > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c7_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','cuda_graph','shard_wait_timeouts')}, d['roofline']['launch_ms'], d['roofline']['frac'])
PY
tail -3 gpurun_out/r2c7_bench.err
