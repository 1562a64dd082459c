mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2c21_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c21_bench_1gpu.json 2> gpurun_out/r2c21_bench_1gpu.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2c21_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'], d['cuda_graph']['ms_per_step'], d['clocks'])
PY
