mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 400 2>&1 | tail -6 | tee gpurun_out/r2_final_pytest.txt
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/r2_final_smoke.txt
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_final_bench_1gpu.json 2> gpurun_out/r2_final_bench_1gpu.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_final_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['cuda_graph']['ms_per_step'], d['gpu_eager_baseline']['value'], d['clocks'])
PY
