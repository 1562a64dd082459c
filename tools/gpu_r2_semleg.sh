mkdir -p gpurun_out
timeout 300 python bench.py --steps 4 --warmup 3 --no-graph > gpurun_out/r2_bench_semantic_leg.json 2> gpurun_out/r2_bench_semantic_leg.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_semantic_leg.json").read().strip().splitlines()[-1])
print(d["value"], d["semantic_conditioner"])
PY
tail -2 gpurun_out/r2_bench_semantic_leg.err
