mkdir -p gpurun_out
for kp in 5 4 5 4; do
  LD_ATTN_KP=$kp timeout 400 python bench.py --steps 10 --warmup 3 --no-eager --no-graph > gpurun_out/r2c28_bench_kp$kp.json 2> gpurun_out/r2c28_bench_kp$kp.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/r2c28_bench_kp$kp.json').read().strip().splitlines()[-1])
print('KP=$kp', d['value'], d['ms_per_step'], 'attn', d['roofline']['launch_ms'], d['roofline']['frac'], d['clocks'])
PY
done 2>&1 | tee gpurun_out/r2c28_kp_ab.txt
