mkdir -p gpurun_out
timeout 300 python tools/kernel_bench.py semantic --iters 5 2>&1 | tee gpurun_out/r2c17_kernel_bench_semantic.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c17_ncu_semantic_launches.csv python tools/kernel_bench.py semantic --iters 1 --warmup 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2c17_ncu_semantic_launches.csv')))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hd = rows[h]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[h + 2:]:
    d = dict(zip(hd, r))
    if d.get('Metric Name') == 'gpu__time_duration.sum':
        k = d['Kernel Name'][:70]
        agg[k][0] += 1; agg[k][1] += float(d['Metric Value'].replace(',', '')) / 1e3
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{t:10.1f} us {n:4d}  {k}")
PY
