mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -5
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log
