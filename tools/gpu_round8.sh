mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench3.log 2>&1; tail -2 gpurun_out/bench3.log
