mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA --timeout 600 2>&1 | tail -40 > gpurun_out/gpu_tests2.log
grep -E "PASSED|FAILED|passed|failed|rel-L2|Error" gpurun_out/gpu_tests2.log | tail -30
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.log 2>&1; tail -5 gpurun_out/bench1.log
