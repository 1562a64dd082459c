mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2c14_pytest.txt
cat gpurun_out/r2c14_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c14_bench_1gpu.json 2> gpurun_out/r2c14_bench_1gpu.err
tail -c 3000 gpurun_out/r2c14_bench_1gpu.json
timeout 300 ncu --set full --clock-control none -k regex:ln_modulate_bulk -c 2 --csv --page raw --log-file gpurun_out/r2c14_ncu_ln.csv python tools/kernel_bench.py rows --iters 1 > gpurun_out/r2c14_ncu_ln.log 2>&1
tail -3 gpurun_out/r2c14_ncu_ln.log
