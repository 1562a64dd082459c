mkdir -p gpurun_out
export LD_WORKER_ONE_GPU=1 MASTER_ADDR=127.0.0.1 LD_ATTN_WAIT_MS=20000 CUDA_MODULE_LOADING=EAGER LD_WORKER_WATCHDOG_S=100
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29533 tests/ring_worker.py dma sp > gpurun_out/r2c31_worker.txt 2>&1
echo "rc=$?"; grep -v "^$" gpurun_out/r2c31_worker.txt | tail -60
