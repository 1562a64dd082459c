# usage: gpu_multi.sh N [check]   (N ranks on one box)
N=$1
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
if [ "$2" = "check" ]; then
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -3 | tee gpurun_out/multigpu_check_$N.log
fi
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -3 | tee gpurun_out/bench_n$N.log
