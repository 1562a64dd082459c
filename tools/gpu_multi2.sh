N=$1
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
export NCCL_P2P_USE_CUDA_MEMCPY=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/multigpu_profile.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -20 | tee gpurun_out/multigpu_profile_${N}_ce.log
