"""Multi-rank parity under `pytest -m gpu`: N processes drive the sequence-parallel attention (copy-engine K|V exchange
through CUDA IPC buffers, in-kernel arrival-flag waits, ONE multi-shard attention launch) and the CFG x sequence-parallel
network step, and every rank compares with the single-rank path it computes itself (tests/ring_worker.py).

On a one-GPU box all ranks share the GPU (gloo control plane): the data path — IPC mapping, peer copies, stream memory
operations, in-kernel polls — is the one the 4- and 8-GPU runs use.  With >= 2 / >= 4 visible GPUs the same worker also
runs one rank per GPU over NCCL, for both transports."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_worker(world, transport, layouts, one_gpu, timeout=900):
    env = dict(os.environ, LD_WORKER_ONE_GPU="1" if one_gpu else "0", MASTER_ADDR="127.0.0.1")
    if one_gpu:
        # the ranks time-slice ONE device: a kernel polling for a peer's shard only progresses when the peer gets its
        # slice, and a lazy module load of a peer's first launch cannot overlap a spinning kernel — be patient and load
        # eagerly (one rank per GPU, the real deployment, needs neither)
        env.update(LD_ATTN_WAIT_MS="60000", CUDA_MODULE_LOADING="EAGER")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tests" / "ring_worker.py"), transport] + list(layouts)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=str(ROOT))
    lines = [l for l in r.stdout.splitlines() if l.startswith("ring_worker")]
    print("\n".join(lines))
    assert r.returncode == 0, f"worker failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    assert len(lines) == len(layouts) and all(l.endswith("OK") for l in lines), lines
    return lines


def test_two_ranks_on_one_gpu_cfg_and_sequence_parallel():
    """world 2: CFG-parallel row exchange (cfg2) and sequence parallel over 2 ranks (sp2, dma transport)."""
    run_worker(2, "dma", ["cfg", "sp"], one_gpu=True)


def test_four_ranks_on_one_gpu_cfg_x_sp2_and_sp4():
    """world 4: the 4-GPU layout (cfg2 x sp2) and a 4-rank sequence-parallel group (the 8-GPU layout's sp4 ring)."""
    run_worker(4, "dma", ["cfg", "sp"], one_gpu=True)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("transport", ["dma", "nccl"])
def test_two_gpus_nccl(transport):
    run_worker(2, transport, ["cfg", "sp"], one_gpu=False)


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs 4 GPUs")
@pytest.mark.parametrize("transport", ["dma", "nccl"])
def test_four_gpus_nccl(transport):
    run_worker(4, transport, ["cfg", "sp"], one_gpu=False)
