"""Worker of tests/test_multirank_gpu.py (also usable by hand under torch.distributed.run): multi-rank parity of the
sequence-parallel attention and of the CFG x sequence-parallel network step against the single-rank path.

  LD_WORKER_ONE_GPU=1   every rank uses cuda:0 and the control plane is gloo — N processes share ONE GPU; CUDA IPC,
                        the copy-engine peer copies, the stream memory operations and the in-kernel arrival-flag waits
                        are exactly what the multi-GPU run uses, only the link underneath differs (and NCCL, which
                        refuses two ranks on one device, is replaced by gloo for the tiny control-plane collectives)
  otherwise             one GPU per rank, NCCL

usage: ring_worker.py <transport: dma|nccl> <layout> [<layout> ...]   with layout = cfg | sp   (cfg: CFG-parallel x
       sequence-parallel world/2; sp: sequence-parallel over the whole world)
Checks per layout (every rank computes the single-rank reference itself):
  1. RingAttention == monolithic attention over the full K/V, 6 calls with fresh data and deliberately skewed ranks
  2. CFGGroup.evaluate == the single-rank network output (both CFG rows), and a 3-step sampler trajectory
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from landiff_b200 import dit, ops, parallel  # noqa: E402
from landiff_b200.factory import DiTShape, build_warp, random_init_  # noqa: E402
from landiff_b200.sampling import VPSDEDPMPP2MSampler  # noqa: E402

TOL = 5e-3


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def check_ring(layout, sp_group, dev, transport):
    """ring attention over the sequence-parallel group == monolithic attention (rows of this rank)."""
    sp, me = layout.sp_size, layout.sp_rank
    ring = parallel.RingAttention(layout, sp_group, dev, transport=transport)
    B, H, R = 1, 3, 348          # 348 = 5 full sub-blocks + 28 keys: ragged shard tails like 4444
    N = R * sp
    ws = dict(q=torch.empty(B, H, R, 64, device=dev, dtype=torch.bfloat16),
              kv=torch.empty(2, B, H, R, 64, device=dev, dtype=torch.bfloat16),
              attn=torch.empty(B, R, H * 64, device=dev, dtype=torch.bfloat16))
    worst = 0.0
    for call in range(6):
        g = torch.Generator(device=dev).manual_seed(100 + call)      # identical full tensors on every rank
        qf = torch.randn(B, H, N, 64, device=dev, generator=g).bfloat16()
        kf = torch.randn(B, H, N, 64, device=dev, generator=g).bfloat16()
        vf = torch.randn(B, H, N, 64, device=dev, generator=g).bfloat16()
        sl = slice(me * R, (me + 1) * R)
        ws["q"].copy_(qf[:, :, sl])
        ws["kv"][0].copy_(kf[:, :, sl])
        ws["kv"][1].copy_(vf[:, :, sl])
        ring.attention(ws)
        if call % 3 == 2:
            torch.cuda._sleep(int(1e7) * (layout.rank + 1))    # skew the ranks: the flags, not luck, must order things
        ref = torch.nn.functional.scaled_dot_product_attention(qf[:, :, sl].float(), kf.float(), vf.float())
        ref = ref.permute(0, 2, 1, 3).reshape(B, R, H * 64)
        worst = max(worst, rel(ws["attn"].float(), ref))
    torch.cuda.synchronize()
    status = ops.attention_status()
    for pg in ring._peer.values():
        pg.close()
    return worst, status


def check_network(layout, sp_group, dev, transport):
    cfg = DiTShape(hidden_size=384, num_heads=6, main_layers=3, control_layers=2, time_embed_dim=128, text_hidden=256,
                   text_length=6, latent_t=2, latent_h=30, latent_w=46)
    assert cfg.n_tok % max(layout.sp_size, 1) == 0

    def make():
        w = build_warp(cfg, device=dev)
        random_init_(w, seed=0, std=0.05)
        return w

    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, cfg.latent_t, 16, cfg.latent_h, cfg.latent_w, generator=g).to(dev)
    ctx = (torch.randn(1, cfg.text_length, cfg.text_hidden, generator=g) * 0.2).bfloat16().to(dev)
    sem = (torch.randn(1, cfg.latent_t, 16, cfg.latent_h, cfg.latent_w, generator=g) * 0.1).bfloat16().to(dev)
    dit.InferValueRegistry.clear()
    dit.InferValueRegistry.register("semantic_feature", sem)
    cond, uc = {"crossattn": ctx}, {"crossattn": torch.zeros_like(ctx)}
    ctx2 = torch.cat([uc["crossattn"], cond["crossattn"]])

    warp1 = make()           # single-rank reference on this rank
    t2 = torch.full((2,), 519.0, device=dev)
    net = warp1(torch.cat([x, x]), t2, {"crossattn": ctx2}, idx=t2).float().clone()
    sampler = VPSDEDPMPP2MSampler(num_steps=50, device="cuda")
    cpu_gen = torch.Generator().manual_seed(42)                      # identical noise on every rank and in both runs
    noise = lambda t: torch.randn(t.shape, generator=cpu_gen).to(t.device)
    traj1 = sampler.sample(warp1, x.clone(), cond, uc, start_step=0, max_steps=3, noise_fn=noise).float().clone()
    torch.cuda.synchronize()
    del warp1

    warp = make()
    if layout.sp_size > 1:
        os.environ["LD_RING_TRANSPORT"] = transport
    parallel.attach(warp, layout, sp_group, dev)
    grp = parallel.CFGGroup(layout)
    net_u, net_c = grp.evaluate(warp, x, 519.0, ctx2)
    r_u, r_c = rel(net_u[0].float(), net[0]), rel(net_c[0].float(), net[1])
    cpu_gen.manual_seed(42)
    traj = sampler.sample(warp, x.clone(), cond, uc, cfg_group=grp, start_step=0, max_steps=3, noise_fn=noise).float()
    r_t = rel(traj, traj1)
    torch.cuda.synchronize()
    status = ops.attention_status()
    for m in (warp.control_model.diffusion_model, warp.main_model.diffusion_model):
        if m.ring is not None:
            for pg in m.ring._peer.values():
                pg.close()
    if layout.sp_size > 1 and transport == "dma":
        assert grp._gather is not None, "the copy-engine output exchange was not used"
    if grp._gather is not None:
        grp._gather.close()
    return max(r_u, r_c), r_t, status


def main():
    import faulthandler

    # a rank stuck in a collective or a stream wait dumps its Python stack and exits instead of hanging the test
    faulthandler.dump_traceback_later(int(os.environ.get("LD_WORKER_WATCHDOG_S", "600")), exit=True)
    transport, layouts = sys.argv[1], sys.argv[2:]
    world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
    one_gpu = os.environ.get("LD_WORKER_ONE_GPU", "0") == "1"
    lr = 0 if one_gpu else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if one_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    ok = True
    for name in layouts:
        layout = parallel.make_layout(world, rank, cfg_parallel=(name == "cfg"))
        sp_group, _ = parallel.new_subgroups(layout)
        errs = [0.0, 0.0, 0.0, 0.0]
        if layout.sp_size > 1:
            errs[0], st = check_ring(layout, sp_group, dev, transport)
            errs[3] = float(st)
        errs[1], errs[2], st = check_network(layout, sp_group, dev, transport)
        errs[3] = max(errs[3], float(st))
        t = torch.tensor(errs, dtype=torch.float64, device="cpu" if one_gpu else dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        good = bool((t[:3] < TOL).all()) and float(t[3]) == 0.0
        ok = ok and good
        if rank == 0:
            print(f"ring_worker world={world} layout=cfg{layout.cfg_size}xsp{layout.sp_size} transport={transport} "
                  f"{'one GPU' if one_gpu else 'one GPU per rank'}: ring-vs-monolithic {t[0]:.3e} network {t[1]:.3e} "
                  f"3-step trajectory {t[2]:.3e} wait-timeouts {int(t[3])} {'OK' if good else 'FAIL'}", flush=True)
        dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
