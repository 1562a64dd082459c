"""Multi-GPU parity on real NCCL (run under torch.distributed.run with 2, 4 or 8 ranks):
every rank evaluates the full CFG batch on its own GPU (single-GPU path), then the same step through the
CFG-parallel (x ring sequence-parallel) path; the assembled (uncond, cond) outputs and a 3-step sampler trajectory
must agree.  Shape: T=2, 30x46 latent -> N = 226 + 690 = 916 tokens (divisible by 4), full width, reduced depth.
"""
import dataclasses
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from landiff_b200 import dit, ops, parallel  # noqa: E402
from landiff_b200.factory import DiTShape, build_warp, random_init_  # noqa: E402
from landiff_b200.sampling import VPSDEDPMPP2MSampler  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def main():
    world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    layout = parallel.make_layout(world, rank)
    sp_group, _ = parallel.new_subgroups(layout)
    cfg = DiTShape(latent_t=2, latent_h=30, latent_w=46, main_layers=4, control_layers=2)
    assert cfg.n_tok % max(layout.sp_size, 1) == 0

    def make():
        w = build_warp(cfg, device=dev)
        random_init_(w, seed=0)
        return w

    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, cfg.latent_t, 16, cfg.latent_h, cfg.latent_w, generator=g).to(dev)
    ctx = (torch.randn(1, cfg.text_length, cfg.text_hidden, generator=g) * 0.2).bfloat16().to(dev)
    sem = (torch.randn(1, cfg.latent_t, 16, cfg.latent_h, cfg.latent_w, generator=g) * 0.1).bfloat16().to(dev)
    dit.InferValueRegistry.clear()
    dit.InferValueRegistry.register("semantic_feature", sem)
    cond, uc = {"crossattn": ctx}, {"crossattn": torch.zeros_like(ctx)}
    ctx2 = torch.cat([uc["crossattn"], cond["crossattn"]])

    # ---- single-GPU reference on this rank
    warp1 = make()
    t2 = torch.full((2,), 519.0, device=dev)
    net = warp1(torch.cat([x, x]), t2, {"crossattn": ctx2}, idx=t2).float().clone()
    sampler = VPSDEDPMPP2MSampler(num_steps=50, device="cuda")
    torch.manual_seed(42)
    traj1 = sampler.sample(warp1, x.clone(), cond, uc, start_step=0, max_steps=3).float().clone()
    torch.cuda.synchronize()
    del warp1

    # ---- parallel path
    warp = make()
    parallel.attach(warp, layout, sp_group, dev)
    grp = parallel.CFGGroup(layout)
    net_u, net_c = grp.evaluate(warp, x, 519.0, ctx2)
    r_u, r_c = rel(net_u[0].float(), net[0]), rel(net_c[0].float(), net[1])
    torch.manual_seed(42)
    traj = sampler.sample(warp, x.clone(), cond, uc, cfg_group=grp, start_step=0, max_steps=3).float()
    r_t = rel(traj, traj1)
    torch.cuda.synchronize()
    errs = torch.tensor([r_u, r_c, r_t], device=dev, dtype=torch.float64)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    if rank == 0:
        ok = bool((errs < 5e-3).all())
        print(f"multigpu_check world={world} layout=cfg{layout.cfg_size}xsp{layout.sp_size} "
              f"rel-L2 uncond {errs[0]:.3e} cond {errs[1]:.3e} 3-step trajectory {errs[2]:.3e} {'OK' if ok else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not bool((errs < 5e-3).all()):
        sys.exit(1)


if __name__ == "__main__":
    main()
