"""GPU-time breakdown of one full-shape sampler step on rank 0 of an N-rank run (torch.profiler, CUDA activities):
kernel busy time by name vs the step's wall time — shows where multi-GPU efficiency goes (short-kernel quantisation,
ring merges, exposed NCCL, idle gaps).  Run under torch.distributed.run like bench.py."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile
from landiff_b200 import dit, ops, parallel
from landiff_b200.factory import FULL, build_warp, random_init_
from landiff_b200.sampling import VPSDEDPMPP2MSampler

world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
layout = parallel.make_layout(world, rank)
sp_group = parallel.new_subgroups(layout)[0] if world > 1 else None
cfg = FULL
warp = build_warp(cfg, device=dev)
random_init_(warp, seed=0)
parallel.attach(warp, layout, sp_group, dev)
grp = parallel.CFGGroup(layout) if world > 1 else None
sampler = VPSDEDPMPP2MSampler(num_steps=50, device="cuda")
g = torch.Generator().manual_seed(1)
x = torch.randn(1, cfg.latent_t, 16, cfg.latent_h, cfg.latent_w, generator=g).to(dev)
ctx = (torch.randn(1, cfg.text_length, cfg.text_hidden, generator=g) * 0.2).bfloat16().to(dev)
dit.InferValueRegistry.clear()
dit.InferValueRegistry.register("semantic_feature", (torch.randn(1, cfg.latent_t, 16, cfg.latent_h, cfg.latent_w, generator=g) * 0.1).bfloat16().to(dev))
cond, uc = {"crossattn": ctx}, {"crossattn": torch.zeros_like(ctx)}
torch.manual_seed(42)
sampler.sample(warp, x, cond, uc, cfg_group=grp, max_steps=3)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sampler.sample(warp, x, cond, uc, cfg_group=grp, start_step=3, max_steps=2)
    e1.record()
    torch.cuda.synchronize()
if rank == 0:
    wall = e0.elapsed_time(e1) / 2
    rows = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            n = ev.name.split("(")[0].replace("void ", "")[:60]
            r = rows.setdefault(n, [0, 0.0])
            r[0] += 1
            r[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    busy = sum(v[1] for v in rows.values()) / 2 / 1e3
    print(f"world {world}: step wall {wall:.2f} ms; summed kernel time on rank 0 {busy:.2f} ms/step (streams overlap, so the sum may exceed wall)")
    for n, (c, t) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f"  {t / 2 / 1e3:8.3f} ms  {c // 2:5d} launches  {n}")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
