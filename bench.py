"""bench.py — seconds per 49-frame 480x720 50-step CFG denoise (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]            ours: hand-written sm_100a kernels via the C-ABI
  python bench.py --impl reference [--gpus N --steps K --warmup W] the reference's CPU path (oracle port) on host cores

A "step" is ONE sampler step of the job: a CFG network evaluation (15-layer control DiT + 30-layer main DiT, batch
[uncond, cond], N = 17 776 tokens) plus the fused denoiser-scale + CFG + DPM++(2M) SDE update.  50 steps = one video,
so value = 50 x (timed seconds / K).  N = 1: both CFG rows on one GPU (BASELINE configs[1]); N = 2: CFG-parallel;
N = 4 / 8: CFG x ring sequence parallel (2 / 4).  Synthetic latents / text features / semantic features of the named
shapes, seeded random-init weights of the named architecture (no checkpoints exist offline).

Timing: W >= 3 untimed warm-up steps; K timed steps bracketed by barrier + cuda synchronize, CUDA events on the
launch stream, max over ranks.  The working set of one step (5.3 GB of weights + ~1.5 GB of activations per
layer) is far larger than the 126 MB L2, so no explicit flush is needed between steps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "seconds per 49-frame 480x720 50-step CFG denoise"
UNIT = "s"
SAMPLER_STEPS = 50
# algorithmic FLOPs: 2MNK per GEMM + 4 N^2 d per attention layer-sample
FLOP_PER_CFG_STEP_FULL = 3.6393e14
FLOP_PER_CFG_STEP_CONFIG1 = 7.793e12


def peaks():
    try:
        p = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, sustained bf16)"
    except Exception:
        return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def ncu_dram_bytes(csv_path, kernel_substr):
    """(dram__bytes_read.sum + dram__bytes_write.sum in bytes, source) of the first kernel whose name contains
    `kernel_substr` in a tools/ncu_summary.py CSV; (None, why) when the file or the columns are missing."""
    import csv

    try:
        rows = list(csv.reader(open(csv_path)))
        hdr, units = rows[0], rows[1]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            if kernel_substr in r[0]:
                return float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]], f"profiles/{Path(csv_path).name}"
        return None, f"{kernel_substr} not in {csv_path}"
    except Exception as e:  # noqa: BLE001
        return None, f"unavailable: {e}"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------- ours
def run_ours(args):
    import torch.distributed as dist

    from landiff_b200 import _C, dit, ops, parallel
    from landiff_b200.factory import FULL, build_warp, random_init_
    from landiff_b200.sampling import VPSDEDPMPP2MSampler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ops.device_check()
    layout = parallel.make_layout(world, rank)
    sp_group = None
    if world > 1:
        sp_group, _ = parallel.new_subgroups(layout)

    cfg = FULL
    warp = build_warp(cfg, device=dev)
    random_init_(warp, seed=0)
    parallel.attach(warp, layout, sp_group, dev)
    cfg_group = parallel.CFGGroup(layout) if world > 1 else None
    # The timed region launches every kernel eagerly so that the dominant kernel can be timed live with CUDA events around
    # each of its launches; the CUDA-graphed step (landiff_b200/graph.py, 1 and 2 GPUs) is measured after it as an
    # extra key — both are GPU-bound, the graph only removes the host's launch work.
    net = warp
    sampler = VPSDEDPMPP2MSampler(num_steps=SAMPLER_STEPS, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "stream3":
        run_stream3(args, world, rank, dev, layout, warp, cfg, cfg_group, barrier)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    g = torch.Generator().manual_seed(1)
    x_host = torch.randn(1, cfg.latent_t, cfg.in_channels, cfg.latent_h, cfg.latent_w, generator=g).pin_memory()
    ctx_host = (torch.randn(1, cfg.text_length, cfg.text_hidden, generator=g) * 0.2).to(torch.bfloat16).pin_memory()
    sem_host = (torch.randn(1, cfg.latent_t, cfg.in_channels, cfg.latent_h, cfg.latent_w, generator=g) * 0.1).to(torch.bfloat16)
    dit.InferValueRegistry.clear()
    dit.InferValueRegistry.register("semantic_feature", sem_host.to(dev))
    cond = {"crossattn": ctx_host.to(dev)}
    uc = {"crossattn": torch.zeros_like(cond["crossattn"])}
    torch.manual_seed(42)  # sampler noise stream, identical on every rank

    # attention-kernel timing hook (the dominant kernel): CUDA events on the launch stream around each launch
    attn_events = []

    def timed(orig):
        def f(*a, **k):
            if not timed.on:
                return orig(*a, **k)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = orig(*a, **k)
            e.record()
            attn_events.append((s, e))
            return r
        return f

    timed.on = False
    ops.attention = timed(ops.attention)                  # single-buffer launch (1 and 2 GPUs)
    ops.attention_shards = timed(ops.attention_shards)    # multi-shard launch of the sequence-parallel layouts
    timed_attention = timed

    def sample_k(x0, k, start=0):
        """k consecutive sampler steps of the 50-step schedule (wrapping around for k > 50)."""
        x, done = x0, 0
        while done < k:
            chunk = min(SAMPLER_STEPS - start, k - done)
            x = sampler.sample(net, x, cond, uc, cfg_group=cfg_group, start_step=start, max_steps=chunk)
            done += chunk
            start = 0
        return x

    x_dev = x_host.to(dev, non_blocking=True)
    barrier()
    sample_k(x_dev, max(args.warmup, 3))
    barrier()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = _C.LAUNCHES[0]
    timed_attention.on = True
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    sample_k(x_dev, args.steps)
    ev1.record()
    barrier()
    timed_attention.on = False
    ms = ev0.elapsed_time(ev1)
    launches = _C.LAUNCHES[0] - launches0
    clk = clocks.stop() if rank == 0 else None
    attn_ms = [s.elapsed_time(e) for s, e in attn_events]

    # e2e: same steps through the public API with HOST buffers: per step H2D of the step's inputs (latent x from pinned
    # memory, text features) and D2H of the step's result (the updated latent)
    x_out_host = torch.empty_like(x_host).pin_memory()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = min(args.steps, 10)
    e0.record()
    for i in range(e2e_steps):
        xd = x_host.to(dev, non_blocking=True)
        cd = {"crossattn": ctx_host.to(dev, non_blocking=True)}
        xo = sampler.sample(net, xd, cd, uc, cfg_group=cfg_group, start_step=i % SAMPLER_STEPS, max_steps=1)
        x_out_host.copy_(xo, non_blocking=True)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps

    # CUDA-graphed step: one graph replay + the fused sampler update per step
    graph_info = None
    if layout.sp_size == 1 and not args.no_graph:
        from landiff_b200.graph import GraphedWarp

        gw = GraphedWarp(warp)
        sampler.sample(gw, x_dev, cond, uc, cfg_group=cfg_group, start_step=0, max_steps=2)   # capture + one replay
        barrier()
        g_steps = min(args.steps, 10)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _C.LAUNCHES[0]
        h0 = time.perf_counter()
        g0.record()
        sampler.sample(gw, x_dev, cond, uc, cfg_group=cfg_group, start_step=0, max_steps=g_steps)
        g1.record()
        host_graph = (time.perf_counter() - h0) / g_steps * 1e3     # host time to ENQUEUE a step (no sync inside)
        barrier()
        l_graph = (_C.LAUNCHES[0] - l0) / g_steps
        graph_info = {"ms_per_step": round(g0.elapsed_time(g1) / g_steps, 3), "steps_timed": g_steps,
                      "host_enqueue_ms_per_step": round(host_graph, 3),
                      "c_abi_calls_per_step": l_graph, "c_abi_calls_per_step_eager": launches / args.steps,
                      "replays": gw.replays,
                      "what": "one CUDA-graph replay of the whole ControlDiffWarp forward + the fused sampler update per step; "
                              "host_enqueue = host time to enqueue a step (no sync inside the loop)"}

    status = ops.attention_status()
    t = torch.tensor([ms, e2e_ms, float(status)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, wait_timeouts = float(t[0]), float(t[1]), int(t[2])
    if rank == 0:
        ms_per_step = ms / args.steps
        value = ms_per_step * SAMPLER_STEPS / 1e3
        tf_peak, bw_peak, how = peaks()
        # dominant kernel: attention.  algorithmic FLOPs per launch = 4 * nq * nkv * 64 * (B*H) for this rank's launch
        b_rows = 2 if layout.cfg_size == 1 else 1
        nq = cfg.n_tok // layout.sp_size
        flop_per_launch = 4.0 * nq * cfg.n_tok * 64 * b_rows * cfg.num_heads   # this rank's queries x ALL keys (one launch)
        attn_avg = sum(attn_ms) / max(len(attn_ms), 1)
        # DRAM traffic of one attention launch: dram__bytes_read.sum + dram__bytes_write.sum of the `ncu --set full` capture
        # of the shipped kernel at the same shape and B = 1 (read from profiles/r2_ncu_attn5.csv); the launch is
        # independent per batch row, so B rows move B times that.  Algorithmic bytes (Q, K, V in, O out) = 273 MB per
        # row: K/V re-reads of the 70 query blocks of a head are served by L2.  Sequence-parallel shards: no capture.
        traffic, traffic_src = None, "no ncu capture at the sequence-parallel shard shape"
        if layout.sp_size == 1:
            traffic, traffic_src = ncu_dram_bytes(ROOT / "profiles" / "r2_ncu_attn5.csv", "attn5_kernel")
            traffic = traffic * b_rows if traffic is not None else None
        achieved = flop_per_launch / (attn_avg * 1e-3) / 1e12 if attn_avg > 0 else 0.0
        line = {
            "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 3), "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "LanDiff ControlDiffWarp (15-layer control + 30-layer main DiT, d=1920, 30 heads) "
                                   "49-frame 480x720 latent 13x16x60x90, 17776 tokens, CFG batch 2, DPM++2M SDE, "
                                   "value = 50 x per-step time",
                       "parallelism": {1: "single GPU", 2: "cfg2", 4: "cfg2 x ring-sp2", 8: "cfg2 x ring-sp4"}.get(world, f"cfg{layout.cfg_size} x sp{layout.sp_size}"),
                       "l2": "per-step working set (>6 GB) exceeds the 126 MB L2; no explicit flush",
                       "weights": "seeded random init N(0, 0.02^2)", "sampler_steps_per_video": SAMPLER_STEPS},
            "tensor_frac_of_peak_whole_step": round(FLOP_PER_CFG_STEP_FULL / (ms_per_step * 1e-3) / world / 1e12 / tf_peak, 4),
            "roofline": {"bound": "tensor", "kernel": "attn5_kernel (tcgen05 flash attention, head_dim 64: double-buffered scores, 16 softmax warps, Q in "
                                   "TMEM, P in place over S, row sums on the tensor core, 4/16 exponential pairs on the FMA pipe, last partial wave split over key ranges"
                                   + ("" if layout.sp_size == 1 else f"; one launch over {layout.sp_size} K/V shards, arrival flags polled in-kernel") + ")",
                         "achieved": round(achieved, 1), "peak": tf_peak, "unit": "TFLOP/s",
                         "frac": round(achieved / tf_peak, 4), "traffic": traffic,
                         "traffic_unit": f"bytes per launch (ncu dram read + write at B = 1 x batch rows, {traffic_src})",
                         "peak_source": how,
                         "launch_ms": round(attn_avg, 4), "launches_timed": len(attn_ms),
                         "share_of_step": round(sum(attn_ms) / ms, 4) if ms > 0 else None},
            "e2e": {"value": round(e2e_ms * SAMPLER_STEPS / 1e3, 4), "unit": UNIT,
                    "h2d_bytes_per_step": x_host.numel() * 4 + ctx_host.numel() * 2,
                    "d2h_bytes_per_step": x_host.numel() * 4, "steps_timed": e2e_steps},
            "gpu_launches": launches,
            "cuda_graph": graph_info,
            "shard_wait_timeouts": wait_timeouts,
            "clocks": clk,
        }
        if world == 1:
            line["cpu_baseline"] = cpu_baseline_sample()
            if not args.no_eager:
                eager = gpu_eager_baseline(warp, cfg, x_dev, torch.cat((uc["crossattn"], cond["crossattn"]), 0),
                                           dit.InferValueRegistry.get_value("semantic_feature"))
                eager["ours_over_eager"] = round(value / eager["value"], 4)
                line["gpu_eager_baseline"] = eager
                try:
                    line["semantic_conditioner"] = semantic_conditioner_leg(dev)
                except Exception as exc:  # noqa: BLE001  (informational leg: never fail the bench line)
                    line["semantic_conditioner"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------- config 5
def run_stream3(args, world, rank, dev, layout, warp, cfg, cfg_group, barrier):
    """BASELINE config 5: streaming long-video diffusion, 3 chained 49-frame chunks (13 latent frames each; chunks 2 and 3
    keep the last 7 latent frames of their predecessor as a fixed prefix and receive their own semantic features), 50
    DPM++ CFG steps per chunk, full shape — landiff_b200/streaming.py over the reference's hooks
    (sampling.py:800-835, diffusion_video.py:287-288, ...video_vq.yaml:213,231).  Prints its own JSON line."""
    import torch.distributed as dist

    from landiff_b200 import _C, dit, ops
    from landiff_b200.sampling import VPSDEDPMPP2MSampler
    from landiff_b200.streaming import StreamPlan, sample_stream

    plan = StreamPlan(n_chunks=3, chunk_frames=cfg.latent_t, prefix_frames=7)
    g = torch.Generator().manual_seed(5)
    ctx_host = (torch.randn(1, cfg.text_length, cfg.text_hidden, generator=g) * 0.2).to(torch.bfloat16).pin_memory()
    sem_host = [(torch.randn(1, cfg.latent_t, cfg.in_channels, cfg.latent_h, cfg.latent_w, generator=g) * 0.1).to(torch.bfloat16).pin_memory()
                for _ in range(plan.n_chunks)]
    out_host = torch.empty(1, plan.total_frames, cfg.in_channels, cfg.latent_h, cfg.latent_w).pin_memory()

    def register(feat):
        dit.InferValueRegistry.clear()
        dit.InferValueRegistry.register("semantic_feature", feat)

    def one_stream(steps_per_chunk):
        cond = {"crossattn": ctx_host.to(dev, non_blocking=True)}                     # H2D: text features
        uc = {"crossattn": torch.zeros_like(cond["crossattn"])}
        feats = [f.to(dev, non_blocking=True) for f in sem_host]                      # H2D: per-chunk semantic features
        mk = lambda k: VPSDEDPMPP2MSampler(num_steps=steps_per_chunk, device="cuda", fixed_frames=k)
        z = sample_stream(warp, mk, plan, (cfg.in_channels, cfg.latent_h, cfg.latent_w), cond, uc, feats, register, device=dev,
                          cfg_group=cfg_group)
        out_host.copy_(z, non_blocking=True)                                         # D2H: the stitched latent
        return z

    torch.manual_seed(42)
    barrier()
    one_stream(2)                    # warm-up: every kernel / buffer / peer mapping of the layout
    barrier()
    clocks = ClockSampler(dev.index or 0)
    if rank == 0:
        clocks.start()
    l0 = _C.LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    z = one_stream(SAMPLER_STEPS)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _C.LAUNCHES[0] - l0
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms, float(ops.attention_status()), float(torch.isfinite(z).all())], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        n_steps = plan.n_chunks * SAMPLER_STEPS
        line = {"metric": "seconds per streamed 3 x 49-frame 480x720 video (3 chained chunks, 50-step CFG denoise each)",
                "value": round(float(t[0]) / 1e3, 4), "unit": UNIT, "n_gpus": world, "steps": n_steps, "warmup": 6,
                "ms_per_step": round(float(t[0]) / n_steps, 3), "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "stream3: BASELINE config 5 — 3 chained chunks of 13 latent frames (49 video frames) at "
                                       "480x720, 7-latent-frame fixed prefix from the previous chunk, per-chunk semantic features, "
                                       f"{plan.total_frames} latent frames = {plan.video_frames()} video frames in total",
                           "parallelism": {1: "single GPU", 2: "cfg2", 4: "cfg2 x sp2", 8: "cfg2 x sp4"}.get(world, "?"),
                           "l2": "per-step working set (>6 GB) exceeds the 126 MB L2; no explicit flush"},
                "e2e": {"value": round(float(t[0]) / 1e3, 4), "unit": UNIT,
                        "h2d_bytes_per_step": (ctx_host.numel() * 2 + sum(f.numel() * 2 for f in sem_host)) // n_steps,
                        "d2h_bytes_per_step": out_host.numel() * 4 // n_steps,
                        "note": "the timed region IS end to end: host text / semantic features in, stitched latent out"},
                "gpu_launches": launches, "shard_wait_timeouts": int(t[1]), "finite": bool(t[2]), "clocks": clk}
        print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- B2: eager GPU
def gpu_eager_baseline(warp, cfg, x_dev, ctx2, sem, steps=3):
    """The reference module graph in PyTorch eager bf16 on the SAME GPU (cuBLAS GEMMs + F.scaled_dot_product_attention +
    ~25 elementwise/norm launches per block) — SURVEY.md section 2.1's "B2" bar: what a user of the reference gets on a
    B200 without this repo.  The graph is the oracle's restatement of dit_video_concat.py:540-664 / :872-1027 run on
    the warp's own bf16 parameters (no copies); N = 1 only, outside the timed region of the product arm."""
    from oracle import dit_oracle as O

    sd = warp.state_dict()
    pick = lambda prefix: {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    sdc, sdm = pick("control_model.diffusion_model."), pick("main_model.diffusion_model.")
    ocfg = O.OracleConfig(hidden_size=cfg.hidden_size, num_heads=cfg.num_heads, main_layers=cfg.main_layers,
                          control_layers=cfg.control_layers, time_embed_dim=cfg.time_embed_dim, text_hidden=cfg.text_hidden,
                          text_length=cfg.text_length, latent_t=cfg.latent_t, latent_h=cfg.latent_h, latent_w=cfg.latent_w,
                          in_channels=cfg.in_channels, interp=cfg.interp)
    x2 = torch.cat([x_dev, x_dev])
    t2 = torch.full((2,), 519.0, device=x_dev.device)
    run = lambda: O.warp_forward(sdc, sdm, ocfg, x2, t2, ctx2, sem)
    run()
    torch.cuda.synchronize()
    clocks = ClockSampler(x_dev.device.index or 0)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = run()
    e1.record()
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1) / steps
    del out
    torch.cuda.empty_cache()
    return {"value": round(ms * SAMPLER_STEPS / 1e3, 4), "unit": UNIT, "ms_per_step": round(ms, 3), "steps_timed": steps,
            "what": "reference module graph (oracle restatement), PyTorch eager bf16 on this GPU: cuBLAS + SDPA, same shape, "
                    "CFG batch 2, network evaluation only (the eager sampler update adds < 0.2 ms)",
            "clocks": clk}


# ---------------------------------------------------------------------------------------------------- row f2
def semantic_conditioner_leg(dev, frames=13, h=30, w=45, iters=5):
    """Once-per-video cost of the semantic conditioner's conv decoder (landiff_b200/semantic.py)
    at the shipped widths on the 480x720 feature grid, next to the same graph in PyTorch eager bf16 (cuDNN).  Informational
    (N = 1, outside the timed region): the headline metric is the 50-step denoise."""
    import torch.nn.functional as F

    from landiff_b200.semantic import SemanticCond

    dec = dict(z_channels=768, resolution=16, in_channels=512, out_ch=64, ch=512, ch_mult=[0.25, 1], num_res_blocks=4,
               attn_resolutions=[], dropout=0.0, use_mid_attention=False, upsample_type="pixelshuffle")
    torch.manual_seed(7)
    m = SemanticCond(semantic_model_config={"target": "torch.nn.Identity"}, upsample_model_config={"target": "-", "params": dec},
                     dtype=torch.bfloat16, out_dim=64, target_dim=16, zero_init_conv_out=False).to(dev)
    feat = torch.randn(1, frames, 768, h, w, device=dev).bfloat16()
    sd = {k: v for k, v in m.state_dict().items()}

    def eager():     # vq_gan_blocks.py:577-606 + condition.py:131-136 with torch ops on the same parameters
        gn = lambda x, n: F.group_norm(x, 32, sd[n + ".weight"], sd[n + ".bias"], eps=1e-6)
        conv = lambda x, n, p=1: F.conv2d(x, sd[n + ".weight"], sd[n + ".bias"], padding=p)
        sw = lambda x: x * torch.sigmoid(x)

        def res(x, n):
            hh = conv(sw(gn(x, n + ".norm1")), n + ".conv1")
            hh = conv(sw(gn(hh, n + ".norm2")), n + ".conv2")
            return (conv(x, n + ".nin_shortcut", 0) if n + ".nin_shortcut.weight" in sd else x) + hh

        u = "upsample_model."
        x = conv(feat[0], u + "conv_in")
        x = res(res(x, u + "mid.block_1"), u + "mid.block_2")
        for j in range(5):
            x = res(x, f"{u}up.1.block.{j}")
        x = conv(F.pixel_shuffle(x, 2), u + "up.1.upsample.conv")
        for j in range(5):
            x = res(x, f"{u}up.0.block.{j}")
        return conv(conv(sw(gn(x, u + "norm_out")), u + "conv_out"), "conv_out")

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            y = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters, y

    ours_ms, y = timed(lambda: m(semantic_feature_before_upsample=feat))
    eager_ms, y_ref = timed(eager)
    rel = ((y[0].float() - y_ref.float()).norm() / y_ref.float().norm()).item()
    del m, sd, y, y_ref
    torch.cuda.empty_cache()
    return {"ms_per_video": round(ours_ms, 3), "eager_bf16_ms_per_video": round(eager_ms, 3),
            "ours_over_eager": round(ours_ms / eager_ms, 3),
            "rel_l2_between_the_two_bf16_runs": round(rel, 5),   # each is ~1e-2 from fp32 (tests/test_semantic.py has the fp32 gates)
            "what": f"SemanticCond upsample path, {frames} frames of 768 x {h} x {w} features -> 16 x {2 * h} x {2 * w}: implicit-GEMM "
                    "convolutions (TMA im2col gather) + GroupNorm kernels; beside it the same graph with torch ops (cuDNN) on the "
                    "same bf16 parameters"}


# ---------------------------------------------------------------------------------------------------- CPU side
CPU_THREADS = 16   # fixed thread count of the CPU legs (fewer if the box has fewer cores), so runs are comparable


def cpu_baseline_sample():
    """The oracle port (plain PyTorch fp32, the reference module graph restated) on the host cores: ONE full-depth
    (15 + 30 layers, d = 1920) CFG step at BASELINE config 1 (5 frames 240x352, N = 886 tokens).  `value` extrapolates
    that measurement to the full job (N = 17 776 tokens, 50 steps) by algorithmic FLOPs and is labelled as such: the CPU
    cannot run the full shape within the bench's time budget (~10 min per step)."""
    from oracle import dit_oracle as O

    cores = min(CPU_THREADS, os.cpu_count() or 1)
    torch.set_num_threads(cores)
    cfg = O.CONFIG1
    sdc = O.cast_state_dict(O.random_state_dict(cfg, True, seed=10), torch.float32)
    sdm = O.cast_state_dict(O.random_state_dict(cfg, False, seed=11), torch.float32)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, cfg.latent_t, 16, cfg.latent_h, cfg.latent_w, generator=g)
    ctx = torch.randn(2, cfg.text_length, cfg.text_hidden, generator=g) * 0.2
    sem = torch.randn(1, cfg.latent_t, 16, cfg.latent_h, cfg.latent_w, generator=g) * 0.1
    t = torch.tensor([519.0, 519.0])
    t0 = time.perf_counter()
    O.warp_forward(sdc, sdm, cfg, x, t, ctx, sem)
    dt = time.perf_counter() - t0
    ratio = FLOP_PER_CFG_STEP_FULL / FLOP_PER_CFG_STEP_CONFIG1
    return {"value": round(dt * ratio * SAMPLER_STEPS, 1), "unit": UNIT, "cores": cores, "kind": "port",
            "extrapolated": True, "same_config": False,
            "sample": f"one full-depth CFG step (batch 2, 15+30 layers, fp32) of the oracle port at BASELINE config 1 "
                      f"(N=886 tokens) took {dt:.2f} s on {cores} threads; value = that x {ratio:.1f} (algorithmic FLOPs "
                      f"of the full 17776-token step; ignores that attention grows quadratically) x 50 steps",
            "sample_seconds": round(dt, 3), "sample_tflops": round(FLOP_PER_CFG_STEP_CONFIG1 / dt / 1e12, 3)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference itself needs the
    un-vendored SwissArmyTransformer and cannot be installed offline).  Rank 0 only.  Every step is one MEASURED
    full-depth config-1 CFG step; `value` is its FLOP-extrapolation to the default arm's job and says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(max(args.warmup, 0) + args.steps):
        r = cpu_baseline_sample()
        if i >= args.warmup:
            vals.append(r)
    value = sum(v["value"] for v in vals) / len(vals)
    secs = sum(v["sample_seconds"] for v in vals) / len(vals)
    last = vals[-1]
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(value / SAMPLER_STEPS * 1e3, 1),
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "extrapolated": True, "same_config": False, "measured_step_seconds_config1": round(secs, 3),
            "config": {"workload": "MEASURED: full-depth (15+30 layers, d=1920) CFG step of the oracle port at BASELINE "
                                   "config 1 (N=886 tokens) on the host cores; value EXTRAPOLATES it to the default arm's job "
                                   "(N=17776, 50 steps) by algorithmic FLOPs — indicative only, not the same job"},
            "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": last["cores"], "kind": "port",
                             "extrapolated": True, "sample": last["sample"]},
            "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="denoise", choices=["denoise", "stream3"],
                    help="denoise: BASELINE configs 2-4 (one 49-frame video); stream3: config 5 (3 chained chunks)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-eager", action="store_true", help="skip the PyTorch-eager bf16 GPU baseline leg (N = 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
