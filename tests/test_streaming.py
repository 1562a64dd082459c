"""Streaming chunk driver (BASELINE config 5; landiff_b200/streaming.py): frame bookkeeping and the prefix hand-off on
CPU with a stub sampler, and on the GPU the whole loop (drop-in network + fused sampler with fixed_frames) against the
CPU oracle's sampler (OracleSampler(fixed_frames=...), which restates sampling.py:800-817, 834-835) on the same noise."""
import dataclasses

import pytest
import torch

from landiff_b200 import streaming as S


def test_plan_bookkeeping_matches_the_yaml_comment():
    # ...video_vq.yaml:213,231: "49 frames, 13 latent, prefix_length=7, gen 1+4*6=25 frames"
    p = S.StreamPlan(n_chunks=3)
    assert (p.chunk_frames, p.prefix_frames, p.new_frames) == (13, 7, 6)
    assert p.total_frames == 13 + 2 * 6 == 25
    assert p.video_frames() == 97          # 49 + 2 * 24
    assert [p.chunk_span(k) for k in range(3)] == [(0, 13), (6, 19), (12, 25)]
    for bad in (dict(n_chunks=0), dict(n_chunks=2, prefix_frames=0), dict(n_chunks=2, prefix_frames=13)):
        with pytest.raises(ValueError):
            S.StreamPlan(**bad)


def test_start_noise_overwrites_the_prefix_frames():
    g = torch.Generator().manual_seed(0)
    prefix = torch.full((1, 2, 3, 4, 4), 7.0)
    x = S.start_noise((1, 5, 3, 4, 4), prefix, "cpu", generator=g)
    assert torch.equal(x[:, :2], prefix) and x[:, 2:].abs().max() < 6 and x[:, 2:].std() > 0.5
    with pytest.raises(ValueError):
        S.start_noise((1, 5, 3, 4, 4), torch.zeros(1, 2, 3, 4, 5), "cpu")


class _StubSampler:
    """Adds (chunk index + 1) to the free frames and keeps the first `fixed_frames` frames, like the real sampler."""
    calls = []

    def __init__(self, fixed_frames):
        self.fixed_frames = fixed_frames

    def sample(self, network, x, cond, uc, cfg_group=None, noise_fn=None):
        k = len(_StubSampler.calls)
        _StubSampler.calls.append((self.fixed_frames, x.clone(), network()))
        out = torch.full_like(x, float(k + 1))
        out[:, :self.fixed_frames] = x[:, :self.fixed_frames]
        return out


def test_stream_loop_hands_the_last_frames_on_and_stitches():
    _StubSampler.calls = []
    plan = S.StreamPlan(n_chunks=3, chunk_frames=5, prefix_frames=2)
    feats = [torch.full((1, 5, 1, 2, 2), 10.0 * (k + 1)) for k in range(3)]
    current = {}
    out = S.sample_stream(lambda: current["feat"][0, 0, 0, 0, 0].item(), _StubSampler, plan, (1, 2, 2), {}, {}, feats,
                          lambda f: current.__setitem__("feat", f), device="cpu")
    assert out.shape == (1, plan.total_frames, 1, 2, 2)
    # chunk k contributes value k+1 on its new frames; its prefix frames carry chunk k-1's values
    assert out[0, :, 0, 0, 0].tolist() == [1.0] * 5 + [2.0] * 3 + [3.0] * 3
    fixed, starts, sem_seen = zip(*_StubSampler.calls)
    assert fixed == (0, 2, 2)
    assert sem_seen == (10.0, 20.0, 30.0)                      # each chunk saw its own semantic feature
    assert torch.equal(starts[1][:, :2], torch.full((1, 2, 1, 2, 2), 1.0))
    assert torch.equal(starts[2][:, :2], torch.full((1, 2, 1, 2, 2), 2.0))
    with pytest.raises(ValueError):
        S.sample_stream(lambda: 0, _StubSampler, plan, (1, 2, 2), {}, {}, feats[:2], lambda f: None, device="cpu")


@pytest.mark.gpu
def test_streaming_against_oracle_sampler():
    from landiff_b200 import dit
    from landiff_b200.factory import DiTShape, build_warp
    from landiff_b200.sampling import VPSDEDPMPP2MSampler
    from oracle import dit_oracle as O

    steps = 6
    cfg_o = dataclasses.replace(O.TINY, latent_t=4)
    cfg_p = DiTShape(hidden_size=128, num_heads=2, main_layers=2, control_layers=1, time_embed_dim=64, text_hidden=64,
                     text_length=6, latent_t=4, latent_h=8, latent_w=12)
    plan = S.StreamPlan(n_chunks=3, chunk_frames=4, prefix_frames=2)
    sdc = O.random_state_dict(cfg_o, True, seed=20, strong=True)
    sdm = O.random_state_dict(cfg_o, False, seed=21, strong=True)
    g = torch.Generator().manual_seed(3)
    ctx = (torch.randn(1, cfg_o.text_length, cfg_o.text_hidden, generator=g) * 0.2).bfloat16().float()
    feats = [(torch.randn(1, 4, 16, 8, 12, generator=g) * 0.1).bfloat16().float() for _ in range(3)]

    # ---- oracle: the same loop restated with OracleSampler (fp32 CPU)
    gen = torch.Generator().manual_seed(9)
    f32 = lambda sd: O.cast_state_dict(sd, torch.float32)
    pieces, prefix = [], None
    for k in range(3):
        x0 = torch.randn(1, 4, 16, 8, 12, generator=gen)
        if prefix is not None:
            x0 = torch.cat([prefix, x0[:, 2:]], dim=1)
        smp = O.OracleSampler(num_steps=steps, fixed_frames=0 if prefix is None else 2)
        net = lambda x2, t2, c2, k=k: O.warp_forward(f32(sdc), f32(sdm), cfg_o, x2, t2, c2, feats[k])
        z = smp(net, x0, ctx, torch.zeros_like(ctx), gen)
        pieces.append(z if k == 0 else z[:, 2:])
        prefix = z[:, 2:].clone()
    ref = torch.cat(pieces, dim=1)

    # ---- CUDA path through the driver, same noise stream
    warp = build_warp(cfg_p, device="cuda", sd_ctrl=sdc, sd_main=sdm)
    gen2 = torch.Generator().manual_seed(9)
    noise = lambda t: torch.randn(t.shape, generator=gen2).to(t.device)

    def register(f):
        dit.InferValueRegistry.clear()
        dit.InferValueRegistry.register("semantic_feature", f.cuda())

    chunks = {}
    out = S.sample_stream(warp, lambda ff: VPSDEDPMPP2MSampler(num_steps=steps, device="cuda", fixed_frames=ff), plan,
                          (16, 8, 12), {"crossattn": ctx.cuda().bfloat16()}, {"crossattn": torch.zeros_like(ctx).cuda().bfloat16()},
                          feats, register, device="cuda", noise_fn=noise,
                          chunk_callback=lambda k, z: chunks.__setitem__(k, z.float().cpu().clone()))
    torch.cuda.synchronize()
    dit.InferValueRegistry.clear()
    assert out.shape == (1, plan.total_frames, 16, 8, 12)
    # the fixed prefix of chunk k+1 is bit-identical to the tail of chunk k
    for k in (1, 2):
        assert torch.equal(chunks[k][:, :2], chunks[k - 1][:, 2:])
    r = ((out.float().cpu().double() - ref.double()).norm() / ref.double().norm()).item()
    assert r <= 2e-2, f"streamed latent rel-L2 {r:.3e} vs the oracle loop"


@pytest.mark.gpu
def test_full_shape_stream_chunks_against_oracle_sampler():
    """BASELINE config 5 at the FULL latent shape (13 x 16 x 60 x 90 per chunk, 7-latent-frame fixed prefix, 50 steps per
    chunk, 3 chained chunks): the streaming driver + fused sampler update on the GPU against the oracle sampler loop on the
    CPU, same noise stream.  The network is the analytic stand-in the reference-generated sampler golden uses
    (oracle/make_golden.py: timestep-, conditioning- and batch-row-dependent, bf16 output) so that the CPU side finishes in
    seconds; the network itself is checked at this shape by test_full_shape_step_against_oracle_fp32_on_gpu."""
    from landiff_b200.sampling import VPSDEDPMPP2MSampler
    from oracle import dit_oracle as O
    from oracle.make_golden import toy_network

    plan = S.StreamPlan(n_chunks=3, chunk_frames=13, prefix_frames=7)
    C, H, W = 16, 60, 90
    g = torch.Generator().manual_seed(4)
    ctx = torch.randn(1, 8, 16, generator=g)
    feats = [torch.zeros(1, 13, C, H, W) for _ in range(3)]       # the stand-in network does not read the registry

    gen = torch.Generator().manual_seed(9)
    pieces, prefix = [], None
    for k in range(3):
        x0 = torch.randn(1, 13, C, H, W, generator=gen)
        if prefix is not None:
            x0 = torch.cat([prefix, x0[:, 7:]], dim=1)
        smp = O.OracleSampler(num_steps=50, fixed_frames=0 if prefix is None else 7)
        z = smp(lambda x2, t2, c2: toy_network(x2, t2, {"crossattn": c2}), x0, ctx, torch.zeros_like(ctx), gen)
        pieces.append(z if k == 0 else z[:, 7:])
        prefix = z[:, 6:].clone()
    ref = torch.cat(pieces, dim=1)

    gen2 = torch.Generator().manual_seed(9)
    noise = lambda t: torch.randn(t.shape, generator=gen2).to(t.device)
    chunks = {}
    out = S.sample_stream(toy_network, lambda ff: VPSDEDPMPP2MSampler(num_steps=50, device="cuda", fixed_frames=ff), plan,
                          (C, H, W), {"crossattn": ctx.cuda()}, {"crossattn": torch.zeros_like(ctx).cuda()}, feats,
                          lambda f: None, device="cuda", noise_fn=noise,
                          chunk_callback=lambda k, z: chunks.__setitem__(k, z.float().cpu().clone()))
    torch.cuda.synchronize()
    assert out.shape == (1, plan.total_frames, C, H, W) == (1, 25, 16, 60, 90)
    for k in (1, 2):   # the fixed prefix of chunk k+1 is bit-identical to the last 7 latent frames of chunk k
        assert torch.equal(chunks[k][:, :7], chunks[k - 1][:, 6:])
    r = ((out.float().cpu().double() - ref.double()).norm() / ref.double().norm()).item()
    assert r <= 5e-3, f"full-shape streamed latent rel-L2 {r:.3e} vs the oracle sampler loop"
