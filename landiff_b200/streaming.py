"""Streaming long-video diffusion (BASELINE config 5): chunk-by-chunk generation of 49-frame chunks, each conditioned
on the last latent frames of the previous chunk and on its own semantic features.

The reference ships the hooks but no chunk loop ("Long-context"):
  * `DiffusionInferenceWrapper.forward(..., vae_feature_prefix=...)` registers the chunk's semantic inputs in the
    process-global `InferValueRegistry` and passes the prefix on (dif_infer.py:152-234);
  * `SATVideoDiffusionEngine.sample(prefix=...)` overwrites the first `prefix.shape[1]` latent frames of the start
    noise with the clean prefix (diffusion_video.py:287-288);
  * `VPSDEDPMPP2MSampler(fixed_frames=k)` re-imposes those frames before every denoiser call and once more at the end
    (sampling.py:800-817, 834-835); the YAML comment fixes the intended numbers: 13 latent frames per chunk,
    prefix_length 7, i.e. 6 new latent frames (24 video frames) per follow-up chunk (...video_vq.yaml:213,231).
This module assembles those hooks into the loop.  It is host logic only: the per-step work is the drop-in network and
the fused sampler update of `landiff_b200.sampling`, on one GPU or across the CFG x ring layout of `parallel.py`
(chunks are sequentially dependent, so they add no parallel axis).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import torch


@dataclass
class StreamPlan:
    """Frame bookkeeping of a streamed video: every chunk has `chunk_frames` latent frames; chunk k > 0 re-uses the
    last `prefix_frames` latent frames of chunk k-1 as its fixed prefix and contributes the remaining ones."""
    n_chunks: int
    chunk_frames: int = 13
    prefix_frames: int = 7

    def __post_init__(self):
        if self.n_chunks < 1:
            raise ValueError("a stream needs at least one chunk")
        if not 0 < self.prefix_frames < self.chunk_frames:
            raise ValueError(f"prefix_frames must be in (0, {self.chunk_frames}), got {self.prefix_frames}")

    @property
    def new_frames(self) -> int:
        return self.chunk_frames - self.prefix_frames

    @property
    def total_frames(self) -> int:
        return self.chunk_frames + (self.n_chunks - 1) * self.new_frames

    def chunk_span(self, k: int):
        """Latent-frame interval [start, stop) of the stream covered by chunk k."""
        start = k * self.new_frames
        return start, start + self.chunk_frames

    def video_frames(self, time_compression: int = 4) -> int:
        """Decoded frames of the whole stream for a causal VAE with the given temporal compression (1 + 4 (T-1))."""
        return 1 + time_compression * (self.total_frames - 1)


def start_noise(shape, prefix: Optional[torch.Tensor], device, generator: Optional[torch.Generator] = None,
                noise_fn: Optional[Callable] = None) -> torch.Tensor:
    """diffusion_video.py:266-288: fp32 N(0,1) start latent whose first prefix.shape[1] frames are the clean prefix."""
    if noise_fn is not None:
        x = noise_fn(torch.empty(shape, dtype=torch.float32, device=device))
    else:
        x = torch.randn(shape, generator=generator, device=device, dtype=torch.float32)
    if prefix is not None:
        if prefix.shape[0] != shape[0] or tuple(prefix.shape[2:]) != tuple(shape[2:]):
            raise ValueError(f"prefix {tuple(prefix.shape)} does not match the chunk latent {tuple(shape)}")
        x = torch.cat([prefix.to(x.dtype), x[:, prefix.shape[1]:]], dim=1)
    return x


def sample_stream(network: Callable, make_sampler: Callable[[int], object], plan: StreamPlan, latent_shape: Sequence[int],
                  cond: Dict, uc: Dict, semantic_features: Sequence[torch.Tensor], register: Callable[[torch.Tensor], None],
                  device="cuda", cfg_group=None, noise_fn: Optional[Callable] = None,
                  chunk_callback: Optional[Callable[[int, torch.Tensor], None]] = None) -> torch.Tensor:
    """Generate `plan.n_chunks` chunks and return the stitched latent [1, plan.total_frames, C, H, W] (fp32).

      network            the drop-in ControlDiffWarp (or any callable with its signature)
      make_sampler(k)    -> a VPSDEDPMPP2MSampler with fixed_frames = k (0 for the first chunk, prefix_frames after)
      latent_shape       (C, H, W) of one latent frame
      semantic_features  one [1, chunk_frames, C, H, W] tensor per chunk (what SemanticCond produces per chunk)
      register(feat)     makes `feat` the current semantic feature (InferValueRegistry.clear + register in the
                         reference, dif_infer.py:161-170)
    """
    if len(semantic_features) != plan.n_chunks:
        raise ValueError(f"{plan.n_chunks} chunks need {plan.n_chunks} semantic feature tensors, got {len(semantic_features)}")
    C, H, W = latent_shape
    shape = (1, plan.chunk_frames, C, H, W)
    pieces: List[torch.Tensor] = []
    prefix = None
    for k in range(plan.n_chunks):
        feat = semantic_features[k]
        if tuple(feat.shape[1:]) != shape[1:]:
            raise ValueError(f"semantic feature of chunk {k} has shape {tuple(feat.shape)}, expected [*, {shape[1:]}]")
        register(feat)
        x0 = start_noise(shape, prefix, device, noise_fn=noise_fn)
        sampler = make_sampler(0 if prefix is None else plan.prefix_frames)
        z = sampler.sample(network, x0, cond, uc, cfg_group=cfg_group, noise_fn=noise_fn)
        if chunk_callback is not None:
            chunk_callback(k, z)
        pieces.append(z if k == 0 else z[:, plan.prefix_frames:])
        prefix = z[:, plan.chunk_frames - plan.prefix_frames:].clone()
    return torch.cat(pieces, dim=1)
