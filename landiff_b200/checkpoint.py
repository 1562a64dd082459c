"""Checkpoint adapter: LanDiff's diffusion-stage checkpoints -> the drop-in modules.

Two files feed the reference's DiT ("Checkpoint / resume"):
  * the CogVideoX-2b base weights, read by `ControlDiffWarp.__init__` itself (dit_video_concat.py:1176-1189; the
    drop-in's constructor does the same), and
  * the LanDiff engine checkpoint `<load>/<iter>/mp_rank_00_model_states.pt` (directory resolved through
    `<load>/latest`), a dict whose `"module"` entry holds the whole SATVideoDiffusionEngine: `model.*` is the
    ControlDiffWarp (`model.main_model.diffusion_model.*`, `model.control_model.diffusion_model.*`), the rest are the
    conditioner (T5) and first stage (VAE) — SAT `load_checkpoint`, dif_infer.py:140-147.
`load_engine_checkpoint` binds the `model.*` part to a drop-in ControlDiffWarp: every DiT tensor must be present with
the right shape (Appendix B of SURVEY.md is the key contract); the control net's `semantic_conditioner.*` tensors
belong to the reference's SemanticCond module, which stays on reference code, and are handed back to the caller.
`verify_md5` checks files against `ckpts/CHECKSUM.md5`-style lists.
"""
from __future__ import annotations

import hashlib
import os
from typing import Dict, List, Tuple

import torch

SEMANTIC_PREFIX = "control_model.diffusion_model.semantic_conditioner."


def resolve_checkpoint_path(load_dir: str) -> str:
    """SAT layout: `<load_dir>/latest` names the iteration directory holding mp_rank_00_model_states.pt."""
    if os.path.isfile(load_dir):
        return load_dir
    latest = os.path.join(load_dir, "latest")
    if os.path.isfile(latest):
        with open(latest) as f:
            it = f.read().strip()
        return os.path.join(load_dir, it, "mp_rank_00_model_states.pt")
    direct = os.path.join(load_dir, "mp_rank_00_model_states.pt")
    if os.path.isfile(direct):
        return direct
    raise FileNotFoundError(f"no SAT checkpoint under {load_dir} (expected `latest` or mp_rank_00_model_states.pt)")


def split_engine_state(module_state: Dict[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor], List[str]]:
    """-> (warp state dict with the `model.` prefix stripped, semantic-conditioner tensors, other top-level prefixes)."""
    warp_sd, sem_sd, others = {}, {}, set()
    for k, v in module_state.items():
        if not k.startswith("model."):
            others.add(k.split(".", 1)[0])
            continue
        k2 = k[len("model."):]
        if k2.startswith(SEMANTIC_PREFIX):
            sem_sd[k2[len(SEMANTIC_PREFIX):]] = v
        else:
            warp_sd[k2] = v
    return warp_sd, sem_sd, sorted(others)


def load_engine_checkpoint(warp, path: str, strict: bool = True) -> Dict[str, torch.Tensor]:
    """Load the `model.*` tensors of a LanDiff engine checkpoint into a drop-in ControlDiffWarp (in place, keeping
    each parameter's dtype/device).  Returns the semantic-conditioner state dict for the reference-side module.
    Raises KeyError / ValueError listing every missing, unexpected or mis-shaped tensor when `strict`."""
    ckpt = torch.load(resolve_checkpoint_path(path), map_location="cpu", weights_only=False)
    module_state = ckpt["module"] if isinstance(ckpt, dict) and "module" in ckpt else ckpt
    warp_sd, sem_sd, _ = split_engine_state(module_state)
    # The control net's semantic conditioner is a reference-side submodule (or nn.Identity): its keys are not part of the
    # DiT key contract checked here.  When a real conditioner is attached its tensors are loaded from `sem_sd` below.
    own_all = warp.state_dict()
    own = {k: v for k, v in own_all.items() if not k.startswith(SEMANTIC_PREFIX)}
    missing = [k for k in own if k not in warp_sd]
    unexpected = [k for k in warp_sd if k not in own]
    bad_shape = [f"{k}: checkpoint {tuple(warp_sd[k].shape)} vs module {tuple(own[k].shape)}"
                 for k in own if k in warp_sd and tuple(warp_sd[k].shape) != tuple(own[k].shape)]
    if strict and (missing or unexpected):
        raise KeyError(f"checkpoint does not match the DiT key contract: missing {missing[:8]}{'...' if len(missing) > 8 else ''} "
                       f"unexpected {unexpected[:8]}{'...' if len(unexpected) > 8 else ''}")
    if bad_shape:
        raise ValueError("shape mismatch: " + "; ".join(bad_shape[:8]))
    with torch.no_grad():
        for k, p in own.items():
            if k in warp_sd:
                p.copy_(warp_sd[k].to(dtype=p.dtype))
    cond = getattr(getattr(getattr(warp, "control_model", None), "diffusion_model", None), "semantic_conditioner", None)
    if cond is not None and not isinstance(cond, torch.nn.Identity) and len(list(cond.state_dict().keys())) > 0:
        res = cond.load_state_dict(sem_sd, strict=False)
        if strict and (res.missing_keys or res.unexpected_keys):
            raise KeyError(f"semantic conditioner tensors do not match: missing {res.missing_keys[:8]} "
                           f"unexpected {res.unexpected_keys[:8]}")
    return sem_sd


def verify_md5(checksum_file: str, root: str = None) -> Dict[str, bool]:
    """`md5sum`-format list (`<hex>  <relative path>` per line) -> {path: ok}.  Missing files count as failures."""
    root = os.path.dirname(os.path.abspath(checksum_file)) if root is None else root
    out = {}
    with open(checksum_file) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            digest, rel = line.split(None, 1)
            rel = rel.lstrip("*").strip()
            path = os.path.join(root, rel)
            if not os.path.isfile(path):
                out[rel] = False
                continue
            h = hashlib.md5()
            with open(path, "rb") as g:
                for block in iter(lambda: g.read(1 << 22), b""):
                    h.update(block)
            out[rel] = h.hexdigest() == digest.lower()
    return out
