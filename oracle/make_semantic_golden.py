"""TEST INFRASTRUCTURE ONLY.  Golden outputs of the REFERENCE's own semantic conditioner
(landiff/diffusion/semantic_models/condition.py + modules/vq_gan_blocks.py imported unchanged from /root/reference), fp32
on CPU, through `SemanticCond.forward(semantic_feature_before_upsample=...)` with the tokenizer replaced by nn.Identity
(the tokenizer's decoder is out of scope; everything after its features is what landiff_b200/semantic.py builds).
Only runnable in the build container.  Run:  python -m oracle.make_semantic_golden

  tests/golden/semantic_ref.pt   "shipped": the shipped decoder widths (768 -> 512 -> 128 -> 64 -> 16, 4 res blocks per level)
                                            on 2 frames of 6 x 9 features;
                                 "small":   widths 128 -> 256 -> 64 -> 64 -> 16, 1 res block, 3 frames of 5 x 7 features
                                 Weights and inputs are regenerated from the seeds by the tests (not stored).
"""
from __future__ import annotations

from pathlib import Path

import torch

from . import sat_shim
from . import semantic_oracle as S

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
CASES = {"shipped": dict(cfg=S.SHIPPED, wseed=21, xseed=22, B=1, T=2, h=6, w=9),
         "small": dict(cfg=S.SMALL, wseed=23, xseed=24, B=1, T=3, h=5, w=7)}


def reference_output(case) -> torch.Tensor:
    sat_shim.install()
    from landiff.diffusion.semantic_models.condition import SemanticCond

    cfg = case["cfg"]
    m = SemanticCond(**cfg.cond_kwargs("landiff.diffusion.semantic_models.modules.vq_gan_blocks.Decoder", torch.float32))
    m.load_state_dict(S.random_state_dict(cfg, case["wseed"]), strict=True)
    m.eval()
    x = S.features_for(cfg, case["xseed"], case["B"], case["T"], case["h"], case["w"])
    with torch.no_grad():
        return m(semantic_feature_before_upsample=x)


def main():
    blob = {}
    for tag, case in CASES.items():
        y = reference_output(case)
        blob[tag] = {"out": y.clone(), **{k: v for k, v in case.items() if k != "cfg"}}
        print(tag, tuple(y.shape), "rms", float(y.pow(2).mean().sqrt()))
    torch.save(blob, OUT / "semantic_ref.pt")


if __name__ == "__main__":
    main()
