"""Attention tail split (wave quantisation) A/B: the per-rank attention shapes of the 1 / 2 / 4 / 8-GPU layouts, timed with
CUDA events, L2 flushed between iterations.  Variant 0 (tail split) against variant 6 (same kernel, one CTA per query block).

usage: python tools/attn_split_bench.py [--iters 7]
"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from landiff_b200 import ops  # noqa: E402
from tools.kernel_bench import flush_l2  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=7)
    a = ap.parse_args()
    dev = "cuda"
    H, NKV = 30, 17776
    torch.manual_seed(0)
    for name, B, nq in (("1 GPU (B=2, 17776 rows)", 2, 17776), ("2 GPUs (B=1, 17776 rows)", 1, 17776),
                        ("4 GPUs (B=1, 8888 rows)", 1, 8888), ("8 GPUs (B=1, 4444 rows)", 1, 4444)):
        q = torch.randn(B, H, nq, 64, device=dev).bfloat16()
        k = torch.randn(B, H, NKV, 64, device=dev).bfloat16()
        v = torch.randn(B, H, NKV, 64, device=dev).bfloat16()
        out = torch.empty(B, nq, H * 64, device=dev, dtype=torch.bfloat16)
        flops = 4.0 * B * H * nq * NKV * 64
        ctas = B * H * ((nq + 255) // 256)
        # the two variants alternate iteration by iteration: under the power cap a kernel timed second runs at lower clocks
        ts = {0: [], 6: []}
        ref = None
        for it in range(a.iters + 2):
            for variant in (6, 0):
                flush_l2()
                s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record()
                ops.attention(q, k, v, out=out, variant=variant)
                e_.record()
                torch.cuda.synchronize()
                if it >= 2:
                    ts[variant].append(s_.elapsed_time(e_))
                if variant == 6 and ref is None:
                    ref = out.clone()
        ms6, ms0 = sorted(ts[6])[len(ts[6]) // 2], sorted(ts[0])[len(ts[0]) // 2]
        err = ((out.float() - ref.float()).norm() / ref.float().norm()).item()
        print(f"  {name:28s} {ctas:5d} query blocks = {ctas / 148:6.2f} waves: no split {ms6:7.3f} ms {flops / ms6 / 1e9:7.1f} TF/s | "
              f"tail split {ms0:7.3f} ms {flops / ms0 / 1e9:7.1f} TF/s ({ms0 / ms6:.3f}x)  rel diff {err:.1e}", flush=True)


if __name__ == "__main__":
    main()
