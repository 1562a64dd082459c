"""Attention under SUSTAINED load (the power cap decides the clock, as inside a sampler step): our variants and torch SDPA
(cuDNN) alternate launch by launch for many rounds, so every candidate sees the same thermal / power state; medians over the
last rounds, with the SM clock and board power sampled by nvidia-smi during the run.

usage: python tools/attn_sustained_bench.py [--rounds 30]
"""
import argparse
import subprocess
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from landiff_b200 import ops  # noqa: E402
from tools.kernel_bench import flush_l2  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=30)
    ap.add_argument("--solo", type=int, default=1, help="also run each candidate alone for --rounds launches")
    a = ap.parse_args()
    dev = "cuda"
    B, H, N = 2, 30, 17776
    torch.manual_seed(0)
    q = torch.randn(B, H, N, 64, device=dev).bfloat16()
    k = torch.randn(B, H, N, 64, device=dev).bfloat16()
    v = torch.randn(B, H, N, 64, device=dev).bfloat16()
    out = torch.empty(B, N, H * 64, device=dev, dtype=torch.bfloat16)
    import os
    cands = {f"ours default (LD_ATTN_KP={os.environ.get('LD_ATTN_KP', 'unset')})": lambda: ops.attention(q, k, v, out=out, variant=0),
             "ours v2 (all MUFU)": lambda: ops.attention(q, k, v, out=out, variant=2),
             "ours v3 (5/16 poly)": lambda: ops.attention(q, k, v, out=out, variant=3),
             "ours v5 (default + truncating pack)": lambda: ops.attention(q, k, v, out=out, variant=5),
             "torch SDPA (cuDNN)": lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v)}
    flops = 4.0 * B * H * N * N * 64

    def smi():
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True).stdout.strip()
        return r

    def timed(fn):
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        fn()
        e_.record()
        return s_, e_

    print(f"interleaved, {a.rounds} rounds (no L2 flush: K/V of one head stay in L2 as in the real step's back-to-back launches)")
    ts = {n: [] for n in cands}
    for r in range(a.rounds):
        evs = [(n, timed(fn)) for n, fn in cands.items()]
        torch.cuda.synchronize()
        if r >= a.rounds // 3:
            for n, (s_, e_) in evs:
                ts[n].append(s_.elapsed_time(e_))
    print("  nvidia-smi clocks.sm, power.draw at the end:", smi())
    for n, t in ts.items():
        t.sort()
        ms = t[len(t) // 2]
        print(f"  {n:34s} {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s")
    if a.solo:
        print(f"each candidate alone, {a.rounds} back-to-back launches (its own power state)")
        for n, fn in cands.items():
            torch.cuda.synchronize()
            evs = [timed(fn) for _ in range(a.rounds)]
            torch.cuda.synchronize()
            state = smi()
            t = sorted(s_.elapsed_time(e_) for s_, e_ in evs[a.rounds // 3:])
            ms = t[len(t) // 2]
            print(f"  {n:34s} {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s   [{state}]")


if __name__ == "__main__":
    main()
