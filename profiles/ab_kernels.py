#!/usr/bin/env python
"""A/B timing of kernel families on one GPU: embed (BGR -> gray stego) and extract (gray stego ->
bits) over F synthetic 1080p frames, CUDA events, median of K launches after W warm-ups.

    python profiles/ab_kernels.py [--frames 600] [--families 2,5] [--ac 63,10] [--h 1080 --w 1920]
"""
import argparse
import json
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svs_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=600)
    ap.add_argument("--families", default="2,5")
    ap.add_argument("--ac", default="63")
    ap.add_argument("--delta", type=float, default=20)
    ap.add_argument("--h", type=int, default=1080)
    ap.add_argument("--w", type=int, default=1920)
    ap.add_argument("--iters", type=int, default=7)
    ap.add_argument("--tag", default="")
    ap.add_argument("--sse", action="store_true", help="embed with the fused per-frame SSE output (SIDE kernels)")
    ap.add_argument("--lib", default="", help="another build of libsvs_b200.so (e.g. variants/libsvs_variants.so)")
    a = ap.parse_args()
    if a.lib:
        svs_b200._native.use_library(os.path.abspath(a.lib))
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    frames = torch.randint(64, 192, (a.frames, a.h, a.w, 3), dtype=torch.uint8, device=dev, generator=g)
    L = svs_b200.lib()
    for n in [int(x) for x in a.ac.split(",")]:
        cap = svs_b200.capacity_bits(a.h, a.w, n)
        total = a.frames * cap
        payload = torch.randint(0, 256, ((total + 7) // 8 + 8,), dtype=torch.uint8, device=dev, generator=g)
        ref = None
        for fam in [int(x) for x in a.families.split(",")]:
            prev = L.svs_debug_kernel_family(fam)
            if prev < 0:
                print(json.dumps({"tag": a.tag, "family": fam, "skipped": "not in this build"}), flush=True)
                continue
            te, tx = [], []
            for it in range(3 + a.iters):
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
                res = svs_b200.embed_frames(frames, payload, total, a.delta, n, want_sse=a.sse)
                e1.record()
                bits = svs_b200.extract_frames(res.stego, a.delta, n)
                e2.record()
                torch.cuda.synchronize()
                if it >= 3:
                    te.append(e0.elapsed_time(e1))
                    tx.append(e1.elapsed_time(e2))
            L.svs_debug_kernel_family(prev)
            sig = (int(res.stego.to(torch.int64).sum().item()), int(bits.to(torch.int64).sum().item()))
            same = None
            if ref is None:
                ref = (res.stego.clone(), bits.clone())
            else:
                same = bool(torch.equal(ref[0], res.stego) and torch.equal(ref[1], bits))
            print(json.dumps({"tag": a.tag, "family": fam, "ac": n, "frames": a.frames, "hw": [a.h, a.w],
                              "embed_ms": round(statistics.median(te), 4), "extract_ms": round(statistics.median(tx), 4),
                              "embed_min": round(min(te), 4), "extract_min": round(min(tx), 4),
                              "same_as_first": same, "sig": sig}), flush=True)


if __name__ == "__main__":
    main()
