#!/usr/bin/env python
"""Throughput of the N2 variants: embed with the fused GRAY2BGR store, extract from a
3-channel stego (what the receiver decodes from the FFV1 file), with the fused SSE (N3) - next to
the headline gray-stego path.  600 x 1080p frames, 63 AC, delta 20, CUDA events, best of 3.
Usage: python profiles/n2_throughput.py > profiles/r2_n2_throughput.txt   (GPU box)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import svs_b200   # noqa: E402

F, H, W, N, D = 600, 1080, 1920, 63, 20


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


g = torch.Generator(device="cuda").manual_seed(3)
frames = torch.randint(64, 192, (F, H, W, 3), dtype=torch.uint8, device="cuda", generator=g)
cap = svs_b200.capacity_bits(H, W, N)
payload = torch.randint(0, 256, (F * cap // 8 + 8,), dtype=torch.uint8, device="cuda", generator=g)
gray_out = torch.empty((F, H, W), dtype=torch.uint8, device="cuda")
bgr_out = torch.empty((F, H, W, 3), dtype=torch.uint8, device="cuda")
bits = torch.empty((F, svs_b200.bits_row_bytes(H, W, N)), dtype=torch.uint8, device="cuda")
px = H * W
rows = [
    ("embed BGR -> gray stego (headline)", lambda: svs_b200.embed_frames(frames, payload, F * cap, D, N, out=gray_out), 4 * px + cap // 8),
    ("embed BGR -> BGR stego (N2: fused GRAY2BGR)", lambda: svs_b200.embed_frames(frames, payload, F * cap, D, N, stego_channels=3, out=bgr_out), 6 * px + cap // 8),
    ("embed BGR -> gray stego + per-frame SSE (N3: fused epilogue)", lambda: svs_b200.embed_frames(frames, payload, F * cap, D, N, out=gray_out, want_sse=True), 4 * px + cap // 8),
    ("extract gray stego (headline)", lambda: svs_b200.extract_frames(gray_out, D, N, out=bits), px + cap // 8),
    ("extract BGR stego (N2: what the FFV1 reader delivers)", lambda: svs_b200.extract_frames(bgr_out, D, N, out=bits), 3 * px + cap // 8),
]
svs_b200.embed_frames(frames, payload, F * cap, D, N, out=gray_out)
svs_b200.embed_frames(frames, payload, F * cap, D, N, stego_channels=3, out=bgr_out)
for name, fn, bytes_per_frame in rows:
    ms = timed(fn)
    print("%-62s %7.3f ms / %d frames  %8.0f frames/s  %6.0f GB/s algorithmic" % (name, ms, F, F / ms * 1e3, F * bytes_per_frame / ms / 1e6))
ext = svs_b200.extract_frames(bgr_out, D, N)
assert torch.equal(ext.reshape(-1), payload[:F * cap // 8]), "round trip through the BGR stego lost bits"
print("round trip through the BGR stego returns the payload: ok")
