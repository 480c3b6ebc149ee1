#!/usr/bin/env python
"""BASELINE.json configs 4 and 5 on one B200 (not the bench contract - bench.py measures configs[2]):

  config 5: delta x AC-count sweep at 1080p - capacity, PSNR(gray, stego), device-timed embed / extract
            throughput, and at every point a bit-exactness check of one frame against the C oracle
            (stego pixels and extracted bits, i.e. including the bits the reference itself gets wrong);
  config 4: 3840x2160 frames, 63 and 10 AC, embed + extract round trip (75 frames = one GPU's share of
            the 600-frame batch on 8 GPUs).

Usage: python profiles/sweep.py > profiles/r2_sweep.jsonl     (GPU box; prints one JSON object per point)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import svs_b200                                            # noqa: E402
from oracle import c_oracle as oc                          # noqa: E402  (checker only)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def point(frames, delta, n, check_frames=1):
    f, h, w = frames.shape[:3]
    cap = svs_b200.capacity_bits(h, w, n)
    g = torch.Generator(device="cuda").manual_seed(1000 * n + int(delta * 8))
    payload = torch.randint(0, 256, ((f * cap + 7) // 8 + 8,), dtype=torch.uint8, device="cuda", generator=g)
    res = svs_b200.embed_frames(frames, payload, f * cap, delta, n, want_sse=True)
    ext = svs_b200.extract_frames(res.stego, delta, n)
    stego = torch.empty_like(res.stego)
    bits = torch.empty((f, svs_b200.bits_row_bytes(h, w, n)), dtype=torch.uint8, device="cuda")
    t_embed = timed(lambda: svs_b200.embed_frames(frames, payload, f * cap, delta, n, out=stego))
    t_extract = timed(lambda: svs_b200.extract_frames(stego, delta, n, out=bits))
    # parity of the first frame(s) against the oracle
    fr = frames[:check_frames].cpu().numpy()
    pk = payload.cpu().numpy()
    s0, _, _ = oc.embed_frames(fr, pk, f * cap, delta, n, threads=oc.max_threads(), want_gray=False)
    same_px = bool(np.array_equal(s0, res.stego[:check_frames].cpu().numpy()))
    want = oc.extract_frames(s0, delta, n, threads=oc.max_threads())
    same_bits = bool(np.array_equal(want, ext[:check_frames].cpu().numpy()))
    # the reference's own round-trip errors: extracted vs embedded bits
    sent = np.unpackbits(pk[:(check_frames * cap + 7) // 8])[:check_frames * cap]
    got = np.unpackbits(want, axis=1)[:, :cap].reshape(-1)
    sse = res.sse.to(torch.float64)
    psnr = float((10.0 * torch.log10(255.0 ** 2 * h * w / sse.clamp(min=1))).mean())
    return {"height": h, "width": w, "frames": f, "delta": delta, "num_ac": n, "capacity_bits_per_frame": cap,
            "psnr_db": round(psnr, 2), "embed_ms": round(t_embed, 4), "extract_ms": round(t_extract, 4),
            "embed_frames_per_s": round(f / t_embed * 1e3), "extract_frames_per_s": round(f / t_extract * 1e3),
            "round_trip_mpixel_per_s": round(f * h * w / (t_embed + t_extract) / 1e3),
            "stego_pixels_identical_to_oracle": same_px, "extracted_bits_identical_to_oracle": same_bits,
            "reference_wrong_bits_in_checked_frames": int((sent != got).sum())}


def main():
    torch.cuda.set_device(0)
    g = torch.Generator(device="cuda").manual_seed(7)
    frames = torch.randint(64, 192, (64, 1080, 1920, 3), dtype=torch.uint8, device="cuda", generator=g)
    for delta in (1, 2, 3, 4, 6, 8, 10, 16, 20, 32, 50, 100):
        for n in (1, 10, 32, 63):
            print(json.dumps(dict(config=5, **point(frames, delta, n))), flush=True)
    del frames
    frames = torch.randint(64, 192, (75, 2160, 3840, 3), dtype=torch.uint8, device="cuda", generator=g)
    for n in (63, 10):
        print(json.dumps(dict(config=4, **point(frames, 20, n))), flush=True)


if __name__ == "__main__":
    main()
