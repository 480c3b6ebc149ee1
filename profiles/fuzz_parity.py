#!/usr/bin/env python
"""Time-bounded random parity run: the CUDA path against the plain-C oracle on random geometries,
deltas, coefficient counts, payload lengths / bit offsets, frame contents, strided input views and both stego layouts.

The -m gpu tests pin fixed cases (golden vectors, the delta x AC sweep, edge geometries); this
script spends a fixed number of seconds on cases nobody wrote down, with the emphasis on what
decides a bit: rounding ties and near-ties of the quantiser (small and fractional deltas, flat and
saturated content), clipping, payloads ending inside a frame.  Every case compares stego pixels,
the gray reference, bits_embedded, the fused SSE and the extracted bits of ALL frames.

    python profiles/fuzz_parity.py [--seconds 40] [--seed 1]        # one JSON line; exit 1 on any mismatch
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SIZES = [(8, 8), (16, 24), (64, 96), (120, 160), (240, 320), (480, 640), (720, 1280), (1080, 1920), (136, 264), (24, 1048)]
DELTAS_F32 = [1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16, 20, 25, 32, 50, 64, 100, 0.5, 0.75, 1.5, 2.5, 7.25, 12.125, 0.125, 0.0625]
DELTAS_OTHER = [0.1, 0.3, 7.3, 19.99, 33.333333333333336]          # not float32 values: scalar kernels (double product)


def make_frames(rng, kind, shape):
    f, h, w = shape[:3]
    if kind == "full":
        a = rng.integers(0, 256, shape, dtype=np.uint8)
    elif kind == "mid":
        a = rng.integers(64, 192, shape, dtype=np.uint8)
    elif kind == "saturated":                      # mostly 0 / 255 with sparse noise: clipping everywhere
        a = np.where(rng.random(shape) < 0.5, 0, 255).astype(np.uint8)
        m = rng.random(shape) < 0.1
        a[m] = rng.integers(0, 256, int(m.sum()), dtype=np.uint8)
    elif kind == "flat":                           # constant blocks + rare pixels: AC coefficients at or near 0 (ties at 0.5 delta multiples)
        base = rng.integers(0, 256, (f, (h + 7) // 8, (w + 7) // 8) + shape[3:], dtype=np.uint8)
        a = np.repeat(np.repeat(base, 8, 1), 8, 2)[:, :h, :w].copy()
        m = rng.random(shape) < 0.02
        a[m] = rng.integers(0, 256, int(m.sum()), dtype=np.uint8)
    else:                                          # smooth ramps + mild noise
        y, x = np.mgrid[0:h, 0:w]
        g = 30 + 190.0 * (x / max(1, w - 1)) * (y / max(1, h - 1)) + 15 * np.sin(x / 9.0) + 10 * np.cos(y / 5.0)
        g = g[None, :, :, None] if len(shape) == 4 else g[None]
        a = np.clip(g + rng.normal(0, 3, shape), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(a)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=40)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--dry", action="store_true", help="oracle side only (no GPU): checks the script itself")
    a = ap.parse_args()
    from oracle import c_oracle
    threads = c_oracle.max_threads()
    if not a.dry:
        import torch
        import svs_b200
        dev = torch.device("cuda:0")
    rng = np.random.default_rng(a.seed)
    t_end = time.time() + a.seconds
    cases = frames_total = 0
    bad = []
    by_kind = {}
    while time.time() < t_end:
        h, w = SIZES[rng.integers(len(SIZES))]
        px = h * w
        f = int(rng.integers(1, max(2, min(7, 6_000_000 // px + 1))))
        ch = 3 if rng.random() < 0.7 else 1
        shape = (f, h, w, 3) if ch == 3 else (f, h, w)
        kind = ["full", "mid", "saturated", "flat", "smooth"][rng.integers(5)]
        delta = float(DELTAS_F32[rng.integers(len(DELTAS_F32))] if rng.random() < 0.85 else DELTAS_OTHER[rng.integers(len(DELTAS_OTHER))])
        n = int(rng.integers(1, 64)) if rng.random() < 0.8 else [63, 64, 100, 10][rng.integers(4)]
        cap = (h // 8) * (w // 8) * min(n, 63)
        r = rng.random()
        total = f * cap if r < 0.5 else int(rng.integers(0, f * cap + 1)) if r < 0.9 else f * cap + 77
        off = int(rng.integers(0, 32)) if rng.random() < 0.3 else 0
        frames = make_frames(rng, kind, shape)
        payload = rng.integers(0, 256, (off + total + 7) // 8 + 1, dtype=np.uint8)
        want_stego, want_gray, want_nbits = c_oracle.embed_frames(frames, payload, total, delta, n, bit_offset=off, threads=threads)
        want_bits = c_oracle.extract_frames(want_stego, delta, n, threads=threads)
        want_sse = ((want_stego.astype(np.int64) - want_gray.astype(np.int64)) ** 2).reshape(f, -1).sum(1)
        cases += 1
        frames_total += f
        by_kind[kind] = by_kind.get(kind, 0) + 1
        if a.dry:
            continue
        d_frames = torch.from_numpy(frames).to(dev)
        if rng.random() < 0.3:                     # a strided view: rows and frames further apart than needed
            wide = torch.zeros((f, h + 8, w + 16) + ((3,) if ch == 3 else ()), dtype=torch.uint8, device=dev)
            view = wide[:, :h, 8:8 + w]
            view.copy_(d_frames)
            d_frames = view
        d_payload = torch.from_numpy(payload).to(dev)
        sc = 3 if rng.random() < 0.3 else 1       # gray stego, or replicated to BGR as the FFV1 writer takes it
        res = svs_b200.embed_frames(d_frames, d_payload, total, delta, n, bit_offset=off, stego_channels=sc,
                                    want_gray=True, want_bits_embedded=True, want_sse=True)
        got_bits = svs_b200.extract_frames(res.stego, delta, n)      # BGR stego: gray(g, g, g) = g
        torch.cuda.synchronize()
        nb = (cap + 7) // 8
        got_stego = res.stego.cpu().numpy()
        if sc == 3:
            assert (got_stego[..., 0] == got_stego[..., 1]).all() and (got_stego[..., 0] == got_stego[..., 2]).all()
            got_stego = got_stego[..., 0]
        diffs = {
            "stego_px": int((got_stego != want_stego).sum()),
            "gray_px": int((res.gray.cpu().numpy() != want_gray).sum()),
            "bits_embedded": int((res.bits_embedded.cpu().numpy() != want_nbits).sum()),
            "sse": int((res.sse.cpu().numpy().astype(np.int64) != want_sse).sum()),
            "bits": int(np.unpackbits(got_bits.cpu().numpy()[:, :nb] ^ want_bits[:, :nb], axis=1)[:, :cap].sum()) if cap else 0,
        }
        if any(diffs.values()):
            bad.append({"h": h, "w": w, "frames": f, "ch": ch, "kind": kind, "delta": delta, "num_ac": n,
                        "total_bits": total, "bit_offset": off, **diffs})
            if len(bad) >= 5:
                break
    print(json.dumps({"fuzz_parity": "cuda path vs oracle/dctqim_oracle.c", "seed": a.seed, "seconds": a.seconds,
                      "cases": cases, "frames": frames_total, "content": by_kind, "mismatching_cases": len(bad),
                      "first_mismatches": bad, "dry": a.dry}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
