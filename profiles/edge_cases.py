#!/usr/bin/env python
"""Edge-case launches of the block and the scalar kernels, compared with each other.

Written for `compute-sanitizer --tool memcheck` (with PYTORCH_NO_CUDA_MEMORY_CACHING=1 every
tensor is its own cudaMalloc, so a read or write one byte past a frame batch, a payload, a
stego buffer or a bit row would be reported) - but compute-sanitizer is closed on this GPU pool,
so the committed run (profiles/r2_edge_cases.txt) is the plain one: 528 embed + extract
launches whose stego, gray, bits_embedded, SSE and extracted bits must agree between the block
kernels and the scalar kernels.  Out-of-bounds WRITES are covered by the guard-band test
tests/test_gpu_parity.py::test_outputs_stay_inside_their_buffers.

The cases are the ones where the block kernels compute addresses differently from the common
path: ragged last groups (blocks per frame not a multiple of 32), a single 8x8 block, payloads
that end exactly at the last byte, a bit offset, the frame in which the payload ends (scalar
kernel for the tail), strided inputs, BGR stego output, the fused gray / SSE outputs, few
coefficients.

    python profiles/edge_cases.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svs_b200  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    L = svs_b200.lib()
    g = torch.Generator(device=dev).manual_seed(7)
    cases = 0
    for (f, h, w) in [(1, 8, 8), (3, 40, 72), (2, 64, 256), (5, 24, 1048), (2, 136, 264)]:
        for ch in (3, 1):
            shape = (f, h, w, 3) if ch == 3 else (f, h, w)
            frames = torch.randint(0, 256, shape, dtype=torch.uint8, device=dev, generator=g)
            # a strided view: rows and frames further apart than they need to be
            wide = torch.zeros((f, h + 8, w + 16) + ((3,) if ch == 3 else ()), dtype=torch.uint8, device=dev)
            view = wide[:, :h, 8:8 + w]
            view.copy_(frames)
            for n in ((1, 10, 63) if h == 40 else (10, 63)):
                cap = svs_b200.capacity_bits(h, w, n)
                for total, off in ((f * cap, 0), (f * cap - cap // 2, 0), (f * cap, 5)):
                    if total <= 0:
                        continue
                    nbytes = (off + total + 7) // 8                       # exactly enough, no padding
                    payload = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device=dev, generator=g)
                    outs = []
                    for fam in (0, 1):                                    # block kernels, scalar kernels
                        prev = L.svs_debug_kernel_family(fam)
                        for src in (frames, view):
                            for sc in (1, 3):
                                r = svs_b200.embed_frames(src, payload, total, 20.0, n, bit_offset=off, stego_channels=sc,
                                                          want_gray=True, want_bits_embedded=True, want_sse=True)
                                bits = svs_b200.extract_frames(r.stego, 20.0, n)
                                outs.append((fam, sc, r.stego.clone(), r.gray.clone(), r.bits_embedded.clone(),
                                             r.sse.clone(), bits.clone()))
                                cases += 1
                        L.svs_debug_kernel_family(prev)
                    torch.cuda.synchronize()
                    ref = {}
                    for fam, sc, *vals in outs:
                        key = sc
                        if key not in ref:
                            ref[key] = vals
                        else:
                            for a, b in zip(ref[key], vals):
                                if not torch.equal(a, b):
                                    raise SystemExit("MISMATCH f=%d %dx%d ch=%d n=%d total=%d off=%d fam=%d sc=%d"
                                                     % (f, h, w, ch, n, total, off, fam, sc))
    torch.cuda.synchronize()
    print("edge cases: %d embed+extract launches, block and scalar kernels agree" % cases)


if __name__ == "__main__":
    main()
