#!/usr/bin/env python
"""Latency of the reference-shaped single-frame call svs_b200.proses_frame_qim_dct (numpy frame and
'0'/'1' string in, numpy / string out - host<->device copies and string conversion included) next to
the unmodified reference function on the same host (the staged copy under oracle/_ref; the loop
port only when nothing is staged) - configs[0] and configs[1] shapes.

Usage: python profiles/dropin_latency.py > profiles/r2_dropin_latency.txt      (GPU box)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import svs_b200                                            # noqa: E402
from oracle import ref_port, stage_ref                     # noqa: E402  (CPU baseline only)

if stage_ref.available():
    cpu_fn, cpu_name = stage_ref.import_reference()["config_and_setup"].proses_frame_qim_dct, "the unmodified reference"
else:
    cpu_fn, cpu_name = ref_port.proses_frame_qim_dct, "CPU port of the reference"
from tests.synth import synth_frames, synth_bits, bits_to_str   # noqa: E402


def best(fn, reps):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return min(t) * 1e3, float(np.median(t)) * 1e3


for (h, w, n, payload_bits, cpu) in ((480, 640, 10, 9168, True), (480, 640, 10, 33744, True), (480, 640, 63, 302400, False),
                                     (720, 1280, 10, 144000, False), (1080, 1920, 63, 2041200, False)):
    frame = synth_frames("lat", (h, w, 3), 64, 192)
    seg = bits_to_str(synth_bits("lat", payload_bits))
    g, s, k = svs_b200.proses_frame_qim_dct(frame, 'embed', 20, seg, num_ac_coeffs_to_use=n)
    e_min, e_med = best(lambda: svs_b200.proses_frame_qim_dct(frame, 'embed', 20, seg, num_ac_coeffs_to_use=n), 20)
    stego3 = np.repeat(s[..., None], 3, 2)
    x_min, x_med = best(lambda: svs_b200.proses_frame_qim_dct(stego3, 'extract', 20, num_ac_coeffs_to_use=n), 20)
    line = "%4dx%-4d AC %2d payload %7d bits: GPU drop-in embed %.2f ms (median %.2f), extract %.2f ms (median %.2f)" % (
        w, h, n, payload_bits, e_min, e_med, x_min, x_med)
    if cpu:
        t0 = time.perf_counter()
        g2, s2, k2 = cpu_fn(frame, 'embed', 20, seg, num_ac_coeffs_to_use=n)
        t1 = time.perf_counter()
        out2 = cpu_fn(stego3, 'extract', 20, num_ac_coeffs_to_use=n)
        t2 = time.perf_counter()
        assert np.array_equal(s, s2) and k == k2
        assert out2 == svs_b200.proses_frame_qim_dct(stego3, 'extract', 20, num_ac_coeffs_to_use=n)
        line += "; %s (1 core) embed %%.0f ms, extract %%.0f ms -> x%%.0f / x%%.0f" % cpu_name % (
            (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t1 - t0) * 1e3 / e_min, (t2 - t1) * 1e3 / x_min)
    print(line, flush=True)
