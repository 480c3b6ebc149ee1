#!/usr/bin/env python
"""Markdown tables from a profiles/sweep.py run:  python profiles/sweep_table.py profiles/r2_sweep.jsonl"""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
c5 = [r for r in rows if r["config"] == 5]
deltas = sorted({r["delta"] for r in c5})
acs = sorted({r["num_ac"] for r in c5})
print("BASELINE config 5: delta x AC sweep, 64 x 1080p frames uniform[64,192), one B200, device-timed.")
print("Every point: stego pixels and extracted bits of the checked frame identical to the C oracle: %s (%d points)."
      % ("yes" if all(r["stego_pixels_identical_to_oracle"] and r["extracted_bits_identical_to_oracle"] for r in c5) else "NO", len(c5)))
print()
print("PSNR(gray, stego) dB / reference's own wrong bits in the checked frame (its extract vs the payload):")
print()
print("| delta | " + " | ".join("AC=%d" % a for a in acs) + " |")
print("|---|" + "---|" * len(acs))
for d in deltas:
    cells = []
    for a in acs:
        r = next(x for x in c5 if x["delta"] == d and x["num_ac"] == a)
        cells.append("%.1f / %d" % (r["psnr_db"], r["reference_wrong_bits_in_checked_frames"]))
    print("| %g | " % d + " | ".join(cells) + " |")
print()
print("Throughput, k frames/s embed / extract (capacity bits per frame in the header):")
print()
print("| delta | " + " | ".join("AC=%d (%d)" % (a, next(x for x in c5 if x["num_ac"] == a)["capacity_bits_per_frame"]) for a in acs) + " |")
print("|---|" + "---|" * len(acs))
for d in deltas:
    cells = []
    for a in acs:
        r = next(x for x in c5 if x["delta"] == d and x["num_ac"] == a)
        cells.append("%.0f / %.0f" % (r["embed_frames_per_s"] / 1e3, r["extract_frames_per_s"] / 1e3))
    print("| %g | " % d + " | ".join(cells) + " |")
print()
print("BASELINE config 4 on one GPU (75 frames of 3840x2160 = one GPU's share of the 600-frame batch on 8 GPUs):")
print()
for r in rows:
    if r["config"] == 4:
        print("* %d AC: embed %.3f ms, extract %.3f ms for %d frames = %.0f k / %.0f k frames/s, %.0f Gpixel/s round trip; identical to the oracle: %s"
              % (r["num_ac"], r["embed_ms"], r["extract_ms"], r["frames"], r["embed_frames_per_s"] / 1e3, r["extract_frames_per_s"] / 1e3,
                 r["round_trip_mpixel_per_s"] / 1e3, r["stego_pixels_identical_to_oracle"] and r["extracted_bits_identical_to_oracle"]))
