#!/usr/bin/env python
"""Print the key timings of the last JSON line in a bench log."""
import json
import sys
lines = [l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")]
d = json.loads(lines[-1])
print("%s value %.0f embed_ms %.3f frac %.3f extract_ms %.3f frac %.3f rt %.3f parity %s" % (
    sys.argv[2] if len(sys.argv) > 2 else "", d["value"], d["roofline"]["launch_ms"], d["roofline"]["frac"],
    d["roofline_extract"]["launch_ms"], d["roofline_extract"]["frac"], d["roofline_round_trip"]["frac"], d["parity_check"]))
