"""Loader for the fixtures written by tests/golden/make_golden.py (outputs of the real reference)."""
import hashlib
import json
import os

import numpy as np

from tests.synth import synth_frames, synth_bits

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

with open(os.path.join(GOLDEN, "cases.json")) as _f:
    _META = json.load(_f)
ENV = _META["env"]
CASES = _META["cases"]
_small = None

# inputs of the digest-only cases are re-synthesised (tests/synth.py is deterministic)
_BIG_INPUTS = {
    "bgr_480x640_d20_ac10_cfg2": lambda: synth_frames("cfg2", (480, 640, 3)),
    "bgr_480x640_d20_ac63": lambda: synth_frames("cfg2b", (480, 640, 3)),
    "bgr_1080p_d20_ac63_fullrange": lambda: synth_frames("cfg3", (1080, 1920, 3)),
    "bgr_1080p_d20_ac10_midrange": lambda: synth_frames("cfg3m", (1080, 1920, 3), 64, 192),
    "bgr_4k_d20_ac63_midrange": lambda: synth_frames("cfg4", (2160, 3840, 3), 64, 192),
    "bgr_4k_d20_ac10_fullrange": lambda: synth_frames("cfg4f", (2160, 3840, 3)),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def small():
    global _small
    if _small is None:
        _small = np.load(os.path.join(GOLDEN, "cases_small.npz"))
    return _small


def case_ids(full=None, max_pixels=None):
    out = []
    for c in CASES:
        if full is not None and c["full"] != full:
            continue
        if max_pixels is not None and c["shape"][0] * c["shape"][1] > max_pixels:
            continue
        out.append(c["name"])
    return out


def get_case(name):
    c = next(c for c in CASES if c["name"] == name)
    if c["full"]:
        frame = small()[name + "/frame"]
    else:
        frame = _BIG_INPUTS[name]()
    assert sha(frame) == c["sha_frame"], "synthetic input drifted for %s" % name
    bits = synth_bits(name, c["nbits"])
    return c, frame, bits


def check_embed(c, gray, stego, n_emb):
    assert int(n_emb) == c["bits_embedded"], (c["name"], int(n_emb), c["bits_embedded"])
    assert sha(gray) == c["sha_gray"], "gray differs from the reference for %s" % c["name"]
    if c["full"]:
        ref = small()[c["name"] + "/stego"]
        diff = np.abs(ref.astype(np.int16) - np.asarray(stego).astype(np.int16))
        assert diff.max() == 0, "%s: %d stego pixels differ (max %d)" % (c["name"], int((diff > 0).sum()), int(diff.max()))
    assert sha(stego) == c["sha_stego"], "stego differs from the reference for %s" % c["name"]


def check_extract(c, bits01, which):
    bits01 = np.asarray(bits01, dtype=np.uint8)
    assert bits01.size == c["n_extracted"], (c["name"], bits01.size, c["n_extracted"])
    if c["full"]:
        ref = np.unpackbits(small()[c["name"] + "/" + which])[:c["n_extracted"]]
        assert np.array_equal(ref, bits01), "%s: %d extracted bits differ" % (c["name"], int((ref != bits01).sum()))
    assert sha(bits01) == c["sha_" + which], "extracted bits differ from the reference for %s" % c["name"]


def golden_stego(c):
    """Stego frame of a full case (for extract-from-stego checks); None for digest-only cases."""
    return small()[c["name"] + "/stego"] if c["full"] else None


def load_e2e():
    z = np.load(os.path.join(GOLDEN, "e2e_payload.npz"))
    with open(os.path.join(GOLDEN, "e2e_receiver_private.pem"), "rb") as f:
        pem = f.read()
    return dict(payload_packed=z["payload_packed"], total_bits=int(z["total_bits"]), image=z["image"],
                sha3=z["sha3"].tobytes(), width=int(z["width"]), height=int(z["height"]), pem=pem)
