#!/usr/bin/env python
"""Wall time of the reference's two pipeline functions on its own clip (media/input/cover_1.mp4,
1280x720 x 192 frames) - SURVEY.md 8f row N1:

  reference      the unmodified functions, CPU hot function (oracle/_ref, staged copy)
  drop-in        the same functions after svs_b200.install(): per-frame GPU calls
  batched        svs_b200.pipeline.install_pipelines(): one launch per batch of decoded frames,
                 decode / encode on worker threads, pinned staging, packed bit reader

for two payloads: the reference's default (image64.png, 10 AC: the payload fits in one frame, so
video I/O dominates everything) and a 256x256 secret at 1 AC (37 frames carry payload).  The
sender's randomness (ephemeral key, salt, nonce) is pinned for the duration of the measurement,
so all three modes embed the SAME payload: their stego videos must be identical frame for frame
and their verdicts (decrypt + SHA3 + pixels; real frames clip, so the reference itself may fail
the AES-GCM tag) must agree.

    python profiles/pipeline_walltime.py > profiles/r2_pipeline_walltime.txt      (GPU box)
"""
import contextlib
import hashlib
import io
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import svs_b200                                    # noqa: E402
from svs_b200 import pipeline                      # noqa: E402
from oracle import stage_ref                       # noqa: E402  (measurement infrastructure)


def main():
    from PIL import Image
    ref = stage_ref.import_reference()
    cs, ep, xp = ref["config_and_setup"], ref["embed_process"], ref["extract_process"]
    orig = {"fn": cs.proses_frame_qim_dct, "embed": ep.embed_gambar_ke_video_final, "extract": xp.ekstraksi_gambar_video_final}
    orig_keygen = cs.buat_pasangan_kunci_ecc
    clip = stage_ref.path("media/input/cover_1.mp4")
    work = tempfile.mkdtemp(prefix="svs_pipe_")
    big = os.path.join(work, "secret256.png")
    rng = np.random.default_rng(5)
    Image.fromarray(rng.integers(0, 256, (256, 256), dtype=np.uint8), "L").save(big)
    priv, pub = cs.buat_pasangan_kunci_ecc()
    pub_bytes = cs.serialisasi_kunci_publik_ecc_compressed(pub)
    skip_cpu_big = os.environ.get("SVS_PIPE_SKIP_CPU_BIG") == "1"

    def restore():
        for m in (cs, ep, xp):
            m.proses_frame_qim_dct = orig["fn"]
        ep.embed_gambar_ke_video_final, xp.ekstraksi_gambar_video_final = orig["embed"], orig["extract"]

    eph = cs.buat_pasangan_kunci_ecc()
    real_urandom = os.urandom

    def pin_randomness():
        state = {"n": 0}

        def fake(k):
            state["n"] += 1
            return hashlib.sha256(b"svs-pipeline-walltime-%d" % state["n"]).digest()[:k] if k <= 32 else real_urandom(k)

        os.urandom = fake
        for m in (cs, ep):
            m.buat_pasangan_kunci_ecc = lambda: eph

    def unpin_randomness():
        os.urandom = real_urandom
        cs.buat_pasangan_kunci_ecc = orig_keygen
        ep.buat_pasangan_kunci_ecc = orig_keygen

    def video_digest(path):
        import cv2
        cap, h, n = cv2.VideoCapture(path), hashlib.sha256(), 0
        while True:
            ok, f = cap.read()
            if not ok:
                break
            h.update(f.tobytes())
            n += 1
        cap.release()
        return "%d frames %s" % (n, h.hexdigest()[:16])

    digests = {}

    def run(mode, secret, num_ac, tag):
        restore()
        if mode == "drop-in":
            svs_b200.install(cs, ep, xp)
        elif mode == "batched":
            pipeline.install_pipelines([ep, xp])
        base = os.path.join(work, "%s_%s" % (mode.replace("-", ""), "".join(ch if ch.isalnum() else "_" for ch in tag)))
        sink = io.StringIO()
        pin_randomness()
        try:
            with contextlib.redirect_stdout(sink):
                t0 = time.perf_counter()
                ok, _, _ = ep.embed_gambar_ke_video_final(clip, secret, base + ".mp4", 20, num_ac, pub_bytes)
                t1 = time.perf_counter()
                ok2 = xp.ekstraksi_gambar_video_final(base + ".avi", base + ".png", 20, num_ac, priv)
                t2 = time.perf_counter()
        finally:
            unpin_randomness()
        same = bool(ok and ok2 and np.array_equal(np.array(Image.open(base + ".png")), np.array(Image.open(secret).convert("L"))))
        verdict = "ok" if same else ("AES-GCM tag / SHA3 rejected" if ok and not ok2 else "embed failed")
        dig = video_digest(base + ".avi") if ok else "-"
        digests.setdefault(tag, {})[mode] = (dig, verdict)
        print("%-10s %-24s embed %7.2f s   extract %6.2f s   verdict: %-28s stego video: %s" % (
            mode, tag, t1 - t0, t2 - t1, verdict, dig), flush=True)
        try:
            os.remove(base + ".avi")
        except OSError:
            pass

    print("cover_1.mp4 1280x720 x 192 frames, delta 20; FFV1 output; wall clock incl. decode, encode, crypto")
    for secret, num_ac, tag in ((stage_ref.path("media/input/image64.png"), 10, "image64/10AC(1 frame)"),
                                (big, 1, "256x256/1AC(37 frames)")):
        for mode in ("reference", "drop-in", "batched"):
            if mode == "reference" and secret == big and skip_cpu_big:
                continue
            run(mode, secret, num_ac, tag)
    restore()
    for tag, by_mode in digests.items():
        vals = set(by_mode.values())
        print("%-24s all modes wrote the same stego video and reached the same verdict: %s" % (tag, "yes" if len(vals) == 1 else "NO %r" % (by_mode,)))
    if any(len(set(v.values())) != 1 for v in digests.values()):
        raise SystemExit(1)


if __name__ == "__main__":
    main()
