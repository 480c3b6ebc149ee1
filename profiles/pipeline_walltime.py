#!/usr/bin/env python
"""Wall time of the reference's two pipeline functions on its own clip (media/input/cover_1.mp4,
1280x720 x 192 frames) - SURVEY.md 8f row N1:

  reference      the unmodified functions, CPU hot function (oracle/_ref, staged copy)
  drop-in        the same functions after svs_b200.install(): per-frame GPU calls
  batched        svs_b200.pipeline.install_pipelines(): one launch per batch of decoded frames,
                 decode / encode on worker threads, pinned staging, packed bit reader

for two payloads: the reference's default (image64.png, 10 AC: the payload fits in one frame, so
video I/O dominates everything) and a 256x256 secret at 1 AC (37 frames carry payload).  Every run
must decrypt, verify SHA3 and return the identical picture.

    python profiles/pipeline_walltime.py > profiles/r2_pipeline_walltime.txt      (GPU box)
"""
import contextlib
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

    def run(mode, secret, num_ac, tag):
        restore()
        if mode == "drop-in":
            svs_b200.install(cs, ep, xp)
        elif mode == "batched":
            pipeline.install_pipelines([ep, xp])
        base = os.path.join(work, "%s_%s" % (mode.replace("-", ""), tag))
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink):
            t0 = time.perf_counter()
            ok, _, _ = ep.embed_gambar_ke_video_final(clip, secret, base + ".mp4", 20, num_ac, pub_bytes)
            t1 = time.perf_counter()
            ok2 = xp.ekstraksi_gambar_video_final(base + ".avi", base + ".png", 20, num_ac, priv)
            t2 = time.perf_counter()
        same = bool(ok and ok2 and np.array_equal(np.array(Image.open(base + ".png")), np.array(Image.open(secret).convert("L"))))
        print("%-10s %-22s embed %7.2f s   extract %6.2f s   decrypt+SHA3+pixels %s" % (
            mode, tag, t1 - t0, t2 - t1, "ok" if same else "FAILED (embed %s, extract %s)" % (ok, ok2)), flush=True)
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


if __name__ == "__main__":
    main()
