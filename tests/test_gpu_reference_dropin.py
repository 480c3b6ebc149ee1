"""The UNMODIFIED reference pipeline running on the B200 frame path.

oracle/stage_ref.py stages the reference's own modules (byte-identical, SHA-256 checked) into the
git-ignored oracle/_ref/, which travels to the GPU box.  Here `svs_b200.install()` rebinds the one
name `proses_frame_qim_dct` in those modules and then the reference's OWN functions run end to
end: embed_gambar_ke_video_final (embed_process.py:17-152) -> FFV1 file ->
ekstraksi_gambar_video_final (extract_process.py:22-216): ECDH + HKDF + AES-GCM decrypts, the
SHA3-256 digest matches and the recovered picture is pixel-identical.  The same run with the
reference's original CPU function must produce the same stego frames for the same payload."""
import os

import numpy as np
import pytest

import svs_b200
from oracle import stage_ref

pytestmark = pytest.mark.gpu

HAVE_REF = stage_ref.available()


class CountingDropIn:
    def __init__(self):
        self.calls = {"embed": 0, "extract": 0}

    def __call__(self, frame, mode, delta, *a, **kw):
        self.calls[mode] = self.calls.get(mode, 0) + 1
        return svs_b200.proses_frame_qim_dct(frame, mode, delta, *a, **kw)


def _frames_of(cv2, path, limit=None):
    cap = cv2.VideoCapture(path)
    out = []
    while limit is None or len(out) < limit:
        ok, f = cap.read()
        if not ok:
            break
        out.append(f)
    cap.release()
    return out


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref is not staged (python oracle/stage_ref.py where /root/reference exists)")
@pytest.mark.parametrize("cover,num_ac,delta,expect_ok", [
    ("clip", 10, 20, True),          # the reference's own defaults on its own clip (app.py:68-69)
    ("synthetic", 63, 20, True),     # maximum capacity on mid-range frames (its round trip is error-free there)
    ("clip", 63, 20, None),          # real frames clip at 63 AC: whatever the reference's verdict is, ours is the same
])
def test_reference_pipeline_runs_on_the_gpu_path(tmp_path, monkeypatch, capsys, cover, num_ac, delta, expect_ok):
    import cv2
    from PIL import Image
    from tests.synth import synth_frames
    ref = stage_ref.import_reference()
    cs, embed_process, extract_process = ref["config_and_setup"], ref["embed_process"], ref["extract_process"]
    original = cs.proses_frame_qim_dct
    monkeypatch.chdir(tmp_path)
    if cover == "clip":              # a short lossless cover from the reference's own clip (6 frames, 720p)
        clip = _frames_of(cv2, stage_ref.path("media/input/cover_1.mp4"), 6)
    else:
        clip = list(synth_frames("refcover", (5, 256, 320, 3), 64, 192))
    assert len(clip) >= 5
    h, w = clip[0].shape[:2]
    wr = cv2.VideoWriter("cover.avi", cv2.VideoWriter_fourcc(*"FFV1"), 24.0, (w, h), isColor=True)
    assert wr.isOpened()
    for f in clip:
        wr.write(f)
    wr.release()
    secret = stage_ref.path("media/input/image64.png")
    priv, pub = cs.buat_pasangan_kunci_ecc()
    pub_bytes = cs.serialisasi_kunci_publik_ecc_compressed(pub)

    drop = CountingDropIn()
    launches0 = svs_b200.lib().svs_kernel_launch_count()
    try:
        for m in (cs, embed_process, extract_process):
            m.proses_frame_qim_dct = drop                       # what svs_b200.install() does, with a call counter
        ok, gray0, stego0 = embed_process.embed_gambar_ke_video_final("cover.avi", secret, "stego_gpu.mp4", delta, num_ac, pub_bytes)
        assert ok is True and gray0 is not None and stego0 is not None
        gpu_ok = extract_process.ekstraksi_gambar_video_final("stego_gpu.avi", "out_gpu.png", delta, num_ac, priv)
    finally:
        for m in (cs, embed_process, extract_process):
            m.proses_frame_qim_dct = original
    assert drop.calls["embed"] >= 1 and drop.calls["extract"] >= 1
    assert svs_b200.lib().svs_kernel_launch_count() > launches0, "the CUDA library was not on the path"
    # the reference's own CPU function reads the GPU-made file: same verdict (decrypt + SHA3), same picture
    cpu_ok = extract_process.ekstraksi_gambar_video_final("stego_gpu.avi", "out_cpu.png", delta, num_ac, priv)
    assert gpu_ok == cpu_ok
    if expect_ok is not None:
        assert gpu_ok is expect_ok
    if gpu_ok:
        want = np.array(Image.open(secret).convert("L"))
        assert np.array_equal(np.array(Image.open("out_gpu.png")), want), "recovered picture differs"
        assert np.array_equal(np.array(Image.open("out_cpu.png")), want)
    # frame by frame the GPU drop-in returns what the reference function returns: the bits of the
    # first and last frame of the file, and a fresh embed of the first cover frame
    stego_frames = _frames_of(cv2, "stego_gpu.avi")
    assert len(stego_frames) == len(clip)
    hh, ww = (h // 8) * 8, (w // 8) * 8
    bits_cpu = None
    for i in (0, len(clip) - 1):
        bits_cpu = original(stego_frames[i][:hh, :ww], 'extract', delta, num_ac_coeffs_to_use=num_ac)
        bits_gpu = svs_b200.proses_frame_qim_dct(stego_frames[i][:hh, :ww], 'extract', delta, num_ac_coeffs_to_use=num_ac)
        assert bits_cpu == bits_gpu
    g_cpu, s_cpu, n_cpu = original(clip[0][:hh, :ww], 'embed', delta, bits_cpu[:5000], num_ac_coeffs_to_use=num_ac)
    g_gpu, s_gpu, n_gpu = svs_b200.proses_frame_qim_dct(clip[0][:hh, :ww], 'embed', delta, bits_cpu[:5000], num_ac_coeffs_to_use=num_ac)
    assert n_cpu == n_gpu and np.array_equal(g_cpu, g_gpu) and np.array_equal(s_cpu, s_gpu)
    capsys.readouterr()


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref is not staged")
def test_install_rebinds_the_three_call_sites():
    ref = stage_ref.import_reference()
    mods = [ref["config_and_setup"], ref["embed_process"], ref["extract_process"]]
    original = mods[0].proses_frame_qim_dct
    try:
        assert sorted(svs_b200.install(*mods)) == ["config_and_setup", "embed_process", "extract_process"]
        assert all(m.proses_frame_qim_dct is svs_b200.proses_frame_qim_dct for m in mods)
    finally:
        for m in mods:
            m.proses_frame_qim_dct = original
