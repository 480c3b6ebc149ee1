"""Host logic of the batched pipeline driver (N1): frame routing, payload slicing, read-as-needed
extraction and the payload byte layout - on CPU, with the C oracle standing in for the kernels.

The last two tests run ONLY where /root/reference is mounted (the build container): they drive
the reference's own, unmodified pipeline functions against the batched ones in both directions
through a real FFV1 file, which pins the payload layout and the frame order to the reference."""
import os
import sys

import numpy as np
import pytest

import svs_b200
from svs_b200 import pipeline
from oracle import c_oracle as oc
from oracle import dctqim_oracle as onp
from tests.synth import synth_frames, synth_bits

REF = "/root/reference"


def oracle_embed(frames, packed, bit_offset, nbits, delta, num_ac, want_gray):
    stego, gray, nb = oc.embed_frames(frames, packed, nbits, delta, num_ac, bit_offset=bit_offset)
    return np.repeat(stego[..., None], 3, 3), (gray[0] if want_gray else None), nb


def oracle_extract(frames, delta, num_ac):
    return oc.extract_frames(frames, delta, num_ac)


class Feed:
    """cv2.VideoCapture.read() over a list of frames."""

    def __init__(self, frames):
        self.frames, self.i = list(frames), 0

    def read(self):
        if self.i >= len(self.frames):
            return False, None
        self.i += 1
        return True, self.frames[self.i - 1]


def reference_loop(frames, bits, delta, num_ac, h, w):
    """The frame loop of embed_process.py:108-144, frame by frame with the NumPy oracle."""
    out, idx, done, first = [], 0, False, (None, None)
    for k, f in enumerate(frames):
        f = f[0:h, 0:w]
        if idx < len(bits) and not done:
            gray, stego, n = onp.embed_frame(np.ascontiguousarray(f), delta, bits[idx:], num_ac)
            if k == 0:
                first = (gray, stego)
            out.append(np.repeat(stego[..., None], 3, 2))
            idx += n
            done = idx >= len(bits)
        else:
            out.append(f)
    return out, done, first


@pytest.mark.parametrize("n_frames,frac,batch", [(7, 2.4, 3), (5, 5.0, 2), (4, 0.3, 8), (3, 3.5, 2)])
def test_embed_stream_equals_the_frame_loop(n_frames, frac, batch):
    h, w, n, delta = 40, 56, 17, 12
    frames = synth_frames("pipe%d" % n_frames, (n_frames, 43, 59, 3), 40, 220)      # cropped to 40x56 like :113
    cap = svs_b200.capacity_bits(h, w, n)
    bits = synth_bits("pipe", int(frac * cap) + 5)
    written = []
    ok, g0, s0, seen = pipeline.embed_frame_stream(Feed(frames).read, lambda a: written.append(np.array(a)),
                                                   np.packbits(bits), bits.size, delta, n, (h, w),
                                                   batch_frames=batch, embed_fn=oracle_embed)
    want, done, first = reference_loop(frames, bits, delta, n, h, w)
    assert ok == done and seen == n_frames and len(written) == n_frames
    for a, b in zip(written, want):
        assert a.shape == (h, w, 3) and np.array_equal(a, b)
    if done:
        assert np.array_equal(g0, first[0]) and np.array_equal(s0, first[1])
    else:
        assert g0 is None and s0 is None


def test_embed_stream_degenerate_cases():
    h, w = 16, 24
    frames = synth_frames("deg", (3, h, w, 3))
    bits = synth_bits("deg", 100)
    for delta, n, total in ((20, 10, 0), (0, 10, 100), (20, 0, 100)):
        written = []
        ok, g0, s0, seen = pipeline.embed_frame_stream(Feed(frames).read, written.append, np.packbits(bits), total,
                                                       delta, n, (h, w), batch_frames=2, embed_fn=oracle_embed)
        assert not ok and g0 is None and seen == 3 and len(written) == 3     # the payload never completes
        if total == 0:
            assert all(np.array_equal(a, b) for a, b in zip(written, frames))   # copied in colour
        else:                                                                # every frame still takes the hot path
            for a, f in zip(written, frames):
                _, stego, k = onp.embed_frame(f, delta, bits, n)
                assert k == 0 and np.array_equal(a[..., 0], stego)


def test_bit_reader_reads_only_what_it_needs():
    h, w, n, delta = 32, 48, 10, 20
    frames = synth_frames("reader", (9, h, w, 3))
    cap = svs_b200.capacity_bits(h, w, n)                        # 240 bits = 30 bytes per frame
    feed = Feed(frames)
    rd = pipeline.StegoBitReader(feed.read, delta, n, (h, w), batch_frames=4, extract_fn=oracle_extract)
    stream = np.concatenate([np.unpackbits(r)[:cap] for r in oc.extract_frames(frames, delta, n)])
    assert rd.take_bytes(2) == np.packbits(stream[:16]).tobytes() and feed.i == 1        # one frame was enough
    assert rd.take_uint(1) == int(np.packbits(stream[16:24])[0]) and feed.i == 1
    got = rd.take_bytes(100)                                     # 800 bits more: 4 frames -> one batch of 3 + ...
    assert got == np.packbits(stream[24:824]).tobytes() and feed.i == 4
    rest = (9 * cap - 824) // 8
    assert rd.take_bytes(rest) == np.packbits(stream[824:824 + 8 * rest]).tobytes() and feed.i == 9
    with pytest.raises(EOFError):
        rd.take_bytes(2)


def test_bit_reader_with_frames_that_do_not_end_on_a_byte():
    """cap % 8 != 0: rows are merged at bit granularity, the cursor may sit inside a byte."""
    h, w, n, delta = 24, 40, 7, 20                                   # 15 blocks x 7 = 105 bits per frame
    frames = synth_frames("reader2", (11, h, w))
    cap = svs_b200.capacity_bits(h, w, n)
    assert cap % 8 == 1
    stream = np.concatenate([np.unpackbits(r)[:cap] for r in oc.extract_frames(frames, delta, n)])
    feed = Feed(frames)
    rd = pipeline.StegoBitReader(feed.read, delta, n, (h, w), batch_frames=3, extract_fn=oracle_extract)
    pos = 0
    for nbytes in (1, 13, 2, 40, 5, 60):
        assert rd.take_bytes(nbytes) == np.packbits(stream[pos:pos + 8 * nbytes]).tobytes()
        pos += 8 * nbytes
    assert rd.available() == rd.frames_read * cap - pos and rd.frames_read <= 11


def test_embed_stream_with_and_without_io_threads_write_the_same_frames():
    h, w, n, delta = 40, 56, 17, 12
    frames = synth_frames("pipe-threads", (9, 43, 59, 3), 40, 220)
    cap = svs_b200.capacity_bits(h, w, n)
    bits = synth_bits("pipe-threads", int(3.3 * cap))
    outs = []
    for overlap in (False, True):
        written = []
        res = pipeline.embed_frame_stream(Feed(frames).read, lambda a: written.append(np.array(a)), np.packbits(bits),
                                          bits.size, delta, n, (h, w), batch_frames=2, embed_fn=oracle_embed,
                                          overlap_io=overlap)
        outs.append((res[0], res[3], written))
    assert outs[0][0] is True and outs[0][:2] == outs[1][:2] and len(outs[0][2]) == 9
    assert all(np.array_equal(a, b) for a, b in zip(outs[0][2], outs[1][2]))


def test_payload_layout_round_trip_and_golden():
    """build_payload reproduces the payload the reference built for image64.png (tests/golden)."""
    from tests import golden_util as G
    e = G.load_e2e()
    raw = e["payload_packed"].tobytes()[:e["total_bits"] // 8]

    class Bytes:
        def __init__(self, b):
            self.b, self.p = b, 0

        def take_bytes(self, n):
            self.p += n
            return self.b[self.p - n:self.p]

        def take_uint(self, n):
            return int.from_bytes(self.take_bytes(n), "big")

    fields = pipeline.parse_payload(Bytes(raw))
    assert fields[:2] == (e["width"], e["height"]) and fields[4] == e["sha3"]
    assert [len(f) for f in fields[2:7]] == [33, 16, 32, 12, 16] and len(fields[7]) == e["width"] * e["height"]
    assert pipeline.build_payload(*fields) == raw
    assert 8 * (len(raw) - len(fields[7])) == pipeline.HEADER_BITS


# ------------------------------------------------------------------ against the real reference
def _reference():
    if not os.path.isdir(REF):
        pytest.skip("reference not mounted")
    cv2 = pytest.importorskip("cv2")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import config_and_setup as cs
    import helpers
    import embed_process
    import extract_process
    import types
    return cv2, types.SimpleNamespace(cs=cs, helpers=helpers), embed_process, extract_process


def _write_cover(cv2, path, frames):
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), 24.0, (frames.shape[2], frames.shape[1]), isColor=True)
    assert wr.isOpened()
    for f in frames:
        wr.write(f)
    wr.release()


@pytest.mark.parametrize("num_ac", [20, 63])
def test_batched_embed_is_read_back_by_the_reference_extract(tmp_path, monkeypatch, capsys, num_ac):
    cv2, ref, _, extract_process = _reference()
    monkeypatch.chdir(tmp_path)
    frames = synth_frames("cover", (6, 128, 136, 3), 64, 192)
    _write_cover(cv2, "cover.avi", frames)
    priv, pub = ref.cs.buat_pasangan_kunci_ecc()
    secret = os.path.join(REF, "media", "input", "image32.png")
    ok, gray, stego = pipeline.embed_gambar_ke_video_final("cover.avi", secret, "stego_out.mp4", 20, num_ac,
                                                           ref.cs.serialisasi_kunci_publik_ecc_compressed(pub), ref=ref,
                                                           batch_frames=4, embed_fn=oracle_embed, verbose=False)
    assert ok and gray.shape == stego.shape == (128, 136) and os.path.exists("stego_out.avi")
    # the UNMODIFIED reference extracts, decrypts and verifies it
    assert extract_process.ekstraksi_gambar_video_final("stego_out.avi", "out.png", 20, num_ac, priv) is True
    from PIL import Image
    assert np.array_equal(np.array(Image.open("out.png")), np.array(Image.open(secret).convert("L")))
    # frame count and pass-through of the frames after the payload
    cap = cv2.VideoCapture("stego_out.avi")
    got = []
    while True:
        ret, f = cap.read()
        if not ret:
            break
        got.append(f)
    cap.release()
    used = -(-(976 + 8 * 32 * 32) // svs_b200.capacity_bits(128, 136, num_ac))
    assert len(got) == 6 and all(np.array_equal(got[i], frames[i]) for i in range(used, 6))
    capsys.readouterr()


def test_reference_embed_is_read_back_by_the_batched_extract(tmp_path, monkeypatch, capsys):
    cv2, ref, embed_process, _ = _reference()
    monkeypatch.chdir(tmp_path)
    frames = synth_frames("cover2", (5, 128, 128, 3), 64, 192)
    _write_cover(cv2, "cover.avi", frames)
    priv, pub = ref.cs.buat_pasangan_kunci_ecc()
    secret = os.path.join(REF, "media", "input", "image32.png")
    ok, _, _ = embed_process.embed_gambar_ke_video_final("cover.avi", secret, "stego_ref.mp4", 20, 20,
                                                         ref.cs.serialisasi_kunci_publik_ecc_compressed(pub))
    assert ok
    assert pipeline.ekstraksi_gambar_video_final("stego_ref.avi", "out.png", 20, 20, priv, ref=ref, batch_frames=2,
                                                 extract_fn=oracle_extract, verbose=False) is True
    from PIL import Image
    assert np.array_equal(np.array(Image.open("out.png")), np.array(Image.open(secret).convert("L")))
    # a wrong key must fail like the reference does (AES-GCM tag)
    other, _ = ref.cs.buat_pasangan_kunci_ecc()
    assert pipeline.ekstraksi_gambar_video_final("stego_ref.avi", "bad.png", 20, 20, other, ref=ref,
                                                 extract_fn=oracle_extract, verbose=False) is False
    capsys.readouterr()
