"""CPU tests: the oracles (NumPy, plain C, loop port) against the reference's golden vectors.

The fixtures in tests/golden/ are outputs of the unmodified reference function
(config_and_setup.py:106-174) - see tests/golden/make_golden.py.  When /root/reference is
present (the build container) the oracles are also compared with the live function.
"""
import os
import sys

import numpy as np
import pytest

from oracle import dctqim_oracle as onp
from oracle import c_oracle as oc
from tests import golden_util as G
from tests.synth import synth_frames, synth_bits, bits_to_str

HAVE_REF = os.path.exists("/root/reference/config_and_setup.py")


def _ref():
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    import config_and_setup
    return config_and_setup


# ---------------------------------------------------------------- transforms vs scipy bit patterns
def test_dct8_bit_patterns_match_scipy():
    fft = pytest.importorskip("scipy.fftpack")
    rng = np.random.default_rng(7)
    x = rng.integers(0, 256, (50000, 8)).astype(np.float32)
    y = (rng.standard_normal((50000, 8)) * 300).astype(np.float32)
    for a in (x, y):
        mine = np.stack(onp.dct8([a[:, i] for i in range(8)]), 1)
        assert np.array_equal(mine.view(np.uint32), fft.dct(a, axis=1, norm='ortho').view(np.uint32))
        mine = np.stack(onp.idct8([a[:, i] for i in range(8)]), 1)
        assert np.array_equal(mine.view(np.uint32), fft.idct(a, axis=1, norm='ortho').view(np.uint32))


def test_dct2_blocks_match_scipy_and_c():
    fft = pytest.importorskip("scipy.fftpack")
    rng = np.random.default_rng(8)
    b = rng.integers(0, 256, (2000, 8, 8)).astype(np.float32)
    want = fft.dct(fft.dct(b, axis=1, norm='ortho'), axis=2, norm='ortho')
    got = onp.dct2_blocks(b)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    back = fft.idct(fft.idct(want, axis=1, norm='ortho'), axis=2, norm='ortho')
    assert np.array_equal(onp.idct2_blocks(want).view(np.uint32), back.view(np.uint32))
    for i in range(0, 2000, 97):
        assert np.array_equal(oc.dct2(b[i]).view(np.uint32), want[i].view(np.uint32))
        assert np.array_equal(oc.idct2(want[i]).view(np.uint32), back[i].view(np.uint32))


def test_gray_formula_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    f = synth_frames("gray", (64, 96, 3))
    assert np.array_equal(onp.bgr_to_gray(f), cv2.cvtColor(f, cv2.COLOR_BGR2GRAY))


# ---------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("name", G.case_ids())
def test_numpy_oracle_against_golden(name):
    c, frame, bits = G.get_case(name)
    gray, stego, n = onp.embed_frame(frame, c["delta"], bits, c["num_ac"])
    G.check_embed(c, gray, stego, n)
    G.check_extract(c, onp.extract_frame_bits(frame, c["delta"], c["num_ac"]), "ext_input")
    G.check_extract(c, onp.extract_frame_bits(stego, c["delta"], c["num_ac"]), "ext_stego")


@pytest.mark.parametrize("name", G.case_ids())
def test_c_oracle_against_golden(name):
    c, frame, bits = G.get_case(name)
    packed = np.packbits(bits)
    stego, gray, n = oc.embed_frames(frame[None], packed, bits.size, c["delta"], c["num_ac"])
    G.check_embed(c, gray[0], stego[0], n[0])
    cap = c["n_extracted"]
    ext = oc.extract_frames(frame[None], c["delta"], c["num_ac"])
    G.check_extract(c, np.unpackbits(ext[0])[:cap], "ext_input")
    ext = oc.extract_frames(stego, c["delta"], c["num_ac"])
    G.check_extract(c, np.unpackbits(ext[0])[:cap], "ext_stego")


@pytest.mark.parametrize("name", G.case_ids(full=True, max_pixels=48 * 64))
def test_loop_port_against_golden(name):
    pytest.importorskip("scipy.fftpack")
    from oracle import ref_port
    c, frame, bits = G.get_case(name)
    gray, stego, n = ref_port.proses_frame_qim_dct(frame, 'embed', c["delta"], bits_to_str(bits),
                                                   num_ac_coeffs_to_use=c["num_ac"])
    G.check_embed(c, gray, stego, n)
    s = ref_port.proses_frame_qim_dct(stego, 'extract', c["delta"], num_ac_coeffs_to_use=c["num_ac"])
    G.check_extract(c, np.frombuffer(s.encode(), np.uint8) - 48, "ext_stego")


def test_loop_port_against_the_480p_digest_case():
    """The loop port at a full 640x480 frame (BASELINE config 2 shape): what bench.py times as the
    CPU baseline when no staged reference is present must itself be pinned at a realistic size."""
    pytest.importorskip("scipy.fftpack")
    from oracle import ref_port
    c, frame, bits = G.get_case("bgr_480x640_d20_ac10_cfg2")
    gray, stego, n = ref_port.proses_frame_qim_dct(frame, 'embed', c["delta"], bits_to_str(bits),
                                                   num_ac_coeffs_to_use=c["num_ac"])
    G.check_embed(c, gray, stego, n)
    s = ref_port.proses_frame_qim_dct(stego, 'extract', c["delta"], num_ac_coeffs_to_use=c["num_ac"])
    G.check_extract(c, np.frombuffer(s.encode(), np.uint8) - 48, "ext_stego")


def test_c_oracle_batch_offsets_and_threads():
    """Frame f consumes payload bits [f*cap, (f+1)*cap) (embed_process.py:115-128)."""
    frames = synth_frames("batch", (5, 32, 40, 3))
    n, delta = 63, 20
    cap = onp.capacity_bits(32, 40, n)
    total = 3 * cap + 500                          # ends inside frame 3; frame 4 untouched
    bits = synth_bits("batch", total + 13)
    off = 13
    packed = np.packbits(bits)
    for threads in (1, 3):
        stego, gray, nemb = oc.embed_frames(frames, packed, total, delta, n, bit_offset=off, threads=threads)
        for f in range(5):
            seg = bits[off + f * cap: off + total] if f * cap < total else bits[:0]
            g, s, k = onp.embed_frame(frames[f], delta, seg, n)
            assert np.array_equal(s, stego[f]) and np.array_equal(g, gray[f]) and k == nemb[f]
        assert np.array_equal(stego[4], gray[4])
        ext = oc.extract_frames(stego, delta, n, threads=threads)
        for f in range(5):
            assert np.array_equal(np.unpackbits(ext[f])[:cap], onp.extract_frame_bits(stego[f], delta, n))


def test_oracle_rejects_bad_shapes():
    with pytest.raises(ValueError, match="Format frame input tidak didukung"):
        onp.proses_frame_qim_dct(np.zeros((8, 8, 4), np.uint8), 'extract', 20)
    with pytest.raises(ValueError):
        onp.embed_frame(np.zeros((12, 8), np.uint8), 20, "1")


# ---------------------------------------------------------------- live reference (build container)
@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present")
@pytest.mark.parametrize("seed", range(6))
def test_oracles_against_live_reference(seed):
    ref = _ref()
    rng = np.random.default_rng(seed)
    h, w = 8 * int(rng.integers(1, 6)), 8 * int(rng.integers(1, 8))
    three = bool(rng.integers(0, 2))
    delta = [20, 7, 3, 1, 50, 2.5][seed]
    n = int(rng.integers(1, 70))
    lo, hi = [(0, 256), (64, 192), (0, 32), (224, 256), (0, 256), (100, 140)][seed]
    frame = synth_frames("live%d" % seed, (h, w, 3) if three else (h, w), lo, hi)
    cap = onp.capacity_bits(h, w, n)
    nbits = int(rng.integers(0, cap + 50))
    bits = synth_bits("live%d" % seed, nbits)
    g0, s0, k0 = ref.proses_frame_qim_dct(frame, 'embed', delta, bits_to_str(bits), num_ac_coeffs_to_use=n)
    e0 = ref.proses_frame_qim_dct(s0, 'extract', delta, num_ac_coeffs_to_use=n)
    g1, s1, k1 = onp.embed_frame(frame, delta, bits, n)
    assert np.array_equal(g0, g1) and np.array_equal(s0, s1) and k0 == k1
    assert e0 == onp.extract_frame(s0, delta, n)
    s2, g2, k2 = oc.embed_frames(frame[None], np.packbits(bits), nbits, delta, n)
    assert np.array_equal(g0, g2[0]) and np.array_equal(s0, s2[0]) and k0 == k2[0]
    e2 = oc.extract_frames(s0[None], delta, n)
    assert e0 == bits_to_str(np.unpackbits(e2[0])[:cap])
