"""CPU tests: C-ABI surface, host-side helpers, and the host build of the kernel arithmetic."""
import ctypes
import os
import re

import numpy as np
import pytest

import svs_b200
from oracle import dctqim_oracle as onp
from tests.host_math import build as hm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """Every function declared in include/svs_b200.h is exported by libsvs_b200.so."""
    hdr = open(os.path.join(ROOT, "include", "svs_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(svs_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = ctypes.CDLL(svs_b200.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), "missing export: %s" % name
    assert declared == set(svs_b200.EXPORTED_SYMBOLS)


def test_abi_scalar_entry_points_without_gpu():
    L = svs_b200.lib()
    assert L.svs_version() >= 100
    assert L.svs_capacity_bits(1080, 1920, 63) == 2041200
    assert L.svs_capacity_bits(1080, 1920, 100) == 2041200
    assert L.svs_capacity_bits(480, 640, 10) == 48000
    assert L.svs_capacity_bits(480, 640, -3) == 0
    assert L.svs_bits_row_bytes(1080, 1920, 63) % 16 == 0
    assert L.svs_bits_row_bytes(1080, 1920, 63) >= 255150
    assert svs_b200.capacity_bits(2160, 3840, 63) == 8164800


def test_argument_errors_are_reported_before_any_cuda_call():
    L = svs_b200.lib()
    rc = L.svs_extract_frames(None, 3, 1, 12, 16, 0, 0, 20.0, 10, None, 0, None)
    assert rc == -1 and "multiples of 8" in svs_b200._native.last_error()
    rc = L.svs_extract_frames(None, 2, 1, 16, 16, 0, 0, 20.0, 10, None, 0, None)
    assert rc == -1
    rc = L.svs_extract_frames(ctypes.c_void_p(256), 3, 1, 16, 16, 16 * 48, 48, float("nan"), 10, None, 0, None)
    assert rc == -4
    rc = L.svs_extract_frames(ctypes.c_void_p(256), 3, 1, 16, 16, 16 * 48, 48, 1e-5, 10, None, 0, None)
    assert rc == -4
    rc = L.svs_extract_frames(ctypes.c_void_p(256), 3, 1, 16, 16, 16 * 48, 40, 20.0, 10, None, 0, None)
    assert rc == -5


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises((svs_b200.SvsError, RuntimeError)):
        svs_b200.proses_frame_qim_dct(np.zeros((16, 16, 3), np.uint8), 'extract', 20)
    with pytest.raises(RuntimeError):
        svs_b200.extract_frames(torch.zeros((1, 16, 16), dtype=torch.uint8), 20, 10)


def test_frame_shape_errors_match_reference():
    with pytest.raises(ValueError, match="Format frame input tidak didukung"):
        svs_b200.proses_frame_qim_dct(np.zeros((16, 16, 4), np.uint8), 'extract', 20)
    with pytest.raises(ValueError, match="kelipatan 8"):
        svs_b200.proses_frame_qim_dct(np.zeros((12, 16, 3), np.uint8), 'extract', 20)


def test_bitstream_helpers():
    s = "1011000111010"
    b = svs_b200.bits_from_str(s)
    assert b.tolist() == [int(c) for c in s]
    assert svs_b200.bits_to_str(b) == s
    packed, n = svs_b200.pack_str(s)
    assert n == 13 and packed.tolist() == [0b10110001, 0b11010000]
    assert svs_b200.unpack_bits(packed, 13).tolist() == b.tolist()
    assert svs_b200.bytes_to_bitstring(b"\x80\x01") == "1000000000000001"
    assert svs_b200.bitstring_to_bytes("100000000000000111") == b"\x80\x01"
    with pytest.raises(ValueError):
        svs_b200.bitstring_to_bytes("101")
    with pytest.raises(ValueError, match="invalid literal"):
        svs_b200.bits_from_str("0102")
    assert svs_b200.bits_from_str("01x", limit=2).tolist() == [0, 1]


def test_install_rebinds_three_names():
    import types
    mods = [types.ModuleType(n) for n in ("config_and_setup", "embed_process", "extract_process")]
    assert svs_b200.install(*mods) == ["config_and_setup", "embed_process", "extract_process"]
    assert all(m.proses_frame_qim_dct is svs_b200.proses_frame_qim_dct for m in mods)


def test_kernel_arithmetic_host_build_matches_oracle():
    """csrc/svs_math.cuh (folded constants, the op order the kernels run) == oracle, bit for bit."""
    L = hm.load()
    rng = np.random.default_rng(11)
    ints = rng.integers(0, 256, (20000, 8, 8)).astype(np.float32)
    got = ints.copy()
    L.hm_dct2_fwd(got.ctypes.data, 20000)
    want = onp.dct2_blocks(ints)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    coefs = want.copy()
    coefs[:, 0, 1:] = np.round(coefs[:, 0, 1:] / 20) * 20
    got = coefs.copy()
    L.hm_dct2_inv(got.ctypes.data, 20000)
    assert np.array_equal(got.view(np.uint32), onp.idct2_blocks(coefs).view(np.uint32))
    wild = (rng.standard_normal((20000, 8)) * 500).astype(np.float32)
    a = wild.copy(); L.hm_dct8_fwd(a.ctypes.data, 20000)
    b = np.stack(onp.dct8([wild[:, i] for i in range(8)]), 1)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    a = wild.copy(); L.hm_dct8_inv(a.ctypes.data, 20000)
    b = np.stack(onp.idct8([wild[:, i] for i in range(8)]), 1)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("delta", [0.0625, 0.5, 1, 2, 2.5, 3, 4, 6, 7, 8, 10, 16, 20, 32, 50, 64, 100, 1000])
def test_quantiser_claims_against_the_ieee_division(delta):
    """csrc/svs_quant.h: the division-free quantiser only ever differs from the reference's
    round(c / delta) (config_and_setup.py:148,160) on coefficients it flags for the exact path, and
    the exact path (rounded reciprocal + two Newton steps + magic-constant rint) never differs -
    on random coefficients, on multiples of 1/8 (what (0,4),(4,0),(4,4) produce), on exact rounding
    ties and on their float32 neighbours."""
    L = hm.load()
    rng = np.random.default_rng(int(delta * 16))
    d32 = np.float32(delta)
    ties = ((np.arange(-3000, 3000) + 0.5) * float(d32)).astype(np.float32)
    ties = ties[np.abs(ties) <= 2040]
    near = np.concatenate([np.nextafter(ties, np.float32(np.inf)), np.nextafter(ties, np.float32(-np.inf)),
                           np.nextafter(np.nextafter(ties, np.float32(np.inf)), np.float32(np.inf))])
    c = np.concatenate([rng.uniform(-2040, 2040, 300000).astype(np.float32),
                        (rng.integers(-16320, 16321, 100000) / 8.0).astype(np.float32),
                        (rng.standard_normal(100000) * 40).astype(np.float32), ties, near,
                        np.array([0.0, -0.0, 2040.0, -2040.0], np.float32)])
    c = np.ascontiguousarray(c)
    stats = np.zeros(8, np.int64)
    L.hm_quant_check(float(delta), c.ctypes.data, c.size, stats.ctypes.data)
    flagged_e, flagged_x, bad_fast_e, bad_fast_x, bad_exact_e, bad_exact_x, embed_ok, extract_ok = stats.tolist()
    assert embed_ok and (extract_ok or delta < 0.25)   # these deltas take the packed kernels
    assert bad_fast_e == 0 and bad_fast_x == 0, stats
    assert bad_exact_e == 0 and bad_exact_x == 0, stats
    if extract_ok:
        assert flagged_x >= ties.size * 0.9            # every exact tie is sent to the exact path
    if delta >= 4:                                     # (small deltas make the constructed ties a large share)
        assert flagged_e + flagged_x < 0.1 * c.size    # ... and the flag stays rare


def test_psnr_from_sse():
    assert svs_b200.psnr_from_sse(0, 8, 8) == float("inf")
    assert abs(svs_b200.psnr_from_sse(64, 8, 8) - 20 * np.log10(255.0)) < 1e-9
