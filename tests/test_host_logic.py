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


def test_psnr_from_sse():
    assert svs_b200.psnr_from_sse(0, 8, 8) == float("inf")
    assert abs(svs_b200.psnr_from_sse(64, 8, 8) - 20 * np.log10(255.0)) < 1e-9
