"""ctypes binding of the plain-C oracle (oracle/dctqim_oracle.c).  TEST INFRASTRUCTURE ONLY.

Used by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs to check the CUDA path
at full frame sizes.  Never imported by the product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsvs_oracle.so")
_lib = None

_u8p = ctypes.POINTER(ctypes.c_uint8)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f32p = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    """Compile the C oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "dctqim_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.svs_oracle_embed_frames.restype = ctypes.c_int
        L.svs_oracle_embed_frames.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
            ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int]
        L.svs_oracle_extract_frames.restype = ctypes.c_int
        L.svs_oracle_extract_frames.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
            ctypes.c_int64, ctypes.c_int64, ctypes.c_double, ctypes.c_int, ctypes.c_void_p,
            ctypes.c_int64, ctypes.c_int]
        for name in ("svs_oracle_dct8", "svs_oracle_idct8", "svs_oracle_dct2", "svs_oracle_idct2"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [_f32p]
        L.svs_oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _geometry(frames):
    a = np.ascontiguousarray(frames, dtype=np.uint8)
    if a.ndim == 4 and a.shape[3] == 3:
        f, h, w, ch = a.shape
    elif a.ndim == 3:
        (f, h, w), ch = a.shape, 1
    else:
        raise ValueError("frames must be (F,H,W,3) or (F,H,W) uint8")
    return a, f, h, w, ch


def embed_frames(frames, payload_packed, total_bits, delta, num_ac, bit_offset=0, threads=1,
                 want_gray=True):
    """Batch embed -> (stego (F,H,W) u8, gray (F,H,W) u8 or None, bits_embedded (F,) int64)."""
    a, f, h, w, ch = _geometry(frames)
    pay = np.ascontiguousarray(payload_packed, dtype=np.uint8)
    if pay.size == 0:
        pay = np.zeros(1, dtype=np.uint8)
    stego = np.empty((f, h, w), dtype=np.uint8)
    gray = np.empty((f, h, w), dtype=np.uint8) if want_gray else None
    nbits = np.zeros(f, dtype=np.int64)
    rc = lib().svs_oracle_embed_frames(
        a.ctypes.data, ch, f, h, w, h * w * ch, w * ch, pay.ctypes.data, int(bit_offset),
        int(total_bits), float(delta), int(num_ac), stego.ctypes.data,
        gray.ctypes.data if want_gray else None, nbits.ctypes.data, int(threads))
    if rc != 0:
        raise ValueError("svs_oracle_embed_frames rc=%d" % rc)
    return stego, gray, nbits


def extract_frames(frames, delta, num_ac, threads=1):
    """Batch extract -> (F, ceil(cap/8)) uint8, MSB-first packed bits."""
    a, f, h, w, ch = _geometry(frames)
    n = max(0, min(int(num_ac), 63))
    cap = (h // 8) * (w // 8) * n
    nbytes = (cap + 7) // 8
    out = np.zeros((f, nbytes), dtype=np.uint8)
    if nbytes:
        rc = lib().svs_oracle_extract_frames(a.ctypes.data, ch, f, h, w, h * w * ch, w * ch,
                                             float(delta), int(num_ac), out.ctypes.data, nbytes,
                                             int(threads))
        if rc != 0:
            raise ValueError("svs_oracle_extract_frames rc=%d" % rc)
    return out


def max_threads():
    return int(lib().svs_oracle_max_threads())


def dct2(block):
    b = np.ascontiguousarray(block, dtype=np.float32).copy()
    lib().svs_oracle_dct2(b.ctypes.data_as(_f32p))
    return b


def idct2(block):
    b = np.ascontiguousarray(block, dtype=np.float32).copy()
    lib().svs_oracle_idct2(b.ctypes.data_as(_f32p))
    return b
