"""Builds tests/host_math/_build/libhost_math.so: csrc/svs_math.cuh compiled for the HOST with g++.

Test infrastructure: lets the CPU test-suite run the exact operation sequence of the CUDA
kernels (folded constants included) against the oracle without a GPU.
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "host_math.cpp")
HDR = os.path.join(ROOT, "secure-video-steganography-using-ecc-and-dct_b200", "csrc", "svs_math.cuh")
HDR2 = os.path.join(os.path.dirname(HDR), "svs_quant.h")
HDRS = [HDR, HDR2] + [os.path.join(os.path.dirname(HDR), n) for n in ("svs_hw.cuh", "svs_block.cuh")]
OUT = os.path.join(HERE, "_build", "libhost_math.so")


def build():
    if os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(p) for p in [SRC] + HDRS):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                    "-I", os.path.dirname(HDR), "-o", OUT, SRC], check=True, capture_output=True)
    return OUT


def load():
    L = ctypes.CDLL(build())
    for name in ("hm_dct2_fwd", "hm_dct2_inv", "hm_dct8_fwd", "hm_dct8_inv"):
        getattr(L, name).restype = None
        getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.c_long]
    L.hm_quant_check.restype = None
    L.hm_quant_check.argtypes = [ctypes.c_double, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
    u8p, c = ctypes.c_void_p, ctypes
    L.hm_blk_embed.restype = c.c_int
    L.hm_blk_embed.argtypes = [c.c_int, u8p, c.c_long, c.c_double, c.c_int, u8p, u8p, u8p]
    L.hm_blk_extract.restype = c.c_int
    L.hm_blk_extract.argtypes = [c.c_int, u8p, c.c_long, c.c_double, c.c_int, u8p]
    return L
