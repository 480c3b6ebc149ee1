"""Loader (and in-tree builder) of libsvs_b200.so - the C ABI declared in include/svs_b200.h.

There is deliberately no fallback: if the shared library is missing or no CUDA device can run
it, every compute entry point raises.  ctypes signatures mirror include/svs_b200.h one to one.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libsvs_b200.so")
SOURCES = [os.path.join(_PKG, "csrc", "svs_b200.cu")]
HEADERS = [os.path.join(_PKG, "csrc", n) for n in ("svs_math.cuh", "svs_quant.h", "svs_hw.cuh", "svs_block.cuh")] + \
          [os.path.join(_ROOT, "include", "svs_b200.h")]
# the round-1 kernel organisations, only compiled into -DSVS_WITH_VARIANTS measurement builds
VARIANT_HEADERS = [os.path.join(_PKG, "csrc", "variants", n) for n in ("svs_fast.cuh", "svs_tile.cuh", "svs_row.cuh")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]

_lib = None
_lib_override = None


def use_library(path):
    """Measurement scripts only (profiles/ab_kernels.py): load another BUILD of the same library,
    e.g. one compiled with -DSVS_WITH_VARIANTS.  Must be called before the first lib()."""
    global _lib_override
    if _lib is not None:
        raise RuntimeError("libsvs_b200.so is already loaded")
    if not os.path.exists(path):
        raise RuntimeError("no such library: %s" % path)
    _lib_override = path


class SvsError(RuntimeError):
    """A non-zero return code of the C ABI that is not an argument error."""


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libsvs_b200.so")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force=False, verbose=False, defines=(), out=None):
    """Compile the kernels for sm_100a in-tree (nvcc cross-compiles without a GPU).

    `defines` / `out`: measurement builds next to the product library, e.g.
    build(defines=["SVS_WITH_VARIANTS"], out="variants/libsvs_variants.so")."""
    target = out or LIB_PATH
    if not force and not defines and out is None and not needs_build():
        return LIB_PATH
    os.makedirs(os.path.dirname(os.path.abspath(target)), exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + \
          ["-I", os.path.join(_ROOT, "include"), "-I", os.path.join(_PKG, "csrc"), "-o", target] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
    if verbose:
        print(res.stderr)
    return target


_c = ctypes
_SIGNATURES = {
    "svs_version": (_c.c_int, []),
    "svs_last_error_string": (_c.c_char_p, []),
    "svs_capacity_bits": (_c.c_int64, [_c.c_int, _c.c_int, _c.c_int]),
    "svs_bits_row_bytes": (_c.c_int64, [_c.c_int, _c.c_int, _c.c_int]),
    "svs_kernel_launch_count": (_c.c_int64, []),
    "svs_debug_kernel_family": (_c.c_int, [_c.c_int]),
    "svs_memcpy_d2d_async": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_void_p]),
    "svs_set_reserved_sms": (_c.c_int, [_c.c_int]),
    "svs_extract_frames": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int64, _c.c_int, _c.c_int, _c.c_int64,
                                      _c.c_int64, _c.c_double, _c.c_int, _c.c_void_p, _c.c_int64, _c.c_void_p]),
    "svs_extract_frames_scatter": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int64, _c.c_int, _c.c_int, _c.c_int64,
                                              _c.c_int64, _c.c_double, _c.c_int, _c.c_void_p, _c.c_int64,
                                              _c.POINTER(_c.c_void_p), _c.c_int, _c.c_void_p]),
    "svs_extract_frames_multicast": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int64, _c.c_int, _c.c_int, _c.c_int64,
                                                _c.c_int64, _c.c_double, _c.c_int, _c.c_void_p, _c.c_void_p,
                                                _c.c_int64, _c.c_void_p]),
    "svs_embed_frames": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int64, _c.c_int, _c.c_int, _c.c_int64,
                                    _c.c_int64, _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_double, _c.c_int,
                                    _c.c_void_p, _c.c_int, _c.c_int64, _c.c_int64, _c.c_void_p, _c.c_void_p,
                                    _c.c_void_p, _c.c_void_p]),
    "svs_ctx_create": (_c.c_int, [_c.c_int, _c.c_int64, _c.POINTER(_c.c_void_p)]),
    "svs_ctx_destroy": (_c.c_int, [_c.c_void_p]),
    "svs_extract_frames_host": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int64, _c.c_int, _c.c_int,
                                           _c.c_int64, _c.c_int64, _c.c_double, _c.c_int, _c.c_void_p, _c.c_int64]),
    "svs_embed_frames_host": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int64, _c.c_int, _c.c_int,
                                         _c.c_int64, _c.c_int64, _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_double,
                                         _c.c_int, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """The loaded C ABI.  Raises if the library has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libsvs_b200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                "this package has no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(_lib_override or LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return lib().svs_last_error_string().decode("utf-8", "replace")


def check(rc, what):
    """Argument errors (<0) -> ValueError, CUDA errors (>0) -> SvsError."""
    if rc == 0:
        return
    msg = "%s failed (%d): %s" % (what, rc, last_error())
    if rc < 0:
        raise ValueError(msg)
    raise SvsError(msg)
