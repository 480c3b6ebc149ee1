"""Deterministic synthetic inputs shared by the tests, the golden-vector generator and bench.py.

Bytes come from SHAKE-256 so that the same (seed, shape) gives the same array on every NumPy
version and on every box; nothing here depends on the reference or on the oracle.
"""
from __future__ import annotations

import hashlib

import numpy as np


def synth_bytes(seed, n):
    return np.frombuffer(hashlib.shake_256(("svs:%s" % (seed,)).encode()).digest(int(n)), dtype=np.uint8)


def synth_frames(seed, shape, lo=0, hi=256):
    """uint8 array of `shape`, values uniform in [lo, hi)."""
    n = int(np.prod(shape))
    raw = synth_bytes(seed, n).astype(np.uint16)
    return (lo + ((raw * (hi - lo)) >> 8)).astype(np.uint8).reshape(shape)


def synth_bits(seed, nbits):
    """uint8 array of 0/1 of length nbits."""
    return np.unpackbits(synth_bytes("bits:%s" % (seed,), (int(nbits) + 7) // 8))[:int(nbits)]


def bits_to_str(bits):
    return (np.asarray(bits, dtype=np.uint8) + 48).tobytes().decode("ascii")


def str_to_bits(s):
    return np.frombuffer(s.encode("ascii"), dtype=np.uint8) - 48


def gradient_frame(h, w, channels=3, seed=0, noise=8):
    """Smooth ramp plus mild noise - a 'natural image'-like case."""
    y, x = np.mgrid[0:h, 0:w]
    base = (40 + 150.0 * (x / max(1, w - 1)) * (y / max(1, h - 1)) + 20 * np.sin(x / 7.0)).astype(np.float64)
    shape = (h, w, channels) if channels == 3 else (h, w)
    nz = synth_frames("g%s" % (seed,), shape, 0, 2 * noise + 1).astype(np.float64) - noise
    if channels == 3:
        base = base[..., None] + np.array([0.0, 10.0, -10.0])
    return np.clip(base + nz, 0, 255).astype(np.uint8)
