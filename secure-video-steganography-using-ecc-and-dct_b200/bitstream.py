"""'0'/'1' strings <-> MSB-first packed bytes (host side).

The reference moves payload and extracted bits around as Python strings of '0'/'1'
(bytes_ke_bitstream / bitstream_ke_bytes, config_and_setup.py:22-41); the kernels use the packed
form with the same bit order (format(b, '08b'): most significant bit first).
"""
from __future__ import annotations

import numpy as np


def bits_from_str(s, limit=None):
    """'0101...' -> uint8 array of 0/1 (only the first `limit` characters are looked at)."""
    if limit is not None:
        s = s[:limit]
    a = np.frombuffer(s.encode("ascii"), dtype=np.uint8) - np.uint8(48)
    if a.size and int(a.max()) > 1:
        bad = s[int(np.argmax(a > 1))]
        raise ValueError("invalid literal for int() with base 10: %r" % bad)   # what int(ch) raises
    return a


def bits_to_str(bits):
    return (np.asarray(bits, dtype=np.uint8) + np.uint8(48)).tobytes().decode("ascii")


def pack_bits(bits01):
    """0/1 array -> packed uint8 (MSB-first), padded with zero bits."""
    return np.packbits(np.asarray(bits01, dtype=np.uint8), bitorder="big")


def unpack_bits(packed, nbits):
    return np.unpackbits(np.asarray(packed, dtype=np.uint8), bitorder="big")[:nbits]


def pack_str(s, limit=None):
    b = bits_from_str(s, limit)
    return pack_bits(b), int(b.size)


def bytes_to_bitstring(data):
    """bytes -> '0'/'1' str, same as bytes_ke_bitstream (config_and_setup.py:22-23)."""
    return bits_to_str(np.unpackbits(np.frombuffer(bytes(data), dtype=np.uint8)))


def bitstring_to_bytes(s):
    """'0'/'1' str -> bytes like bitstream_ke_bytes (:25-30): trailing bits beyond a multiple of
    8 are dropped; an input that becomes empty that way raises."""
    rest = len(s) % 8
    if rest:
        s = s[:-rest]
        if not s:
            raise ValueError("Bitstream kosong setelah dipotong.")
    return pack_bits(bits_from_str(s)).tobytes()
