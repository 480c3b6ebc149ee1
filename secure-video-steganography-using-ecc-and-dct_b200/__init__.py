"""B200-native (sm_100a) implementation of the reference's per-frame block-DCT + QIM hot path.

Drop-in for ``proses_frame_qim_dct`` (config_and_setup.py:106-174) plus a batched,
device-resident API; see DESIGN.md and INTEGRATION.md at the repository root.
"""
from ._native import build, lib, LIB_PATH, EXPORTED_SYMBOLS, SvsError
from .bitstream import (bits_from_str, bits_to_str, pack_bits, unpack_bits, pack_str,
                        bytes_to_bitstring, bitstring_to_bytes)
from .frame_path import (proses_frame_qim_dct, install, embed_frames, extract_frames, capacity_bits,
                         bits_row_bytes, psnr_from_sse, EmbedResult, MAX_AC)

from . import sharding
from . import pipeline

__all__ = [
    "sharding", "pipeline",
    "build", "lib", "LIB_PATH", "EXPORTED_SYMBOLS", "SvsError",
    "bits_from_str", "bits_to_str", "pack_bits", "unpack_bits", "pack_str", "bytes_to_bitstring",
    "bitstring_to_bytes", "proses_frame_qim_dct", "install", "embed_frames", "extract_frames",
    "capacity_bits", "bits_row_bytes", "psnr_from_sse", "EmbedResult", "MAX_AC",
]
