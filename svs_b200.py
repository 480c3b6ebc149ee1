"""Importable alias for the package directory ``secure-video-steganography-using-ecc-and-dct_b200``
(a hyphenated directory name cannot be written in an ``import`` statement).

    import svs_b200
    svs_b200.proses_frame_qim_dct(frame, 'extract', 20, num_ac_coeffs_to_use=10)
"""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("secure-video-steganography-using-ecc-and-dct_b200")
sys.modules[__name__] = _pkg
