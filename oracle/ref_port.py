"""Loop-structured CPU port of the reference hot path.  TEST / BASELINE INFRASTRUCTURE ONLY.

Where ``dctqim_oracle.py`` is the *fast* checker (vectorised over blocks), this module keeps the
reference's own execution structure - one ``scipy.fftpack`` call pair per 8x8 block and one
``round()`` per coefficient in the interpreter (/root/reference/config_and_setup.py:129-169) - so
that timing it on the GPU box's host cores reproduces what the reference costs there.  It is
what ``bench.py`` times for ``cpu_baseline`` (kind "port") and for ``--impl reference``: the
reference is pure Python and /root/reference does not exist on the GPU box, so there is no
``oracle/_ref`` build.  Checked against the live reference and the golden fixtures in
``tests/test_oracle.py``.  Never imported by the product package.
"""
from __future__ import annotations

import numpy as np

try:                                   # the reference's own dependencies (requirements.txt:1-5)
    import cv2
except Exception:                      # pragma: no cover - image without OpenCV
    cv2 = None
from scipy.fftpack import dct, idct

from .dctqim_oracle import bgr_to_gray

_LIMIT = 63


def _gray_of(frame):
    # config_and_setup.py:111-116
    if frame.ndim == 3 and frame.shape[2] == 3:
        if cv2 is not None:
            return cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
        return bgr_to_gray(frame)
    if frame.ndim == 2:
        return frame.copy()
    raise ValueError("Format frame input tidak didukung.")


def _fwd(block):
    return dct(dct(block, axis=0, norm='ortho'), axis=1, norm='ortho')    # :135


def _inv(block):
    return idct(idct(block, axis=0, norm='ortho'), axis=1, norm='ortho')  # :168


def embed_frame_loop(frame, delta, payload, num_ac=63):
    """Per-block / per-coefficient embed loop (config_and_setup.py:129-158,166-172)."""
    gray = _gray_of(np.asarray(frame))
    work = np.float32(gray)
    result = work.copy()
    rows, cols = work.shape
    total = len(payload) if payload else 0
    cursor = 0
    for top in range(0, rows, 8):
        if cursor >= total:
            break
        for left in range(0, cols, 8):
            if cursor >= total:
                break
            spectrum = _fwd(work[top:top + 8, left:left + 8])
            line = spectrum.flatten()
            edited = line.copy()
            for j in range(min(num_ac, _LIMIT)):
                if cursor >= total:
                    break
                if delta <= 0:
                    continue
                want = int(payload[cursor])
                level = int(round(line[j + 1] / delta))
                level = level - (level % 2) + want
                edited[j + 1] = float(level * delta)
                cursor += 1
            result[top:top + 8, left:left + 8] = _inv(edited.reshape(8, 8))
    return gray, np.uint8(np.clip(result, 0, 255)), cursor


def extract_frame_loop(frame, delta, num_ac=63):
    """Per-block / per-coefficient extract loop (config_and_setup.py:129-163,173-174)."""
    gray = _gray_of(np.asarray(frame))
    work = np.float32(gray)
    rows, cols = work.shape
    out = []
    for top in range(0, rows, 8):
        for left in range(0, cols, 8):
            line = _fwd(work[top:top + 8, left:left + 8]).flatten()
            for j in range(min(num_ac, _LIMIT)):
                if delta <= 0:
                    out.append('0')
                    continue
                out.append(str(int(round(line[j + 1] / delta)) % 2))
    return "".join(out)


def proses_frame_qim_dct(frame_bgr_input, mode, delta, bit_payload_segment=None,
                         enable_debug_prints_extract=False, num_ac_coeffs_to_use=63):
    """Same call surface as config_and_setup.py:106-109."""
    if mode == 'embed':
        return embed_frame_loop(frame_bgr_input, delta, bit_payload_segment, num_ac_coeffs_to_use)
    if mode == 'extract':
        return extract_frame_loop(frame_bgr_input, delta, num_ac_coeffs_to_use)
    _gray_of(np.asarray(frame_bgr_input))
    return None
