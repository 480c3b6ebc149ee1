"""CPU oracle for the per-frame 8x8 block-DCT + parity-QIM path.  TEST INFRASTRUCTURE ONLY.

This file is a vectorised NumPy restatement of the reference function
``proses_frame_qim_dct`` (/root/reference/config_and_setup.py:106-174).  It is the checker
that the CUDA path is compared against; nothing in the product package may import it.  Only
``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` use it.

Pinning: the reference has no tests or golden vectors for this path, so the
oracle is pinned to *outputs of the reference itself*: ``tests/golden/make_golden.py`` imports
the real function from /root/reference and stores its outputs; ``tests/test_oracle.py`` checks
this restatement bit-for-bit against those fixtures (and against the live reference whenever
/root/reference is present).

The float arithmetic lives in third-party code that the reference does not vendor or pin:
  * scipy.fftpack.dct/idct (DUCC/pocketfft real-FFT plan for N=8), scipy 1.18.1 in this image,
    called at config_and_setup.py:135 and :168;
  * cv2.cvtColor(BGR2GRAY), OpenCV 4.13.0 fixed-point formula, called at config_and_setup.py:112;
  * NumPy >= 2 scalar promotion (float32 / python number -> float32), NumPy 2.3.5 here,
    config_and_setup.py:148,160.
Their published algorithms are restated below op for op (every line is one IEEE binary32
operation, round-to-nearest-even, no FMA contraction - NumPy never fuses).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# fl32(cos(k*pi/16)), k = 1..7, as bit patterns so that no libm is involved.
_T_BITS = (0x3F7B14BE, 0x3F6C835E, 0x3F54DB31, 0x3F3504F3, 0x3F0E39DA, 0x3EC3EF15, 0x3E47C5C2)
T = (None,) + tuple(np.array([b], dtype=np.uint32).view(np.float32)[0] for b in _T_BITS)
W = T[4]                                                        # fl32(sqrt(1/2))
S2 = np.array([0x3FB504F3], dtype=np.uint32).view(np.float32)[0]  # fl32(sqrt(2))
QUARTER = F32(0.25)
HALF = F32(0.5)
TWO = F32(2.0)

MAX_AC = 63
BLOCK = 8


# ----------------------------------------------------------------------------------------------
# 8-point transforms (scipy.fftpack.dct/idct, type 2, norm='ortho', float32)
# ----------------------------------------------------------------------------------------------
def dct8(x):
    """Forward orthonormal DCT-II of 8 float32 arrays x[0..7] -> list X[0..7].

    Follows pocketfft's rfftp plan for N=8 (radix-2 pass, then radix-4 pass) wrapped in its
    DCT-II pre/post-processing, as scipy 1.18.1 executes it for config_and_setup.py:135.
    """
    c = [np.asarray(v, dtype=F32) for v in x]
    c[0] = TWO * c[0]
    c[7] = TWO * c[7]
    for k in (1, 3, 5):
        t = c[k + 1]
        c[k + 1] = c[k + 1] - c[k]
        c[k] = c[k] + t
    h = [None] * 8
    h[0] = c[0] + c[7]
    h[4] = c[0] - c[7]
    h[3] = TWO * c[3]
    h[7] = -(TWO * c[4])
    h[1] = c[1] + c[5]
    tr = c[1] - c[5]
    ti = c[2] + c[6]
    h[2] = c[2] - c[6]
    wti = W * ti
    wtr = W * tr
    h[6] = wti + wtr
    h[5] = wtr - wti
    o = [None] * 8
    for k in (0, 1):
        a0, a1, a2, a3 = h[4 * k], h[4 * k + 1], h[4 * k + 2], h[4 * k + 3]
        p = a0 + a3
        m = a0 - a3
        u = TWO * a1
        v = TWO * a2
        o[k] = p + u
        o[k + 4] = p - u
        o[k + 6] = m + v
        o[k + 2] = m - v
    c = [QUARTER * v for v in o]
    X = [None] * 8
    for k, kc in ((1, 7), (2, 6), (3, 5)):
        t1 = T[k] * c[kc] + T[kc] * c[k]
        t2 = T[k] * c[k] - T[kc] * c[kc]
        X[k] = HALF * (t1 + t2)
        X[kc] = HALF * (t1 - t2)
    X[4] = c[4] * T[4]
    X[0] = c[0] * (S2 * HALF)
    return X


def idct8(X):
    """Inverse (DCT-III) of 8 float32 arrays, scipy.fftpack.idct(type=2, norm='ortho') order.

    config_and_setup.py:168.
    """
    c = [np.asarray(v, dtype=F32) for v in X]
    c[0] = c[0] * S2
    for k, kc in ((1, 7), (2, 6), (3, 5)):
        t1 = c[k] + c[kc]
        t2 = c[k] - c[kc]
        c[k] = T[k] * t2 + T[kc] * t1
        c[kc] = T[k] * t1 - T[kc] * t2
    c[4] = c[4] * (TWO * T[4])
    h = [None] * 8
    for k in (0, 1):
        r1 = c[k + 6] + c[k + 2]
        h[4 * k + 2] = c[k + 6] - c[k + 2]
        r2 = c[k] + c[k + 4]
        h[4 * k + 1] = c[k] - c[k + 4]
        h[4 * k] = r2 + r1
        h[4 * k + 3] = r2 - r1
    o = [None] * 8
    o[0] = h[0] + h[4]
    o[7] = h[0] - h[4]
    o[4] = -h[7]
    o[3] = h[3]
    wh5 = W * h[5]
    wh6 = W * h[6]
    tr = wh5 + wh6
    ti = wh6 - wh5
    o[1] = h[1] + tr
    o[5] = h[1] - tr
    o[2] = ti + h[2]
    o[6] = ti - h[2]
    c = [QUARTER * v for v in o]
    for k in (1, 3, 5):
        t = c[k]
        c[k] = c[k] - c[k + 1]
        c[k + 1] = c[k + 1] + t
    return c


def dct2_blocks(blocks):
    """(N,8,8) float32 -> (N,8,8): axis-0 pass (down the columns) then axis-1 pass.

    config_and_setup.py:135 - dct(dct(b, axis=0), axis=1); each pass rounds to float32.
    """
    b = np.asarray(blocks, dtype=F32)
    col = np.stack(dct8([b[:, r, :] for r in range(8)]), axis=1)      # transform over rows index
    return np.stack(dct8([col[:, :, v] for v in range(8)]), axis=2)   # then over columns index


def idct2_blocks(coefs):
    """Inverse of dct2_blocks in the reference's order (axis 0 first, then axis 1); :168."""
    b = np.asarray(coefs, dtype=F32)
    col = np.stack(idct8([b[:, r, :] for r in range(8)]), axis=1)
    return np.stack(idct8([col[:, :, v] for v in range(8)]), axis=2)


# ----------------------------------------------------------------------------------------------
# Frame-level pieces
# ----------------------------------------------------------------------------------------------
def bgr_to_gray(frame):
    """cv2.cvtColor(COLOR_BGR2GRAY) of OpenCV 4.13: 15-bit fixed point, config_and_setup.py:112."""
    f = np.asarray(frame)
    b = f[..., 0].astype(np.int32)
    g = f[..., 1].astype(np.int32)
    r = f[..., 2].astype(np.int32)
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)


def to_gray(frame):
    """Shape dispatch of config_and_setup.py:111-116 (ValueError text kept verbatim)."""
    f = np.asarray(frame)
    if f.ndim == 3 and f.shape[2] == 3:
        return bgr_to_gray(f)
    if f.ndim == 2:
        return f.copy()
    raise ValueError("Format frame input tidak didukung.")


def _to_blocks(img):
    h, w = img.shape
    return img.reshape(h // 8, 8, w // 8, 8).transpose(0, 2, 1, 3).reshape(-1, 8, 8)


def _from_blocks(blocks, h, w):
    return blocks.reshape(h // 8, w // 8, 8, 8).transpose(0, 2, 1, 3).reshape(h, w)


def quant_index(coefs, delta):
    """int(round(c / delta)) of config_and_setup.py:148,160.

    NumPy >= 2: float32 scalar / python number is a float32 division with delta cast to
    float32 first; builtins.round on np.float32 is round-half-to-even.
    """
    t = np.asarray(coefs, dtype=F32) / F32(delta)
    return np.rint(t).astype(np.int64)


def _requant_value(qnew, delta):
    """float(q * delta) stored into a float32 array (config_and_setup.py:156)."""
    if isinstance(delta, (int, np.integer)):
        return (qnew * int(delta)).astype(np.float64).astype(F32)
    return (qnew.astype(np.float64) * float(delta)).astype(F32)


def bits_from_any(bits):
    """'0'/'1' str, bytes of 0/1, or a 0/1 integer array -> uint8 array of 0/1."""
    if bits is None:
        return np.zeros(0, dtype=np.uint8)
    if isinstance(bits, str):
        a = np.frombuffer(bits.encode("ascii"), dtype=np.uint8) - 48
        if a.size and a.max() > 1:
            raise ValueError("payload string must contain only '0' and '1'")
        return a
    return np.asarray(bits, dtype=np.uint8)


def capacity_bits(h, w, num_ac):
    """Bits one frame carries: blocks * min(num_ac, 63)  (config_and_setup.py:138)."""
    n = max(0, min(int(num_ac), MAX_AC))
    return (h // 8) * (w // 8) * n


def embed_frame(frame, delta, payload_bits, num_ac=63):
    """mode='embed' of proses_frame_qim_dct -> (gray u8, stego u8, bits_embedded).

    payload_bits may be longer than the frame's capacity (the caller passes the whole
    remaining payload, embed_process.py:116-121).
    """
    gray = to_gray(frame)
    h, w = gray.shape
    if h % 8 or w % 8:
        raise ValueError("frame height and width must be multiples of 8")
    bits = bits_from_any(payload_bits)
    nbits = int(bits.size)
    n = min(int(num_ac), MAX_AC)
    if nbits == 0:                                   # :125-126,130 - loop breaks at once
        return gray, gray.copy(), 0
    blocks = _to_blocks(gray.astype(F32))
    nblk = blocks.shape[0]
    active = n > 0 and delta > 0                     # :143-145 'continue' consumes nothing
    if active:
        used_before = np.arange(nblk, dtype=np.int64) * n
        k = np.clip(nbits - used_before, 0, n)       # coefficients modified in each block
        processed = used_before < nbits              # :130,132 early exit
        embedded = int(min(nbits, nblk * n))
    else:
        k = np.zeros(nblk, dtype=np.int64)
        processed = np.ones(nblk, dtype=bool)        # index never advances -> every block
        embedded = 0
    coefs = dct2_blocks(blocks)
    flat = coefs.reshape(nblk, 64).copy()
    if active:
        idx = np.arange(n, dtype=np.int64)
        sel = idx[None, :] < k[:, None]              # (nblk, n) coefficient is touched
        pos = np.minimum(used_before[:, None] + idx[None, :], nbits - 1)
        bitv = bits[pos].astype(np.int64)
        q = quant_index(flat[:, 1:n + 1], delta)
        qnew = q - (q % 2) + bitv                    # :149-155 asymmetric parity fix-up
        newc = _requant_value(qnew, delta)
        flat[:, 1:n + 1] = np.where(sel, newc, flat[:, 1:n + 1])
    pix = idct2_blocks(flat.reshape(nblk, 8, 8))
    out = np.where(processed[:, None, None], pix, blocks)
    stego_f = _from_blocks(out, h, w)
    stego = np.clip(stego_f, 0, 255).astype(np.uint8)     # clip then truncate, :171
    return gray, stego, embedded


def extract_frame_bits(frame, delta, num_ac=63):
    """mode='extract' -> uint8 array of 0/1, length blocks*min(num_ac,63) (:159-163,173-174)."""
    gray = to_gray(frame)
    h, w = gray.shape
    if h % 8 or w % 8:
        raise ValueError("frame height and width must be multiples of 8")
    n = min(int(num_ac), MAX_AC)
    nblk = (h // 8) * (w // 8)
    if n <= 0:
        return np.zeros(0, dtype=np.uint8)
    if not delta > 0:
        return np.zeros(nblk * n, dtype=np.uint8)
    coefs = dct2_blocks(_to_blocks(gray.astype(F32))).reshape(nblk, 64)
    q = quant_index(coefs[:, 1:n + 1], delta)
    return (q % 2).astype(np.uint8).reshape(-1)


def extract_frame(frame, delta, num_ac=63):
    """mode='extract' with the reference's return type (a '0'/'1' str)."""
    return bits_to_str(extract_frame_bits(frame, delta, num_ac))


def proses_frame_qim_dct(frame_bgr_input, mode, delta, bit_payload_segment=None,
                         enable_debug_prints_extract=False, num_ac_coeffs_to_use=63):
    """Oracle with the reference's exact signature (config_and_setup.py:106-109)."""
    if mode == 'embed':
        return embed_frame(frame_bgr_input, delta, bit_payload_segment, num_ac_coeffs_to_use)
    if mode == 'extract':
        return extract_frame(frame_bgr_input, delta, num_ac_coeffs_to_use)
    to_gray(frame_bgr_input)
    return None


# ----------------------------------------------------------------------------------------------
# Bit-string helpers (MSB-first packing = bytes_ke_bitstream, config_and_setup.py:22-23)
# ----------------------------------------------------------------------------------------------
def bits_to_str(bits):
    return (np.asarray(bits, dtype=np.uint8) + 48).tobytes().decode("ascii")


def pack_bits(bits):
    return np.packbits(bits_from_any(bits), bitorder="big")


def unpack_bits(packed, nbits):
    return np.unpackbits(np.asarray(packed, dtype=np.uint8), bitorder="big")[:nbits]
