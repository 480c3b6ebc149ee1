"""Host side of the B200 DCT-QIM frame path.

Two surfaces over the C ABI (include/svs_b200.h):

* ``proses_frame_qim_dct`` - the reference's own call surface (config_and_setup.py:106-109):
  numpy frame in; ``(gray, stego, bits_embedded)`` or a '0'/'1' string out.  ``install()``
  rebinds it into the reference's modules (embed_process.py:8-14, extract_process.py:9-14).
* ``embed_frames`` / ``extract_frames`` - batched, device-resident (torch CUDA tensors, packed
  bits), stream-ordered; what bench.py and the multi-GPU driver call.

torch is only plumbing here (device memory, streams); all arithmetic is in csrc/svs_b200.cu.
There is no CPU path: without the built library and a CUDA device these functions raise.
"""
from __future__ import annotations

import collections
import ctypes
import math
import threading

import numpy as np

from . import _native
from .bitstream import bits_from_str, bits_to_str, pack_bits

MAX_AC = 63

EmbedResult = collections.namedtuple("EmbedResult", "stego gray bits_embedded sse")


def capacity_bits(height, width, num_ac):
    """Bits per frame: (H/8)*(W/8)*min(num_ac, 63) (config_and_setup.py:138)."""
    return (int(height) // 8) * (int(width) // 8) * max(0, min(int(num_ac), MAX_AC))


def bits_row_bytes(height, width, num_ac):
    """Row pitch of the packed-bit output that enables 32-bit stores in the extract kernel."""
    return int(_native.lib().svs_bits_row_bytes(int(height), int(width), int(num_ac)))


def psnr_from_sse(sse, height, width):
    """cv2.PSNR(gray, stego) from the fused sum of squared errors (embed_process.py:205)."""
    sse = float(sse)
    if sse <= 0:
        return float("inf")
    return 10.0 * math.log10(255.0 * 255.0 * height * width / sse)


# ----------------------------------------------------------------------------------------------
# reference-compatible single-frame entry point (host buffers through svs_*_frames_host)
# ----------------------------------------------------------------------------------------------
class _HostContext:
    """One svs_ctx per (thread, device); the C context must not be shared between threads."""

    _local = threading.local()

    @classmethod
    def get(cls, device=None):
        import ctypes
        if device is None:
            device = 0
            try:
                import torch
                if torch.cuda.is_available():
                    device = torch.cuda.current_device()
            except Exception:
                pass
        table = getattr(cls._local, "table", None)
        if table is None:
            table = cls._local.table = {}
        ctx = table.get(device)
        if ctx is None:
            handle = ctypes.c_void_p()
            _native.check(_native.lib().svs_ctx_create(int(device), 0, ctypes.byref(handle)), "svs_ctx_create")
            ctx = table[device] = handle
        return ctx


def _frame_geometry(frame):
    """(array, channels, H, W, row_stride) for an HxWx3 / HxW uint8 array; copies only if needed."""
    if not isinstance(frame, np.ndarray):
        frame = np.asarray(frame)
    if frame.ndim == 3 and frame.shape[2] == 3:
        ch = 3
    elif frame.ndim == 2:
        ch = 1
    else:
        raise ValueError("Format frame input tidak didukung.")            # config_and_setup.py:116
    if frame.dtype != np.uint8:
        frame = frame.astype(np.uint8)
    h, w = frame.shape[:2]
    if h == 0 or w == 0 or h % 8 or w % 8:
        raise ValueError("tinggi/lebar frame harus kelipatan 8 (got %dx%d)" % (h, w))
    inner_ok = frame.strides[-1] == 1 and (ch == 1 or frame.strides[1] == 3)
    if not inner_ok or frame.strides[0] < w * ch:
        frame = np.ascontiguousarray(frame)          # e.g. channel-reversed or transposed views
    return frame, ch, h, w, frame.strides[0]


def proses_frame_qim_dct(frame_bgr_input, mode, delta, bit_payload_segment=None,
                         enable_debug_prints_extract=False, num_ac_coeffs_to_use=63):
    """Drop-in for config_and_setup.py:106-174 running on the GPU.

    embed  -> (gray HxW uint8, stego HxW uint8, bits_embedded int)
    extract -> str of '0'/'1', length (H/8)*(W/8)*min(num_ac,63)
    """
    frame, ch, h, w, row_stride = _frame_geometry(frame_bgr_input)
    if mode not in ("embed", "extract"):
        return None                                   # the reference falls off the end
    L = _native.lib()
    ctx = _HostContext.get()
    num_ac = int(num_ac_coeffs_to_use)
    delta = float(delta)
    if mode == "embed":
        cap = capacity_bits(h, w, num_ac) if delta > 0 else 0
        seg = bit_payload_segment if bit_payload_segment else ""
        if cap > 0:
            bits = bits_from_str(seg, cap)            # the caller passes the whole remaining payload
            total = int(bits.size)
        else:                                         # delta <= 0 / num_ac <= 0: only emptiness matters
            bits = np.zeros(1, np.uint8)
            total = 1 if len(seg) else 0
        packed = pack_bits(bits) if total else np.zeros(4, np.uint8)
        stego = np.empty((h, w), np.uint8)
        gray = np.empty((h, w), np.uint8)
        nbits = np.zeros(1, np.int64)
        rc = L.svs_embed_frames_host(ctx, frame.ctypes.data, ch, 1, h, w, h * row_stride, row_stride,
                                     packed.ctypes.data, 0, total, delta, num_ac,
                                     stego.ctypes.data, 1, gray.ctypes.data, nbits.ctypes.data, None)
        _native.check(rc, "svs_embed_frames_host")
        return gray, stego, int(nbits[0])
    cap = capacity_bits(h, w, num_ac)
    if cap == 0:
        return ""
    nbytes = (cap + 7) // 8
    out = np.zeros(nbytes, np.uint8)
    rc = L.svs_extract_frames_host(ctx, frame.ctypes.data, ch, 1, h, w, h * row_stride, row_stride,
                                   delta, num_ac, out.ctypes.data, nbytes)
    _native.check(rc, "svs_extract_frames_host")
    return bits_to_str(np.unpackbits(out)[:cap])


def install(*modules):
    """Rebind the hot function in the reference's modules.

    With no arguments, patches config_and_setup / embed_process / extract_process if they are
    already imported; returns the list of patched module names.
    """
    import sys
    if not modules:
        modules = [sys.modules[m] for m in ("config_and_setup", "embed_process", "extract_process") if m in sys.modules]
    done = []
    for m in modules:
        setattr(m, "proses_frame_qim_dct", proses_frame_qim_dct)
        done.append(getattr(m, "__name__", str(m)))
    return done


# ----------------------------------------------------------------------------------------------
# batched, device-resident API
# ----------------------------------------------------------------------------------------------
def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device visible: the svs_b200 frame path has no CPU fallback")
    return torch


def _batch_geometry(frames):
    torch = _torch()
    if not (isinstance(frames, torch.Tensor) and frames.is_cuda and frames.dtype == torch.uint8):
        raise TypeError("frames must be a CUDA uint8 tensor of shape (F,H,W,3) or (F,H,W)")
    if frames.dim() == 4 and frames.shape[3] == 3:
        ch = 3
        if frames.stride(3) != 1 or frames.stride(2) != 3:
            frames = frames.contiguous()
    elif frames.dim() == 3:
        ch = 1
        if frames.stride(2) != 1:
            frames = frames.contiguous()
    else:
        raise ValueError("Format frame input tidak didukung.")
    f, h, w = frames.shape[:3]
    if h % 8 or w % 8 or h == 0 or w == 0:
        raise ValueError("frame height and width must be positive multiples of 8 (got %dx%d)" % (h, w))
    return frames, ch, int(f), int(h), int(w), int(frames.stride(0)), int(frames.stride(1))


def _stream_handle(stream):
    torch = _torch()
    if stream is None:
        stream = torch.cuda.current_stream()
    return stream.cuda_stream


def embed_frames(frames, payload, total_bits, delta, num_ac=63, *, bit_offset=0, stego_channels=1,
                 want_gray=False, want_bits_embedded=False, want_sse=False, out=None, stream=None):
    """Embed a batch resident in HBM.  Enqueues ONE kernel on `stream`; nothing synchronises.

    frames   (F,H,W,3) BGR or (F,H,W) gray CUDA uint8 (strided views are fine)
    payload  CUDA uint8 tensor, MSB-first packed bits, 4-byte aligned; frame f consumes bits
             [bit_offset + f*cap, ...) while total_bits lasts (embed_process.py:115-128)
    returns  EmbedResult(stego (F,H,W[,3]), gray|None, bits_embedded (F,) int64|None, sse (F,) uint64-as-int64|None)
    """
    torch = _torch()
    frames, ch, f, h, w, fs, rs = _batch_geometry(frames)
    dev = frames.device
    if out is None:
        shape = (f, h, w) if stego_channels == 1 else (f, h, w, 3)
        out = torch.empty(shape, dtype=torch.uint8, device=dev)
    gray = torch.empty((f, h, w), dtype=torch.uint8, device=dev) if want_gray else None
    nbits = torch.empty((f,), dtype=torch.int64, device=dev) if want_bits_embedded else None
    sse = torch.zeros((f,), dtype=torch.int64, device=dev) if want_sse else None
    if payload is None or int(total_bits) <= 0:
        pay_ptr, total_bits = None, 0
    else:
        if not (payload.is_cuda and payload.dtype == torch.uint8 and payload.is_contiguous()):
            raise TypeError("payload must be a contiguous CUDA uint8 tensor (packed bits)")
        if payload.numel() * 8 < int(bit_offset) + int(total_bits):
            raise ValueError("payload holds %d bits, bit_offset + total_bits = %d"
                             % (payload.numel() * 8, int(bit_offset) + int(total_bits)))
        if payload.numel() % 4 or payload.data_ptr() % 4:
            # the kernels read whole 32-bit words (include/svs_b200.h): give them a padded, aligned copy
            padded = torch.zeros((payload.numel() + 3) // 4 * 4, dtype=torch.uint8, device=payload.device)
            padded[:payload.numel()] = payload
            payload = padded
        pay_ptr = payload.data_ptr()
    with torch.cuda.device(dev):
        rc = _native.lib().svs_embed_frames(
            frames.data_ptr(), ch, f, h, w, fs, rs, pay_ptr, int(bit_offset), int(total_bits), float(delta),
            int(num_ac), out.data_ptr(), int(stego_channels), int(out.stride(0)), int(out.stride(1)),
            gray.data_ptr() if want_gray else None, nbits.data_ptr() if want_bits_embedded else None,
            sse.data_ptr() if want_sse else None, _stream_handle(stream))
    _native.check(rc, "svs_embed_frames")
    return EmbedResult(out, gray, nbits, sse)


def extract_frames(frames, delta, num_ac=63, *, out=None, stream=None, peer_ptrs=None, multicast_ptr=None):
    """Extract a batch resident in HBM -> (F, ceil(cap/8)) uint8 view of MSB-first packed bits.

    The returned tensor is a view of a buffer whose row pitch is bits_row_bytes(...) so that the
    kernel can use 32-bit stores; pass `out` (F, pitch) to reuse a buffer.

    Fused all-gather of a frame-sharded job (sharding.FusedExtractGather): `peer_ptrs` = device
    addresses (ints) of the same rows inside every OTHER rank's gathered buffer (peer-mapped over
    NVLink), or `multicast_ptr` = one NVSwitch multicast address of those rows on ALL ranks; the
    extract kernel then stores every packed word there as well (svs_extract_frames_scatter /
    svs_extract_frames_multicast) - no separate collective.
    """
    torch = _torch()
    frames, ch, f, h, w, fs, rs = _batch_geometry(frames)
    cap = capacity_bits(h, w, num_ac)
    nbytes = (cap + 7) // 8
    if cap == 0:
        return torch.empty((f, 0), dtype=torch.uint8, device=frames.device)
    if out is None:
        out = torch.empty((f, bits_row_bytes(h, w, num_ac)), dtype=torch.uint8, device=frames.device)
    L = _native.lib()
    with torch.cuda.device(frames.device):
        if multicast_ptr:
            rc = L.svs_extract_frames_multicast(frames.data_ptr(), ch, f, h, w, fs, rs, float(delta), int(num_ac),
                                                int(multicast_ptr), out.data_ptr(), int(out.stride(0)),
                                                _stream_handle(stream))
            what = "svs_extract_frames_multicast"
        elif peer_ptrs:
            arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
            rc = L.svs_extract_frames_scatter(frames.data_ptr(), ch, f, h, w, fs, rs, float(delta), int(num_ac),
                                              out.data_ptr(), int(out.stride(0)), arr, len(peer_ptrs),
                                              _stream_handle(stream))
            what = "svs_extract_frames_scatter"
        else:
            rc = L.svs_extract_frames(frames.data_ptr(), ch, f, h, w, fs, rs, float(delta), int(num_ac),
                                      out.data_ptr(), int(out.stride(0)), _stream_handle(stream))
            what = "svs_extract_frames"
    _native.check(rc, what)
    return out[:, :nbytes]
