"""Batched video pipeline around the frame path.

The reference drives the hot function one frame at a time from two Python loops:
``embed_process.py:108-144`` (read frame -> embed the REST of the payload string -> GRAY2BGR ->
write; once the payload is exhausted the remaining frames are copied in colour) and
``extract_process.py:55-86,173-182`` (read frame -> extract -> append to an ever growing '0'/'1'
string until the header, then the ciphertext, are complete).  With the frame path on a B200 those
loops, the per-frame host<->device hops and the string handling are what is left, so this module
replaces just them:

* `embed_frame_stream`   reads N frames, embeds them with ONE batched launch (frame f gets
                         payload bits [f*cap, (f+1)*cap), the fused GRAY2BGR store of N2 writes
                         what the FFV1 writer consumes) and passes the frames after the payload
                         through untouched - the same frames, in the same order, as the loop;
* `StegoBitReader`       serves the extracted stream by bytes (N4: packed, no strings), pulling
                         and extracting further frames in batches only when a caller asks for
                         more bits than it has - exactly the reference's read-as-needed order;
* `embed_gambar_ke_video_final` / `ekstraksi_gambar_video_final`
                         the reference's two pipeline functions with their signatures, prints
                         reduced to the essentials, built from the two pieces above.  Crypto
                         (ECDH P-256, HKDF, AES-256-GCM, SHA3-256), key handling and the secret
                         image codec are NOT re-implemented: they are called through the
                         reference's own modules (`config_and_setup`, `helpers`), which must be
                         importable (or passed as `ref=`).  Only the byte layout of the payload
                         (``embed_process.py:62-74``; every field is byte aligned) is restated.

`embed_fn` / `extract_fn` default to the CUDA kernels; the host-logic tests inject CPU stand-ins.
"""
from __future__ import annotations

import math
import os
import struct
import types

import numpy as np

from . import frame_path
from .bitstream import bytes_to_bitstring, bitstring_to_bytes

HEADER_BITS = 976        # everything before the ciphertext for P-256 / 16-byte salt / SHA3-256 / GCM


# ----------------------------------------------------------------------------------------------
# compute back-ends (GPU by default)
# ----------------------------------------------------------------------------------------------
class GpuFrameOps:
    """Batched embed / extract on one CUDA device.

    The packed payload is uploaded once; frames travel through reusable PINNED host buffers
    (asynchronous H2D / D2H on the device's copy engines, no pageable staging copies inside the
    driver), and the results are handed back as views of pinned memory that stay valid until the
    next call of the same kind - the caller (the writer thread of `embed_frame_stream`) consumes
    them before that."""

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device visible: the svs_b200 frame path has no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._payload = None
        self._payload_src = None
        self._pinned = {}

    def _device_payload(self, packed):
        if self._payload_src is not packed:
            pad = (-len(packed)) % 4 + 4                      # word loads may touch the next word
            host = np.concatenate([np.asarray(packed, np.uint8), np.zeros(pad, np.uint8)])
            self._payload = self.torch.from_numpy(host).to(self.device)
            self._payload_src = packed
        return self._payload

    def _pin(self, key, shape):
        """A pinned uint8 buffer of at least `shape` (grown geometrically, reused across batches)."""
        n = int(np.prod(shape))
        buf = self._pinned.get(key)
        if buf is None or buf.numel() < n:
            buf = self.torch.empty(max(n, 2 * (buf.numel() if buf is not None else 0)), dtype=self.torch.uint8).pin_memory()
            self._pinned[key] = buf
        return buf[:n].view(*shape)

    def _upload(self, frames):
        frames = np.asarray(frames)
        stage = self._pin("in", frames.shape)
        stage.numpy()[...] = frames                            # one host copy into pinned memory
        return stage.to(self.device, non_blocking=True)

    def embed(self, frames, packed, bit_offset, nbits, delta, num_ac, want_gray):
        """frames (k,h,w,3|none) u8 -> (stego_bgr (k,h,w,3), gray of frame 0 or None, bits per frame)."""
        t = self.torch
        d = self._upload(frames)
        res = frame_path.embed_frames(d, self._device_payload(packed), int(nbits), delta, num_ac,
                                      bit_offset=int(bit_offset), stego_channels=3, want_gray=bool(want_gray),
                                      want_bits_embedded=True)
        out = self._pin("stego", tuple(res.stego.shape))
        out.copy_(res.stego, non_blocking=True)
        gray0 = res.gray[0].cpu().numpy() if want_gray else None
        nb = res.bits_embedded.cpu().numpy()                   # synchronises the stream: `out` is complete
        t.cuda.current_stream(self.device).synchronize()
        return out.numpy(), gray0, nb

    def extract(self, frames, delta, num_ac):
        """frames (k,h,w[,3]) u8 -> (k, ceil(cap/8)) packed bits."""
        d = self._upload(frames)
        return frame_path.extract_frames(d, delta, num_ac).cpu().numpy()


_default_ops = None


def _ops():
    global _default_ops
    if _default_ops is None:
        _default_ops = GpuFrameOps()
    return _default_ops


# ----------------------------------------------------------------------------------------------
# embed: the frame loop of embed_process.py:108-144
# ----------------------------------------------------------------------------------------------
class _Prefetch:
    """Decode ahead on a worker thread: batches of up to `n` frames (cv2 releases the GIL while it
    decodes, so this overlaps the GPU call and the encoder)."""

    def __init__(self, read_frame, n, crop, depth=2):
        import queue
        import threading
        self.q = queue.Queue(maxsize=depth)
        self.err = None

        def work():
            try:
                while True:
                    batch = []
                    while len(batch) < n:
                        ok, frame = read_frame()
                        if not ok:
                            break
                        batch.append(crop(frame))
                    self.q.put(batch)
                    if len(batch) < n:
                        if batch:
                            self.q.put([])
                        return
            except BaseException as exc:            # surfaced on the consumer side
                self.err = exc
                self.q.put([])

        self.t = threading.Thread(target=work, daemon=True)
        self.t.start()

    def next(self):
        batch = self.q.get()
        if self.err is not None:
            raise self.err
        return batch


class _Writer:
    """Encode behind on a worker thread (FFV1 encoding is the slowest stage of the pipeline)."""

    def __init__(self, write_frame, depth=2):
        import queue
        import threading
        self.q = queue.Queue(maxsize=depth)
        self.err = None

        def work():
            while True:
                frames = self.q.get()
                if frames is None:
                    return
                try:
                    if self.err is None:
                        for f in frames:
                            write_frame(f)
                except BaseException as exc:
                    self.err = exc

        self.t = threading.Thread(target=work, daemon=True)
        self.t.start()

    def put(self, frames):
        if self.err is not None:
            raise self.err
        self.q.put(frames)

    def close(self):
        self.q.put(None)
        self.t.join()
        if self.err is not None:
            raise self.err


def embed_frame_stream(read_frame, write_frame, payload_packed, total_bits, delta, num_ac, out_hw, *,
                       batch_frames=32, embed_fn=None, log=None, overlap_io=True):
    """Embed `total_bits` of `payload_packed` (MSB-first) into the frames `read_frame()` yields.

    read_frame()  -> (ok, frame_bgr) like cv2.VideoCapture.read; frames are cropped to out_hw
    write_frame(a)   receives one (h,w,3) uint8 BGR frame, like cv2.VideoWriter.write
    overlap_io       decode the next batch and encode the previous one on worker threads while the
                     current one is on the GPU (same frames, same order; False = strictly serial)
    Returns (all_embedded, first_gray, first_stego_gray, frames_seen): what the reference's
    loop leaves behind (embed_process.py:147-152).
    """
    h, w = out_hw
    embed_fn = embed_fn or _ops().embed
    log = log or (lambda *_: None)
    total_bits = int(total_bits)
    cap = frame_path.capacity_bits(h, w, num_ac) if delta > 0 else 0
    # how many leading frames go through the hot path; with cap == 0 every frame does and the
    # payload never completes (the reference's behaviour for delta <= 0 or num_ac <= 0)
    n_embed = math.inf if cap == 0 else -(-total_bits // cap)
    seen = 0
    first_gray = first_stego = None
    embedded = 0
    done = False
    crop = lambda frame: frame[0:h, 0:w]
    if overlap_io:
        source, sink = _Prefetch(read_frame, batch_frames, crop), _Writer(write_frame)
        next_batch, emit = source.next, sink.put
    else:
        def next_batch():
            batch = []
            while len(batch) < batch_frames:
                ok, frame = read_frame()
                if not ok:
                    break
                batch.append(crop(frame))
            return batch

        def emit(frames):
            for f in frames:
                write_frame(f)
    try:
        while True:
            batch = next_batch()
            if not batch:
                break
            k = 0 if total_bits == 0 else int(min(len(batch), max(0, n_embed - seen)))
            if k:
                stack = np.stack([np.asarray(f) for f in batch[:k]])
                off = seen * cap
                stego, gray0, nbits = embed_fn(stack, payload_packed, off, total_bits - off, delta, num_ac, seen == 0)
                if seen == 0:
                    first_gray, first_stego = gray0, np.ascontiguousarray(stego[0, :, :, 0])
                # (the stego batch may live in a reused pinned buffer: hand the writer its own copy)
                emit([np.array(stego[i]) for i in range(k)] if overlap_io else [stego[i] for i in range(k)])
                for i in range(k):
                    embedded += int(nbits[i])
                    log("    Frame %d: %d bits disisipkan. Total disisipkan: %d/%d" % (seen + i + 1, int(nbits[i]), embedded, total_bits))
                if embedded >= total_bits:
                    done = True
            rest = []
            for f in batch[k:]:                               # past the payload: copied in colour (:131-140)
                f = np.asarray(f)
                rest.append(f if f.ndim == 3 else np.repeat(f[..., None], 3, 2))
            if rest:
                emit(rest)
            seen += len(batch)
    finally:
        if overlap_io:
            sink.close()
    if not done:
        log("    Warning: Video selesai sebelum semua payload (%d bits) disisipkan." % total_bits)
        return False, None, None, seen
    return True, first_gray, first_stego, seen


# ----------------------------------------------------------------------------------------------
# extract: the read-as-needed loops of extract_process.py:55-86 and :173-182
# ----------------------------------------------------------------------------------------------
class StegoBitReader:
    """The extracted bit stream of a stego video, served in payload order by whole bytes.

    The stream is kept PACKED (one bytearray + a bit cursor), never as one byte or character per
    bit: appending a batch and taking a field both cost time proportional to what is appended /
    taken, not to what has accumulated (the reference re-concatenates an ever growing string,
    extract_process.py:76,181)."""

    def __init__(self, read_frame, delta, num_ac, out_hw, *, batch_frames=8, extract_fn=None, log=None):
        self.read_frame = read_frame
        self.delta, self.num_ac = delta, num_ac
        self.h, self.w = out_hw
        self.batch_frames = max(1, int(batch_frames))
        self.extract_fn = extract_fn or _ops().extract
        self.log = log or (lambda *_: None)
        self.cap = frame_path.capacity_bits(self.h, self.w, num_ac)
        self.buf = bytearray()                                # packed MSB-first; self.end valid bits
        self.end = 0
        self.pos = 0                                          # bit cursor
        self.frames_read = 0
        self.exhausted = False

    def available(self):
        return self.end - self.pos

    def _append_rows(self, packed):
        """Append `cap` bits of every row of `packed` (k, ceil(cap/8)) to the stream."""
        packed = np.asarray(packed, np.uint8)
        if self.cap % 8 == 0 and self.end % 8 == 0:           # every BASELINE geometry: plain byte append
            self.buf += packed[:, :self.cap // 8].tobytes()
            self.end += packed.shape[0] * self.cap
            return
        # general case: merge at bit granularity, touching only the partial tail byte and the new rows
        tail_bits = self.end % 8
        head = np.unpackbits(np.frombuffer(bytes(self.buf[-1:]), np.uint8))[:tail_bits] if tail_bits else np.zeros(0, np.uint8)
        fresh = np.unpackbits(packed, axis=1)[:, :self.cap].reshape(-1)
        if tail_bits:
            del self.buf[-1:]
        self.buf += np.packbits(np.concatenate([head, fresh])).tobytes()
        self.end += fresh.size

    def ensure(self, nbits):
        """Make at least `nbits` unread bits available; False if the video ends first."""
        while self.available() < nbits and not self.exhausted:
            need = nbits - self.available()
            want = self.batch_frames if self.cap == 0 else min(self.batch_frames, -(-need // self.cap))
            frames = []
            while len(frames) < max(1, want):
                ok, frame = self.read_frame()
                if not ok:
                    self.exhausted = True
                    break
                frames.append(np.asarray(frame)[0:self.h, 0:self.w])
            if not frames:
                break
            if self.cap == 0:
                raise ValueError("no bits can be extracted (num_ac <= 0)")
            self._append_rows(self.extract_fn(np.stack(frames), self.delta, self.num_ac))
            self.frames_read += len(frames)
            self.log("    %d frame diekstrak (total %d), %d bit tersedia" % (len(frames), self.frames_read, self.available()))
        return self.available() >= nbits

    def take_bytes(self, n):
        if not self.ensure(8 * n):
            raise EOFError("video ended before %d more bytes could be extracted" % n)
        first, shift = self.pos >> 3, self.pos & 7
        if shift == 0:
            out = bytes(self.buf[first:first + n])
        else:                                                 # unaligned cursor: shift the window only
            window = np.frombuffer(bytes(self.buf[first:first + n + 1]), np.uint8)
            out = np.packbits(np.unpackbits(window)[shift:shift + 8 * n]).tobytes()
        self.pos += 8 * n
        return out

    def take_uint(self, nbytes):
        return int.from_bytes(self.take_bytes(nbytes), "big")


# ----------------------------------------------------------------------------------------------
# payload layout (embed_process.py:62-74 / extract_process.py:89-165): all fields byte aligned
# ----------------------------------------------------------------------------------------------
def build_payload(width, height, eph_pub, salt, digest, nonce, tag, ciphertext):
    """meta(16+16 bit) | len8 pub | len8 salt | len8 sha3 | len8 nonce | len8 tag | len32 ciphertext."""
    out = struct.pack(">HH", width, height)
    for field in (eph_pub, salt, digest, nonce, tag):
        if len(field) > 255:
            raise ValueError("Nilai %d di luar jangkauan untuk 8 bit." % len(field))
        out += bytes([len(field)]) + bytes(field)
    return out + struct.pack(">I", len(ciphertext)) + bytes(ciphertext)


def parse_payload(reader):
    """Inverse of build_payload over a StegoBitReader (frames are pulled only as needed)."""
    width, height = reader.take_uint(2), reader.take_uint(2)
    if width == 0 or height == 0:
        raise ValueError("Error: Metadata gambar 0x0.")
    fields = [reader.take_bytes(reader.take_uint(1)) for _ in range(5)]
    ciphertext = reader.take_bytes(reader.take_uint(4))
    return (width, height, *fields, ciphertext)


# ----------------------------------------------------------------------------------------------
# the reference's two pipeline functions
# ----------------------------------------------------------------------------------------------
def reference_modules():
    """The reference's own crypto / image-codec modules (never re-implemented here)."""
    import importlib
    return types.SimpleNamespace(cs=importlib.import_module("config_and_setup"), helpers=importlib.import_module("helpers"))


def _video_io():
    import cv2
    return cv2


def embed_gambar_ke_video_final(path_video_input, path_gambar_rahasia, path_video_output_base, delta_kuantisasi,
                                num_ac_coeffs, kunci_publik_ecc_penerima_bytes_compressed, *, ref=None,
                                batch_frames=32, embed_fn=None, verbose=True):
    """Same contract as embed_process.py:17-152: -> (ok, first_gray, first_stego_gray)."""
    ref = ref or reference_modules()
    cs, hp = ref.cs, ref.helpers
    say = print if verbose else (lambda *_: None)
    say("\n=== MEMULAI PROSES EMBEDDING GAMBAR KE VIDEO (batched, B200) ===")
    lebar, tinggi, bitstream = hp.gambar_ke_bitstream(path_gambar_rahasia)
    if bitstream is None:
        return False, None, None
    try:
        image_bytes = bitstring_to_bytes(bitstream)
        digest = cs.hitung_sha3_256(image_bytes)
        eph_priv, eph_pub = cs.buat_pasangan_kunci_ecc()
        receiver_pub = cs.deserialisasi_kunci_publik_ecc_compressed(kunci_publik_ecc_penerima_bytes_compressed)
        shared = cs.buat_shared_secret_ecdh(eph_priv, receiver_pub)
        salt = os.urandom(16)
        key = cs.derive_kunci_aes_dari_shared_secret(shared, salt, 32)
        ciphertext, nonce, tag = cs.enkripsi_aes_gcm(image_bytes, key)
        payload = build_payload(lebar, tinggi, cs.serialisasi_kunci_publik_ecc_compressed(eph_pub), salt, digest,
                                nonce, tag, ciphertext)
    except Exception as exc:                       # the reference reports and returns (False, None, None)
        say("    Error: persiapan kriptografi / payload gagal: %s" % exc)
        return False, None, None
    total_bits = 8 * len(payload)
    say("    Total bit payload yang akan disisipkan: %d bits." % total_bits)

    cv2 = _video_io()
    cap = cv2.VideoCapture(path_video_input)
    if not cap.isOpened():
        say("    Error: Video input '%s' tidak bisa dibuka." % path_video_input)
        return False, None, None
    w0, h0 = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
    fps = cap.get(cv2.CAP_PROP_FPS)
    out_w, out_h = (w0 // 8) * 8, (h0 // 8) * 8
    if out_w == 0 or out_h == 0:
        say("    Error: Dimensi video terlalu kecil.")
        cap.release()
        return False, None, None
    out_path = os.path.splitext(path_video_output_base)[0] + ".avi"
    writer = cv2.VideoWriter(out_path, cv2.VideoWriter_fourcc(*"FFV1"), fps, (out_w, out_h), isColor=True)
    if not writer.isOpened():
        say("    ERROR: Gagal VideoWriter FFV1 '%s'." % out_path)
        cap.release()
        return False, None, None
    try:
        ok, gray, stego, _ = embed_frame_stream(cap.read, writer.write, np.frombuffer(payload, np.uint8), total_bits,
                                                delta_kuantisasi, num_ac_coeffs, (out_h, out_w),
                                                batch_frames=batch_frames, embed_fn=embed_fn, log=say)
    finally:
        cap.release()
        writer.release()
    say("  Proses embedding selesai. Video output: '%s'." % out_path if ok else
        "  Proses embedding selesai, namun TIDAK semua data berhasil disisipkan.")
    return (True, gray, stego) if ok else (False, None, None)


def ekstraksi_gambar_video_final(path_stego_video, path_gambar_output, delta_kuantisasi, num_ac_coeffs,
                                 kunci_privat_ecc_penerima, bits_untuk_dimensi=16, *, ref=None, batch_frames=8,
                                 extract_fn=None, verbose=True):
    """Same contract as extract_process.py:22-216: -> bool."""
    if bits_untuk_dimensi != 16:
        raise ValueError("the payload layout of the reference stores the image size as 2 x 16 bits")
    ref = ref or reference_modules()
    cs, hp = ref.cs, ref.helpers
    say = print if verbose else (lambda *_: None)
    say("\n=== MEMULAI PROSES EKSTRAKSI GAMBAR DARI VIDEO (batched, B200) ===")
    cv2 = _video_io()
    cap = cv2.VideoCapture(path_stego_video)
    if not cap.isOpened():
        say("  Error: Tidak bisa membuka stego-video '%s'." % path_stego_video)
        return False
    try:
        w0, h0 = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
        out_w, out_h = (w0 // 8) * 8, (h0 // 8) * 8
        if out_w == 0 or out_h == 0:
            say("  Error: Dimensi video terlalu kecil.")
            return False
        reader = StegoBitReader(cap.read, delta_kuantisasi, num_ac_coeffs, (out_h, out_w), batch_frames=batch_frames,
                                extract_fn=extract_fn, log=say)
        try:
            lebar, tinggi, eph_pub, salt, digest, nonce, tag, ciphertext = parse_payload(reader)
        except (EOFError, ValueError) as exc:
            say("  Ekstraksi GAGAL: %s" % exc)
            return False
    finally:
        cap.release()
    try:
        shared = cs.buat_shared_secret_ecdh(kunci_privat_ecc_penerima, cs.deserialisasi_kunci_publik_ecc_compressed(eph_pub))
        key = cs.derive_kunci_aes_dari_shared_secret(shared, salt, 32)
    except Exception as exc:
        say("  Error saat ECDH atau derivasi kunci AES penerima: %s" % exc)
        return False
    plain = cs.dekripsi_aes_gcm(ciphertext, key, nonce, tag)
    if plain is None:
        say("    Dekripsi GAGAL.")
        return False
    say("    Verifikasi Hash SHA3-256 %s" % ("BERHASIL: Gambar tidak korup." if cs.hitung_sha3_256(plain) == digest
                                           else "GAGAL: Gambar mungkin korup atau telah diubah!"))
    image = hp.bitstream_ke_gambar(bytes_to_bitstring(plain), lebar, tinggi)
    if not image:
        say("  Gagal merekonstruksi gambar.")
        return False
    try:
        image.save(path_gambar_output)
    except Exception as exc:
        say("  Error simpan gambar: %s" % exc)
        return False
    say("    Gambar berhasil diekstrak dan disimpan sebagai '%s'." % path_gambar_output)
    return True


def install_pipelines(modules=None):
    """Rebind the reference's pipeline functions (embed_process.embed_gambar_ke_video_final,
    extract_process.ekstraksi_gambar_video_final) to the batched versions; returns what was patched."""
    import sys
    done = []
    mods = modules or [sys.modules.get("embed_process"), sys.modules.get("extract_process")]
    for m in mods:
        if m is None:
            continue
        if hasattr(m, "embed_gambar_ke_video_final"):
            m.embed_gambar_ke_video_final = embed_gambar_ke_video_final
            done.append(m.__name__ + ".embed_gambar_ke_video_final")
        if hasattr(m, "ekstraksi_gambar_video_final"):
            m.ekstraksi_gambar_video_final = ekstraksi_gambar_video_final
            done.append(m.__name__ + ".ekstraksi_gambar_video_final")
    return done
