#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (it needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports ``proses_frame_qim_dct`` from /root/reference/config_and_setup.py and the
reference's crypto/payload helpers, feeds them the deterministic inputs of tests/synth.py and
stores what they return:

  * cases_small.npz / cases.json - per-frame embed + extract results (full arrays for small
    frames, SHA-256 digests for the 640x480 and 1080p frames);
  * e2e_payload.npz + e2e_receiver_private.pem - a complete ECDH/HKDF/AES-GCM/SHA3 payload for
    media/input/image64.png built by the reference's own functions (embed_process.py:24-74), so
    that the GPU tests can check "decrypts, tag verifies, SHA3 matches, pixels identical"
    without the reference being present.

Nothing here is product code; the fixtures are what pins oracle/ (SURVEY.md section 8c).
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from tests.synth import synth_frames, synth_bits, bits_to_str, gradient_frame  # noqa: E402

import config_and_setup as ref  # noqa: E402  (the reference, read-only)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# name, frame builder, delta, num_ac, payload bits (None -> fill capacity + 100), store arrays?
def build_cases():
    cases = []

    def add(name, frame, delta, num_ac, nbits=None, full=True):
        cases.append(dict(name=name, frame=frame, delta=delta, num_ac=num_ac, nbits=nbits, full=full))

    add("bgr_48x64_d20_ac10", synth_frames("a", (48, 64, 3)), 20, 10)
    add("bgr_48x64_d20_ac63", synth_frames("b", (48, 64, 3)), 20, 63)
    add("gray_40x56_d7_ac63_midblock", synth_frames("c", (40, 56)), 7, 63, nbits=1234)
    add("bgr_32x32_d3_ac5_short", synth_frames("d", (32, 32, 3)), 3, 5, nbits=77)
    add("bgr_32x40_d1_ac63", synth_frames("e", (32, 40, 3)), 1, 63)
    add("bgr_32x40_d100_ac63", synth_frames("f", (32, 40, 3)), 100, 63)
    add("bgr_24x24_d2p5_ac63", synth_frames("g", (24, 24, 3)), 2.5, 63)
    add("gray_24x24_d0p1_ac100", synth_frames("h", (24, 24)), 0.1, 100)
    add("bgr_16x24_d0_ac10", synth_frames("i", (16, 24, 3)), 0, 10, nbits=50)
    add("bgr_16x24_d20_ac0", synth_frames("j", (16, 24, 3)), 20, 0, nbits=50)
    add("bgr_16x16_d20_ac63_empty", synth_frames("k", (16, 16, 3)), 20, 63, nbits=0)
    add("black_16x32_d20_ac63", np.zeros((16, 32, 3), np.uint8), 20, 63)
    add("white_16x32_d20_ac63", np.full((16, 32, 3), 255, np.uint8), 20, 63)
    chk = np.zeros((16, 32), np.uint8)
    chk[::2, 1::2] = 255
    chk[1::2, ::2] = 255
    add("checker_16x32_d20_ac63", chk, 20, 63)
    add("midrange_32x48_d20_ac63", synth_frames("m", (32, 48, 3), 64, 192), 20, 63)
    add("gradient_48x64_d8_ac32", gradient_frame(48, 64, 3, seed=1), 8, 32)
    add("bgr_40x40_d20_ac1", synth_frames("n", (40, 40, 3)), 20, 1)
    add("bgr_40x40_d6_ac17_mid", synth_frames("o", (40, 40, 3)), 6, 17, nbits=200)
    try:
        import cv2
        cap = cv2.VideoCapture("/root/reference/media/input/cover_1.mp4")
        ok, fr = cap.read()
        cap.release()
        if ok:
            add("cover1_crop_64x96_d20_ac10", np.ascontiguousarray(fr[200:264, 400:496]), 20, 10)
            add("cover1_crop_64x96_d7_ac63", np.ascontiguousarray(fr[300:364, 600:696]), 7, 63)
    except Exception as e:  # pragma: no cover
        print("cover crop skipped:", e)
    # digests only
    add("bgr_480x640_d20_ac10_cfg2", synth_frames("cfg2", (480, 640, 3)), 20, 10, nbits=33744, full=False)
    add("bgr_480x640_d20_ac63", synth_frames("cfg2b", (480, 640, 3)), 20, 63, full=False)
    add("bgr_1080p_d20_ac63_fullrange", synth_frames("cfg3", (1080, 1920, 3)), 20, 63, full=False)
    add("bgr_1080p_d20_ac10_midrange", synth_frames("cfg3m", (1080, 1920, 3), 64, 192), 20, 10, full=False)
    # round 2: BASELINE config 4 (3840x2160); generated with `make_golden.py --append <names>`
    add("bgr_4k_d20_ac63_midrange", synth_frames("cfg4", (2160, 3840, 3), 64, 192), 20, 63, full=False)
    add("bgr_4k_d20_ac10_fullrange", synth_frames("cfg4f", (2160, 3840, 3)), 20, 10, full=False)
    return cases


def run_case(c):
    frame, delta, num_ac = c["frame"], c["delta"], c["num_ac"]
    h, w = frame.shape[:2]
    cap = (h // 8) * (w // 8) * max(0, min(num_ac, 63))
    nbits = cap + 100 if c["nbits"] is None else c["nbits"]
    bits = synth_bits(c["name"], nbits)
    seg = bits_to_str(bits)
    t0 = time.time()
    gray, stego, n_emb = ref.proses_frame_qim_dct(frame, 'embed', delta, seg, num_ac_coeffs_to_use=num_ac)
    ext_stego = ref.proses_frame_qim_dct(stego, 'extract', delta, num_ac_coeffs_to_use=num_ac)
    ext_input = ref.proses_frame_qim_dct(frame, 'extract', delta, num_ac_coeffs_to_use=num_ac)
    dt = time.time() - t0
    es = np.frombuffer(ext_stego.encode(), np.uint8) - 48
    ei = np.frombuffer(ext_input.encode(), np.uint8) - 48
    meta = dict(name=c["name"], shape=list(frame.shape), delta=delta, num_ac=num_ac, nbits=int(nbits),
                bits_embedded=int(n_emb), n_extracted=len(ext_stego), full=c["full"],
                sha_frame=sha(frame), sha_gray=sha(gray), sha_stego=sha(stego),
                sha_ext_stego=sha(es), sha_ext_input=sha(ei), ref_seconds=round(dt, 3))
    arrays = {}
    if c["full"]:
        p = c["name"]
        arrays = {p + "/frame": frame, p + "/gray": gray, p + "/stego": stego,
                  p + "/ext_stego": np.packbits(es), p + "/ext_input": np.packbits(ei)}
    return meta, arrays


def make_e2e():
    """Payload for image64.png via the reference's own helpers (embed_process.py:24-74)."""
    import helpers as steg_helpers
    from cryptography.hazmat.primitives import serialization
    from cryptography.hazmat.primitives.asymmetric import ec
    w, h, img_bits = steg_helpers.gambar_ke_bitstream("/root/reference/media/input/image64.png")
    img_bytes = ref.bitstream_ke_bytes(img_bits)
    digest = ref.hitung_sha3_256(img_bytes)
    recv_priv = ec.generate_private_key(ec.SECP256R1())
    recv_pub_comp = ref.serialisasi_kunci_publik_ecc_compressed(recv_priv.public_key())
    eph_priv, eph_pub = ref.buat_pasangan_kunci_ecc()
    shared = ref.buat_shared_secret_ecdh(eph_priv, ref.deserialisasi_kunci_publik_ecc_compressed(recv_pub_comp))
    salt = os.urandom(16)
    key = ref.derive_kunci_aes_dari_shared_secret(shared, salt, 32)
    eph_pub_bytes = ref.serialisasi_kunci_publik_ecc_compressed(eph_pub)
    ct, nonce, tag = ref.enkripsi_aes_gcm(img_bytes, key)
    b = ref.bytes_ke_bitstream
    i2b = ref.int_ke_bitstream
    payload = (steg_helpers.buat_metadata_bitstream(w, h)
               + i2b(len(eph_pub_bytes), 8) + b(eph_pub_bytes)
               + i2b(len(salt), 8) + b(salt)
               + i2b(len(digest), 8) + b(digest)
               + i2b(len(nonce), 8) + b(nonce)
               + i2b(len(tag), 8) + b(tag)
               + i2b(len(ct), 32) + b(ct))
    bits = np.frombuffer(payload.encode(), np.uint8) - 48
    pem = recv_priv.private_bytes(serialization.Encoding.PEM, serialization.PrivateFormat.PKCS8,
                                  serialization.NoEncryption())
    with open(os.path.join(HERE, "e2e_receiver_private.pem"), "wb") as f:
        f.write(pem)
    np.savez_compressed(os.path.join(HERE, "e2e_payload.npz"),
                        payload_packed=np.packbits(bits), total_bits=np.int64(bits.size),
                        image=np.frombuffer(img_bytes, np.uint8).reshape(h, w),
                        sha3=np.frombuffer(digest, np.uint8), width=np.int64(w), height=np.int64(h))
    print("e2e payload: %d bits for %dx%d image" % (bits.size, w, h))


def append(names):
    """Run only the named digest-only cases and merge them into cases.json (the existing fixtures -
    in particular the randomly keyed e2e payload - stay untouched)."""
    with open(os.path.join(HERE, "cases.json")) as f:
        doc = json.load(f)
    have = {c["name"] for c in doc["cases"]}
    for c in build_cases():
        if c["name"] in names and c["name"] not in have:
            assert not c["full"], "--append is for digest-only cases"
            m, _ = run_case(c)
            doc["cases"].append(m)
            print("%-36s emb=%-8d ext=%-8d %.2fs" % (m["name"], m["bits_embedded"], m["n_extracted"], m["ref_seconds"]))
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(doc, f, indent=1)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--append":
        return append(set(sys.argv[2:]))
    metas, arrays = [], {}
    for c in build_cases():
        m, a = run_case(c)
        metas.append(m)
        arrays.update(a)
        print("%-36s emb=%-8d ext=%-8d %.2fs" % (m["name"], m["bits_embedded"], m["n_extracted"], m["ref_seconds"]))
    np.savez_compressed(os.path.join(HERE, "cases_small.npz"), **arrays)
    import scipy, cv2
    env = dict(numpy=np.__version__, scipy=scipy.__version__, opencv=cv2.__version__,
               python=sys.version.split()[0],
               reference="erc-a/Secure-Video-Steganography-using-ECC-and-DCT config_and_setup.py:106-174")
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(dict(env=env, cases=metas), f, indent=1)
    make_e2e()


if __name__ == "__main__":
    main()
