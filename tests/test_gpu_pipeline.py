"""GPU tests of the batched pipeline driver (N1): the same frame streams through the CUDA kernels
and through the oracle must be identical, and the reference-built ECDH/AES-GCM/SHA3 payload must
survive embed -> (frames) -> read-as-needed extract -> parse -> decrypt.  Nothing here reads
/root/reference."""
import numpy as np
import pytest

import svs_b200
from svs_b200 import pipeline
from tests import golden_util as G
from tests.synth import synth_frames, synth_bits
from tests.test_pipeline_host import Feed, oracle_embed, oracle_extract

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,n,delta,frac,batch", [
    ((7, 40, 56, 3), 17, 12, 2.4, 3),
    ((6, 480, 640, 3), 10, 20, 3.2, 4),
    ((3, 720, 1280, 3), 63, 20, 1.5, 8),
])
def test_embed_stream_gpu_equals_oracle(shape, n, delta, frac, batch):
    frames = synth_frames("gpipe%s" % (shape,), shape, 30, 230)
    h, w = shape[1:3]
    cap = svs_b200.capacity_bits(h, w, n)
    bits = synth_bits("gpipe", int(frac * cap) + 11)
    out_gpu, out_cpu = [], []
    before = svs_b200.lib().svs_kernel_launch_count()
    r_gpu = pipeline.embed_frame_stream(Feed(frames).read, lambda a: out_gpu.append(np.array(a)), np.packbits(bits),
                                        bits.size, delta, n, (h, w), batch_frames=batch)
    assert svs_b200.lib().svs_kernel_launch_count() > before
    r_cpu = pipeline.embed_frame_stream(Feed(frames).read, lambda a: out_cpu.append(np.array(a)), np.packbits(bits),
                                        bits.size, delta, n, (h, w), batch_frames=batch, embed_fn=oracle_embed)
    assert r_gpu[0] is True and r_cpu[0] is True and r_gpu[3] == r_cpu[3] == shape[0]
    assert np.array_equal(r_gpu[1], r_cpu[1]) and np.array_equal(r_gpu[2], r_cpu[2])
    assert len(out_gpu) == len(out_cpu) == shape[0]
    for a, b in zip(out_gpu, out_cpu):
        assert np.array_equal(a, b)
    # and the stream reads back, batch by batch, to the payload the oracle reader sees
    rd_gpu = pipeline.StegoBitReader(Feed(out_gpu).read, delta, n, (h, w), batch_frames=2)
    rd_cpu = pipeline.StegoBitReader(Feed(out_cpu).read, delta, n, (h, w), batch_frames=2, extract_fn=oracle_extract)
    nbytes = bits.size // 8
    assert rd_gpu.take_bytes(nbytes) == rd_cpu.take_bytes(nbytes)


def test_reference_payload_survives_the_batched_round_trip():
    pytest.importorskip("cryptography")
    from cryptography.hazmat.primitives import hashes, serialization
    from cryptography.hazmat.primitives.asymmetric import ec
    from cryptography.hazmat.primitives.ciphers.aead import AESGCM
    from cryptography.hazmat.primitives.kdf.hkdf import HKDF

    e = G.load_e2e()                                            # payload built by the reference's own helpers
    h, w, n, delta = 240, 320, 10, 20                            # 12,000 bits per frame -> 3 frames for 33,744 bits
    frames = synth_frames("gpipe-e2e", (5, h, w, 3), 64, 192)
    written = []
    ok, _, _, seen = pipeline.embed_frame_stream(Feed(frames).read, lambda a: written.append(np.array(a)),
                                                 e["payload_packed"], e["total_bits"], delta, n, (h, w), batch_frames=2)
    assert ok and seen == 5 and all(np.array_equal(written[i], frames[i]) for i in (3, 4))
    feed = Feed(written)
    reader = pipeline.StegoBitReader(feed.read, delta, n, (h, w), batch_frames=2)
    width, height, eph_pub, salt, digest, nonce, tag, ct = pipeline.parse_payload(reader)
    assert feed.i == 3                                          # only the frames that carry payload were read
    priv = serialization.load_pem_private_key(e["pem"], password=None)
    shared = priv.exchange(ec.ECDH(), ec.EllipticCurvePublicKey.from_encoded_point(ec.SECP256R1(), eph_pub))
    key = HKDF(algorithm=hashes.SHA256(), length=32, salt=salt, info=b'kunci aes untuk steganografi video').derive(shared)
    plain = AESGCM(key).decrypt(nonce, ct + tag, None)
    h3 = hashes.Hash(hashes.SHA3_256())
    h3.update(plain)
    assert h3.finalize() == digest == e["sha3"]
    assert np.array_equal(np.frombuffer(plain, np.uint8).reshape(height, width), e["image"])
