"""world_size-2 gloo tests (CPU) of the frame-sharding host logic: partition, payload slices,
all-gather reassembly in frame order, ragged shards.  The per-rank compute is a CPU stand-in
(the C oracle) injected through embed_fn / extract_fn; the GPU path is covered by -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import svs_b200
from svs_b200 import sharding
from oracle import c_oracle as oc
from tests.synth import synth_frames, synth_bits

H, W, N_AC, DELTA = 32, 48, 63, 20


def _cpu_embed(frames, payload, total_bits, delta, num_ac, *, bit_offset=0, **kw):
    stego, gray, nb = oc.embed_frames(frames.numpy(), payload.numpy(), total_bits, delta, num_ac, bit_offset=bit_offset)
    return svs_b200.EmbedResult(torch.from_numpy(stego), torch.from_numpy(gray), torch.from_numpy(nb), None)


def _cpu_extract(frames, delta, num_ac, out=None):
    return torch.from_numpy(oc.extract_frames(frames.numpy(), delta, num_ac))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, total_bits, global_payload, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames = torch.from_numpy(synth_frames("shard", (n_frames, H, W, 3), 64, 192))
        bits = synth_bits("shard", total_bits)
        packed = torch.from_numpy(np.packbits(bits))
        cap = svs_b200.capacity_bits(H, W, N_AC)
        f0, f1 = sharding.frame_range(n_frames, rank, world)
        if global_payload:
            res = sharding.embed_shard(frames[f0:f1], packed, total_bits, DELTA, N_AC, n_frames_total=n_frames,
                                       embed_fn=_cpu_embed)
        else:
            off, nb = sharding.payload_slice(total_bits, cap, f0, f1)
            assert off % 8 == 0
            mine = packed[off // 8:(off + nb + 7) // 8]
            res = sharding.embed_shard(frames[f0:f1], mine, nb, DELTA, N_AC, n_frames_total=n_frames,
                                       payload_is_global=False, embed_fn=_cpu_embed)
        full = sharding.extract_allgather(res.stego, DELTA, N_AC, n_frames_total=n_frames, extract_fn=_cpu_extract)
        np.save(os.path.join(result_dir, "bits%d.npy" % rank), full.numpy())
        np.save(os.path.join(result_dir, "stego%d.npy" % rank), res.stego.numpy())
        np.save(os.path.join(result_dir, "nb%d.npy" % rank), res.bits_embedded.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames,global_payload,fill", [(4, True, 1.0), (5, True, 0.7), (5, False, 1.0), (3, True, 0.2)])
def test_two_rank_shard_and_gather(tmp_path, n_frames, global_payload, fill):
    world = 2
    cap = svs_b200.capacity_bits(H, W, N_AC)
    total_bits = int(n_frames * cap * fill)
    mp.spawn(_worker, args=(world, _free_port(), n_frames, total_bits, global_payload, str(tmp_path)),
             nprocs=world, join=True)
    frames = synth_frames("shard", (n_frames, H, W, 3), 64, 192)
    bits = synth_bits("shard", total_bits)
    stego, _, nb = oc.embed_frames(frames, np.packbits(bits), total_bits, DELTA, N_AC)
    want_bits = oc.extract_frames(stego, DELTA, N_AC)
    got0 = np.load(tmp_path / "bits0.npy")
    got1 = np.load(tmp_path / "bits1.npy")
    assert np.array_equal(got0, want_bits) and np.array_equal(got1, want_bits)     # same on every rank, frame order
    st = np.concatenate([np.load(tmp_path / "stego0.npy"), np.load(tmp_path / "stego1.npy")])
    assert np.array_equal(st, stego)                                               # shards == single-process result
    assert np.concatenate([np.load(tmp_path / "nb0.npy"), np.load(tmp_path / "nb1.npy")]).tolist() == nb.tolist()
    # the gathered stream starts with the payload (mid-range frames: error-free round trip)
    stream = np.unpackbits(want_bits, axis=1)[:, :cap].reshape(-1)
    assert np.array_equal(stream[:total_bits], bits)


def test_partition_helpers():
    assert [sharding.frame_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sharding.frame_counts(1800, 8) == [225] * 8
    assert sharding.frame_counts(3, 4) == [1, 1, 1, 0]
    assert sharding.payload_slice(1000, 300, 0, 2) == (0, 600)
    assert sharding.payload_slice(1000, 300, 2, 4) == (600, 400)
    assert sharding.payload_slice(1000, 300, 4, 6) == (1000, 0)
