"""2-GPU NCCL test of the frame-sharded path (skipped with fewer than 2 devices): each rank embeds
its frame range with the CUDA kernels, the overlapped extract + all-gather returns the whole
bitstream in frame order on both ranks, and everything equals the single-process oracle."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

H, W, N_AC, DELTA, F_TOTAL = 64, 256, 63, 20, 12


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    import torch
    import torch.distributed as dist
    import svs_b200
    from svs_b200 import sharding
    from tests.synth import synth_frames, synth_bits
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        frames = synth_frames("mgpu", (F_TOTAL, H, W, 3), 64, 192)
        cap = svs_b200.capacity_bits(H, W, N_AC)
        total = F_TOTAL * cap
        packed = torch.from_numpy(np.packbits(synth_bits("mgpu", total))).cuda()
        f0, f1 = sharding.frame_range(F_TOTAL, rank, world)
        mine = torch.from_numpy(frames[f0:f1]).cuda()
        res = sharding.embed_shard(mine, packed, total, DELTA, N_AC, n_frames_total=F_TOTAL)
        og = sharding.OverlappedExtractGather(f1 - f0, svs_b200.bits_row_bytes(H, W, N_AC), mine.device, chunks=3)
        og.run(res.stego, DELTA, N_AC)
        full = og.wait()[:, :(cap + 7) // 8]
        plain = sharding.extract_allgather(res.stego, DELTA, N_AC, n_frames_total=F_TOTAL)
        torch.cuda.synchronize()
        assert torch.equal(full, plain)
        # the fused path: the extract kernel stores into both ranks' symmetric buffers (peer stores, then
        # the NVSwitch multicast address when the fabric has one); twice, to cover buffer reuse
        for use_mc in (False, True):
            fg = sharding.FusedExtractGather(f1 - f0, svs_b200.bits_row_bytes(H, W, N_AC), mine.device, use_multicast=use_mc)
            for _ in range(2):
                fg.gathered.zero_()
                fg.hdl.barrier(channel=2)
                got = fg.run(res.stego, DELTA, N_AC)[:, :(cap + 7) // 8]
                torch.cuda.synchronize()
                assert torch.equal(got, plain), "fused gather (%s) differs" % fg.mode
        # ... and the copy-engine variants (DMA pushes on side streams, overlappable with the next batch):
        # one copy per peer, and ONE copy to the NVSwitch multicast address of this rank's rows
        for mc in (False, "force"):
            cg = sharding.CopyEngineGather(f1 - f0, svs_b200.bits_row_bytes(H, W, N_AC), mine.device, n_streams=2, use_multicast=mc)
            for _ in range(3):                               # three rounds over two buffers: covers buffer reuse
                cg.gathered.zero_()
                cg.hdl.barrier(channel=2)
                cg.run(res.stego, DELTA, N_AC)
                got = cg.wait()[:, :(cap + 7) // 8]
                torch.cuda.synchronize()
                assert torch.equal(got, plain), "copy-engine gather differs (%s)" % cg.mode
        np.save(os.path.join(outdir, "bits%d.npy" % rank), full.cpu().numpy())
        np.save(os.path.join(outdir, "stego%d.npy" % rank), res.stego.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_two_gpu_shard_embed_and_overlapped_gather(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import c_oracle as oc
    from tests.synth import synth_frames, synth_bits
    import svs_b200
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    frames = synth_frames("mgpu", (F_TOTAL, H, W, 3), 64, 192)
    cap = svs_b200.capacity_bits(H, W, N_AC)
    bits = synth_bits("mgpu", F_TOTAL * cap)
    stego, _, _ = oc.embed_frames(frames, np.packbits(bits), bits.size, DELTA, N_AC, threads=4)
    want = oc.extract_frames(stego, DELTA, N_AC, threads=4)
    assert np.array_equal(np.concatenate([np.load(tmp_path / "stego0.npy"), np.load(tmp_path / "stego1.npy")]), stego)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / ("bits%d.npy" % r)), want)
    assert np.array_equal(np.unpackbits(want, axis=1)[:, :cap].reshape(-1), bits)
