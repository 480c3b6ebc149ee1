#!/usr/bin/env python
"""Can a COPY ENGINE write to an NVSwitch multicast address?  (torchrun, >= 2 ranks)

Each rank owns rows [r*N, (r+1)*N) of a symmetric buffer.  It copies its rows from a private
staging buffer to (multicast pointer + its row offset) with ONE cudaMemcpyAsync - if the copy
engines can address the multicast mapping, the switch replicates the rows into every rank's
buffer (all-gather with 1/world of the outbound traffic and no SMs).  Verifies the result on
every rank and times it next to per-peer copies."""
import ctypes
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_bytes = int(os.environ.get("CE_MC_BYTES", str(460 << 20))) // 256 * 256
    g = symm.empty((world, n_bytes), dtype=torch.uint8, device=dev)
    hdl = symm.rendezvous(g, dist.group.WORLD)
    mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
    if rank == 0:
        print("world %d, %d MB per rank, multicast_ptr %s" % (world, n_bytes >> 20, hex(mc)), flush=True)
    if not mc:
        if rank == 0:
            print("no multicast mapping on this fabric")
        dist.destroy_process_group()
        return
    cudart = ctypes.CDLL("libcudart.so.12")
    cudart.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    cudart.cudaMemcpyAsync.restype = ctypes.c_int
    stage = torch.full((n_bytes,), rank + 1, dtype=torch.uint8, device=dev)
    stage[::4097] = 200 + rank
    st = torch.cuda.Stream(dev)
    g.zero_()
    torch.cuda.synchronize()
    hdl.barrier(channel=0)

    def push_multicast():
        return cudart.cudaMemcpyAsync(ctypes.c_void_p(mc + rank * n_bytes), ctypes.c_void_p(stage.data_ptr()), n_bytes, 3,
                                      ctypes.c_void_p(st.cuda_stream))

    rc = push_multicast()
    st.synchronize()
    err = torch.cuda.current_stream().query()
    torch.cuda.synchronize()
    hdl.barrier(channel=1)
    torch.cuda.synchronize()
    ok = rc == 0 and all(bool(torch.equal(g[r], torch.full_like(stage, r + 1).index_put_((torch.arange(0, n_bytes, 4097, device=dev),), torch.tensor(200 + r, dtype=torch.uint8, device=dev)))) for r in range(world))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("cudaMemcpyAsync to the multicast address: rc %d, every rank holds every rank's rows: %s" % (rc, bool(flag.item())), flush=True)
    if flag.item():
        for name, fn in (("one copy to the multicast address", push_multicast),):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                fn()
            st.synchronize()
            dist.barrier()
            dt = (time.perf_counter() - t0) / 5
            if rank == 0:
                print("%-40s %.2f ms per all-gather of %d x %d MB (%.0f GB/s into every GPU)" % (name, dt * 1e3, world, n_bytes >> 20, world * n_bytes / dt / 1e9), flush=True)
    # reference: per-peer copy-engine pushes (what CopyEngineGather does)
    peers = [hdl.get_buffer(r, (world, n_bytes), torch.uint8)[rank] for r in range(world) if r != rank]
    streams = [torch.cuda.Stream(dev) for _ in range(min(4, len(peers)))]
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        for i, p in enumerate(peers):
            with torch.cuda.stream(streams[i % len(streams)]):
                p.copy_(stage, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    dt = (time.perf_counter() - t0) / 5
    if rank == 0:
        print("%-40s %.2f ms per all-gather (%.0f GB/s into every GPU)" % ("one copy per peer (copy engines)", dt * 1e3, (world - 1) * n_bytes / dt / 1e9), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
