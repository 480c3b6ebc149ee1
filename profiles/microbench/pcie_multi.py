#!/usr/bin/env python
"""Concurrent pinned host<->device copies on N GPUs of one box: what the host side of the PCIe
links can deliver in aggregate.  This is the ceiling of bench.py's `e2e` at N > 1 (its timed region
copies 5.1 GB in and 1.4 GB out per rank and step), next to the single-GPU link rate of pcie.py.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 profiles/microbench/pcie_multi.py          (one line per mode, rank 0)

Every rank copies a 1 GiB pinned buffer 6 times: H2D alone, D2H alone, and both directions at
once on two streams; all ranks start together (barrier) and the slowest rank sets the time."""
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 1 << 30
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(rank + 1)
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.full((n,), 7, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    reps = 6

    def run(h2d, d2h):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.item())

    run(True, True)
    for name, h2d, d2h in (("H2D alone", True, False), ("D2H alone", False, True), ("H2D + D2H together", True, True)):
        dt = run(h2d, d2h)
        per_dir = n * reps / dt / 1e9
        if rank == 0:
            print("%d GPUs  %-20s per GPU: %s  aggregate: %s" % (
                world, name,
                " + ".join("%.1f GB/s %s" % (per_dir, k) for k, on in (("H2D", h2d), ("D2H", d2h)) if on),
                " + ".join("%.1f GB/s %s" % (per_dir * world, k) for k, on in (("H2D", h2d), ("D2H", d2h)) if on)), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
