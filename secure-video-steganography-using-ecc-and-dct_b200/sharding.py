"""Frame-sharded multi-GPU driver: one process per GPU, torch.distributed for the plumbing.

The path shards by frame: frame f only needs payload bits
[f*cap, (f+1)*cap), so rank r owns a contiguous frame range and the matching payload slice and
embeds with no communication.  Extraction has one real exchange step: every rank ends up with
the whole bitstream in frame (= payload) order, an all-gather of fixed-size per-frame packed
bit rows - NCCL over NVLink on GPUs, gloo in the CPU tests of this host logic.

The compute callables default to the CUDA kernels; tests inject CPU stand-ins to exercise the
partitioning / reassembly logic with world_size 2 on gloo.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import frame_path


def frame_range(n_frames, rank, world):
    """Contiguous, balanced partition: the first n_frames % world ranks get one extra frame."""
    base, rem = divmod(int(n_frames), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def frame_counts(n_frames, world):
    return [frame_range(n_frames, r, world)[1] - frame_range(n_frames, r, world)[0] for r in range(world)]


def payload_slice(total_bits, cap, f0, f1):
    """(bit_offset, nbits) of the payload that frames [f0, f1) consume (embed_process.py:115-128)."""
    total_bits, cap = int(total_bits), int(cap)
    start = min(total_bits, f0 * cap)
    stop = min(total_bits, f1 * cap)
    return start, stop - start


def embed_shard(frames_local, payload, total_bits, delta, num_ac, *, n_frames_total, rank=None, world=None,
                payload_is_global=True, embed_fn=None, **kw):
    """Embed this rank's frames.  `payload` is either the whole packed payload (every rank holds
    it; the kernel is pointed at bit f0*cap) or just this rank's slice (payload_is_global=False,
    which must start at a byte boundary, true whenever cap % 8 == 0)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    embed_fn = embed_fn or frame_path.embed_frames
    h, w = frames_local.shape[1:3]
    cap = frame_path.capacity_bits(h, w, num_ac) if delta > 0 else 0
    f0, f1 = frame_range(n_frames_total, rank, world)
    if f1 - f0 != frames_local.shape[0]:
        raise ValueError("rank %d owns frames [%d,%d) but was handed %d" % (rank, f0, f1, frames_local.shape[0]))
    if payload_is_global:
        off, nbits = payload_slice(total_bits, cap, f0, n_frames_total)     # everything from f0 on
        if cap == 0:
            off, nbits = 0, int(total_bits)
        return embed_fn(frames_local, payload, nbits, delta, num_ac, bit_offset=off, **kw)
    return embed_fn(frames_local, payload, int(total_bits), delta, num_ac, bit_offset=0, **kw)


def all_gather_bits(local_bits, out=None, counts=None, group=None):
    """All-gather per-frame packed bit rows -> (sum(counts), row_bytes) in frame order on every rank.

    Equal shard sizes take one all_gather_into_tensor straight into `out`; ragged shards are
    padded to the largest count and trimmed after the exchange.
    """
    world = dist.get_world_size(group)
    local_bits = local_bits if local_bits.is_contiguous() else local_bits.contiguous()
    nloc, row = local_bits.shape
    if counts is None:
        counts = [nloc] * world
    if len(set(counts)) == 1:
        if out is None:
            out = torch.empty((world * nloc, row), dtype=local_bits.dtype, device=local_bits.device)
        dist.all_gather_into_tensor(out, local_bits, group=group)
        return out
    big = max(counts)
    send = local_bits
    if nloc < big:
        send = torch.zeros((big, row), dtype=local_bits.dtype, device=local_bits.device)
        send[:nloc] = local_bits
    recv = torch.empty((world * big, row), dtype=local_bits.dtype, device=local_bits.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    parts = [recv[r * big:r * big + c] for r, c in enumerate(counts)]
    res = torch.cat(parts, 0)
    if out is not None:
        out.copy_(res)
        return out
    return res


def extract_allgather(frames_local, delta, num_ac, *, n_frames_total=None, group=None, extract_fn=None,
                      local_out=None, out=None):
    """Extract this rank's frames and all-gather: returns (F_total, ceil(cap/8)) packed bits in
    frame order, identical on every rank."""
    world = dist.get_world_size(group)
    extract_fn = extract_fn or frame_path.extract_frames
    local = extract_fn(frames_local, delta, num_ac) if local_out is None else extract_fn(frames_local, delta, num_ac, out=local_out)
    nbytes = local.shape[1]
    counts = None
    if n_frames_total is not None:
        counts = frame_counts(n_frames_total, world)
    # gather the padded rows if the extract buffer has a pitch (avoids a compaction copy)
    base = local
    if local_out is not None and local_out.shape[1] != nbytes:
        base = local_out
    elif not local.is_contiguous() and getattr(local, "_base", None) is not None \
            and local._base.dim() == 2 and local._base.shape[0] == local.shape[0]:
        base = local._base
    full = all_gather_bits(base, out=out, counts=counts, group=group)
    return full[:, :nbytes]


class OverlappedExtractGather:
    """Extract this rank's frames in chunks and all-gather each chunk on a side stream while the
    next chunk is being extracted (and, across calls, while the next batch is being embedded).

    The gathered stream is 2.5 % of the bytes the kernels touch, but every rank has to RECEIVE
    (R-1)/R of it: at 8 ranks that is several milliseconds per 1800-frame batch, the same order
    as the extract kernel itself, so it must not sit on the critical path.  Result: `gathered`,
    (world * F_local, pitch) packed bit rows in frame order, valid after `wait()`.
    """

    def __init__(self, n_local, pitch, device, chunks=8, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_local, self.pitch = int(n_local), int(pitch)
        self.local = torch.empty((self.n_local, self.pitch), dtype=torch.uint8, device=device)
        self.gathered = torch.empty((self.world * self.n_local, self.pitch), dtype=torch.uint8, device=device)
        chunks = max(1, min(int(chunks), self.n_local))
        step = -(-self.n_local // chunks)
        self.bounds = [(c0, min(self.n_local, c0 + step)) for c0 in range(0, self.n_local, step)]
        self.comm = torch.cuda.Stream(device=device)
        self.extracted = [torch.cuda.Event() for _ in self.bounds]
        self.gathered_ev = [None] * len(self.bounds)

    def run(self, frames_local, delta, num_ac, extract_fn=None):
        extract_fn = extract_fn or frame_path.extract_frames
        cur = torch.cuda.current_stream()
        for i, (c0, c1) in enumerate(self.bounds):
            if self.gathered_ev[i] is not None:
                cur.wait_event(self.gathered_ev[i])            # the previous gather of this chunk has read `local`
            extract_fn(frames_local[c0:c1], delta, num_ac, out=self.local[c0:c1])
            self.extracted[i].record(cur)
            self.comm.wait_event(self.extracted[i])
            with torch.cuda.stream(self.comm):
                if len(self.bounds) == 1:
                    dist.all_gather_into_tensor(self.gathered, self.local, group=self.group)
                else:
                    outs = [self.gathered[r * self.n_local + c0: r * self.n_local + c1] for r in range(self.world)]
                    dist.all_gather(outs, self.local[c0:c1], group=self.group)
                ev = torch.cuda.Event()
                ev.record(self.comm)
                self.gathered_ev[i] = ev
        return self.gathered

    def wait(self):
        """Make the current stream wait for every outstanding gather."""
        cur = torch.cuda.current_stream()
        for ev in self.gathered_ev:
            if ev is not None:
                cur.wait_event(ev)
        return self.gathered


class FusedExtractGather:
    """Extract + all-gather as ONE kernel over NVLink peer memory (no NCCL on the data path).

    The gathered buffer (world * F_local rows of `pitch` bytes, frame order) lives in symmetric
    memory: every rank maps every other rank's copy (CUDA IPC / fabric handles through
    torch.distributed._symmetric_memory).  The extract kernel stores each packed word of this
    rank's frames into ITS rows of every rank's copy - through one NVSwitch multicast address
    when the fabric offers it (each word leaves the GPU once and the switch replicates it), else
    with plain stores to each peer mapping.  A symmetric-memory barrier on the stream separates
    producers from consumers: no SMs are set aside for a collective and the transfer overlaps the
    transform tile by tile.  Result: `gathered` on every rank after `run()` (stream-ordered).
    """

    def __init__(self, n_local, pitch, device, group=None, use_multicast=True):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.n_local, self.pitch = int(n_local), int(pitch)
        self.gathered = symm.empty((self.world * self.n_local, self.pitch), dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.gathered, self.group)
        off = self.rank * self.n_local * self.pitch
        self.local = self.gathered[self.rank * self.n_local:(self.rank + 1) * self.n_local]
        ptrs = list(self.hdl.buffer_ptrs)
        self.peer_ptrs = [int(p) + off for r, p in enumerate(ptrs) if r != self.rank]
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0) if use_multicast else 0
        self.multicast_ptr = mc + off if mc else 0
        self.mode = "multicast" if self.multicast_ptr else "peer stores"

    def run(self, frames_local, delta, num_ac, extract_fn=None):
        """Enqueue barrier -> extract (+ remote stores) -> barrier on the current stream."""
        extract_fn = extract_fn or frame_path.extract_frames
        self.hdl.barrier(channel=0)          # every rank is done reading the previous result
        if self.multicast_ptr:
            extract_fn(frames_local, delta, num_ac, out=self.local, multicast_ptr=self.multicast_ptr)
        else:
            extract_fn(frames_local, delta, num_ac, out=self.local, peer_ptrs=self.peer_ptrs)
        self.hdl.barrier(channel=1)          # every rank's rows have landed everywhere
        return self.gathered


class CopyEngineGather:
    """Extract on the SMs, all-gather on the copy engines, overlapped with whatever runs next.

    At 8 ranks every GPU has to RECEIVE 7/8 of the gathered stream (3.2 GB per 1800-frame 1080p
    batch): more NVLink time than the extract kernel needs to produce its share, so stores issued
    from inside the kernel (FusedExtractGather) back-pressure it.  Here the kernel writes its rows
    into this rank's symmetric buffer only, and the rows are then pushed into every peer's buffer
    by DMA (cudaMemcpyAsync into the peer mappings, no SMs, no NCCL) on side streams, so the
    transfer overlaps the NEXT batch's embed kernel.  The symmetric buffer is double-buffered so
    that the next extract does not have to wait for the push that is still reading the previous
    result.  `wait()` makes the current stream wait for the latest push and for a
    symmetric-memory barrier that tells it every rank's rows have landed; `gathered` / `local`
    are the buffers of the latest `run()`.
    """

    def __init__(self, n_local, pitch, device, group=None, n_streams=4, n_buffers=2, use_multicast=False):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.n_local, self.pitch = int(n_local), int(pitch)
        shape = (self.world * self.n_local, self.pitch)
        r0, r1 = self.rank * self.n_local, (self.rank + 1) * self.n_local
        order = [(self.rank + k) % self.world for k in range(1, self.world)]       # spread the inbound load
        self._bufs = []
        self.multicast = False
        for _ in range(max(1, int(n_buffers))):
            g = symm.empty(shape, dtype=torch.uint8, device=device)
            hdl = symm.rendezvous(g, self.group)
            # (use_multicast="force": also with two ranks, where one peer copy is much faster - 0.66 vs
            #  1.43 ms for 460 MB; that is what the 2-GPU test uses)
            want_mc = use_multicast == "force" or (bool(use_multicast) and self.world > 2)
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0) if want_mc else 0
            b = {"gathered": g, "hdl": hdl, "done": None, "mc": 0}
            if mc:
                # NVSwitch multicast (OFF by default): the extract kernel writes into a PRIVATE staging
                # buffer and ONE copy-engine copy per stream sends it to the multicast address of this
                # rank's rows - the switch replicates them into every rank's buffer, this rank's included
                # (so source and destination never overlap).  1/(world-1) of the outbound traffic and HBM
                # reads of the per-peer pushes, but MEASURED SLOWER on 8 x B200: 1,717,456 frames/s against
                # 1,944,604 with one copy per peer (the exchange takes 8.4 instead of 7.4 ms per step).
                b["mc"] = mc + r0 * self.pitch
                b["local"] = torch.empty((self.n_local, self.pitch), dtype=torch.uint8, device=device)
                b["peers"] = []
                self.multicast = True
            else:
                b["local"] = g[r0:r1]
                b["peers"] = [hdl.get_buffer(r, shape, torch.uint8)[r0:r1] for r in order]
            self._bufs.append(b)
        n_side = max(1, min(n_streams, len(order))) if not self.multicast else max(1, min(int(n_streams), 4))
        self.side = [torch.cuda.Stream(device=device) for _ in range(n_side)]
        self._turn = 0
        self._last = self._bufs[0]
        self.mode = ("copy engines -> NVSwitch multicast address (one outbound copy, replicated by the switch), %d streams, %d buffers"
                     if self.multicast else "copy engines (DMA into peer mappings), %d streams, %d buffers") % (len(self.side), len(self._bufs))

    gathered = property(lambda self: self._last["gathered"])
    local = property(lambda self: self._last["local"])
    hdl = property(lambda self: self._last["hdl"])

    def run(self, frames_local, delta, num_ac, extract_fn=None):
        extract_fn = extract_fn or frame_path.extract_frames
        b = self._bufs[self._turn % len(self._bufs)]
        self._turn += 1
        self._last = b
        cur = torch.cuda.current_stream()
        if b["done"] is not None:
            cur.wait_event(b["done"])                  # the previous push out of this buffer has read it
        extract_fn(frames_local, delta, num_ac, out=b["local"])
        ready = torch.cuda.Event()
        ready.record(cur)
        lead = self.side[0]
        lead.wait_event(ready)
        with torch.cuda.stream(lead):
            b["hdl"].barrier(channel=0)                # every rank is done with this buffer's previous result
            go = torch.cuda.Event()
            go.record(lead)
        if b["mc"]:
            from . import _native
            total = self.n_local * self.pitch
            step = -(-total // len(self.side)) // 256 * 256 or total
            off, i = 0, 0
            while off < total:
                n = min(step, total - off)
                st = self.side[i % len(self.side)]
                if st is not lead:
                    st.wait_event(go)
                _native.check(_native.lib().svs_memcpy_d2d_async(b["mc"] + off, b["local"].data_ptr() + off, n, st.cuda_stream),
                              "svs_memcpy_d2d_async")
                off += n
                i += 1
        for i, rows in enumerate(b["peers"]):
            st = self.side[i % len(self.side)]
            if st is not lead:
                st.wait_event(go)
            with torch.cuda.stream(st):
                rows.copy_(b["local"], non_blocking=True)
        joins = []
        for st in self.side[1:]:
            ev = torch.cuda.Event()
            ev.record(st)
            joins.append(ev)
        with torch.cuda.stream(lead):
            for ev in joins:
                lead.wait_event(ev)
            b["hdl"].barrier(channel=1)                # every rank's rows have landed everywhere
            b["done"] = torch.cuda.Event()
            b["done"].record(lead)
        return b["gathered"]

    def wait(self):
        if self._last["done"] is not None:
            torch.cuda.current_stream().wait_event(self._last["done"])
        return self._last["gathered"]
