#!/usr/bin/env python
"""bench.py - 1080p frames/s of the DCT-QIM embed+extract round trip on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2]): 1800 synthetic 1920x1080 BGR frames PER GPU, 63 AC
coefficients per block (maximum capacity), delta 20, payload = random bits filling every frame.
One step = one pass of the hot path over that batch: embed every frame (BGR in, gray stego
out), then extract every stego frame (packed bits out); with N > 1 each rank owns a contiguous
frame range + its payload slice and the extracted bitstreams are all-gathered over NCCL.
Inputs are resident in HBM for `value`; `e2e` runs the same round trip through the host-buffer
C ABI (pinned host memory, H2D + D2H inside the timed region).  See DESIGN.md section 6.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, CH = 1080, 1920, 3
FRAMES_PER_GPU = int(os.environ.get("SVS_BENCH_FRAMES", "1800"))
NUM_AC, DELTA = 63, 20
LO, HI = 64, 192            # mid-range noise: the reference's own round trip is error-free here
METRIC = "1080p frames/s embed+extract (device-timed)"
UNIT = "frames/s"


def workload_name():
    return "1080p x %d frames per GPU, BGR u8, 63 AC, delta 20, uniform[64,192) noise" % FRAMES_PER_GPU


def algorithmic_bytes(frames, cap_bits):
    """embed = 3HW + HW + cap/8, extract(gray) = HW + cap/8 per frame."""
    px = H * W
    nb = (cap_bits + 7) // 8
    return frames * (3 * px + px + nb), frames * (px + nb)


def fp32_pipe_roofline(frames, embed_ms, clocks):
    """FP32-pipe lane operations the op-exact embed kernel must issue vs the pipe's peak."""
    blocks = (H // 8) * (W // 8)
    lane_ops = frames * blocks * (32 * 54 + 2 * min(NUM_AC, 63) + 64)
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    peak = 148 * 128 * sm_mhz * 1e6
    ach = lane_ops / (embed_ms / 1000.0)
    return {"bound": "fp32 issue (secondary; not an HBM or tensor bound)", "kernel": "fast::embed_fast_kernel<3,1,true>",
            "achieved": ach / 1e12, "peak": peak / 1e12, "unit": "T lane-op/s", "frac": ach / peak,
            "lane_ops_per_block": 32 * 54 + 2 * min(NUM_AC, 63) + 64}


def bind_to_gpu_numa_node(device_index):
    """Multi-rank runs: keep the rank (and with it the first-touch placement of its pinned host
    buffers) on the CPUs NVML reports as local to its GPU, so that the e2e copies do not cross the
    socket interconnect.  Best effort; returns the number of CPUs bound to (0 = left alone)."""
    if os.environ.get("SVS_BENCH_NUMA_BIND", "1") == "0":
        return 0
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            tok = vis.split(",")[device_index].strip()
            h = pynvml.nvmlDeviceGetHandleByUUID(tok) if tok.startswith("GPU-") else pynvml.nvmlDeviceGetHandleByIndex(int(tok))
        else:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (int(words[i // 64]) >> (i % 64)) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes per embed launch from the last committed `ncu --set full` capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        if t.get("frames") and t.get("embed_dram_bytes"):
            return t["embed_dram_bytes"] / t["frames"] * FRAMES_PER_GPU
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.2] or [r for (_, r) in self.rows]
        sm, smax, reasons, power = [], 0.0, set(), 0.0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                smax = max(smax, float(r[2]))
                power = max(power, float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax or None,
                "power_w_max": power or None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes oracle/)
# ----------------------------------------------------------------------------------------------
def _port_round_trip(job):
    """One 1080p/63 round trip (or the top `rows` pixel rows of one) through the loop-structured
    port of the reference, in one process."""
    seed, rows = job
    import numpy as np
    from oracle import ref_port
    from tests.synth import synth_frames, synth_bits, bits_to_str
    frame = synth_frames("cpu%d" % seed, (H, W, CH), LO, HI)[:rows]
    cap = (rows // 8) * (W // 8) * NUM_AC
    seg = bits_to_str(synth_bits("cpu%d" % seed, cap))
    t0 = time.perf_counter()
    _, stego, k = ref_port.proses_frame_qim_dct(frame, 'embed', DELTA, seg, num_ac_coeffs_to_use=NUM_AC)
    out = ref_port.proses_frame_qim_dct(stego, 'extract', DELTA, num_ac_coeffs_to_use=NUM_AC)
    dt = time.perf_counter() - t0
    assert k == cap and out == seg
    return dt


def cpu_port_baseline(rows=H, pool=None, step=0):
    """Reference-structured CPU path on all host cores: one process per core, disjoint frames
    (`rows` < 1080 times a strip of each frame; the per-block cost does not depend on the strip)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(cores)
    try:
        t0 = time.perf_counter()
        per = pool.map(_port_round_trip, [(step * cores + i, rows) for i in range(cores)])
        wall = time.perf_counter() - t0
    finally:
        if own:
            pool.close()
            pool.join()
    frac = rows / float(H)
    return {"value": cores * frac / max(per), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d processes x %d of 1080 pixel rows of one 1080p/63-AC/delta-20 frame each, embed+extract "
                      "through oracle/ref_port.py (per-block scipy.fftpack + per-coefficient round(), the "
                      "reference's own loop structure); rate = frames / slowest process; wall %.1fs" % (cores, rows, wall),
            "per_core_frames_per_s": frac / statistics.mean(per)}


def cpu_c_oracle_rate(frames=32):
    """The plain-C op-exact oracle, all host threads (a far stronger CPU figure than the reference)."""
    import numpy as np
    from oracle import c_oracle
    from tests.synth import synth_frames, synth_bits
    threads = c_oracle.max_threads()
    f = synth_frames("cpuc", (frames, H, W, CH), LO, HI)
    cap = (H // 8) * (W // 8) * NUM_AC
    packed = np.packbits(synth_bits("cpuc", frames * cap))
    t0 = time.perf_counter()
    stego, _, _ = c_oracle.embed_frames(f, packed, frames * cap, DELTA, NUM_AC, threads=threads, want_gray=False)
    c_oracle.extract_frames(stego, DELTA, NUM_AC, threads=threads)
    dt = time.perf_counter() - t0
    return {"value": frames / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d frames through oracle/dctqim_oracle.c (op-exact plain C, pthreads)" % frames}


def run_reference_arm(args, rank):
    """--impl reference: the reference's CPU implementation of the path on the host cores.

    A step is a bounded sample of the workload: every host core runs the top `rows` pixel rows of
    one 1080p frame; `rows` is chosen so that warmup + steps fit SVS_REF_BUDGET_S (default 150 s)."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    budget = float(os.environ.get("SVS_REF_BUDGET_S", "150"))
    per_frame_s = 6.0                                   # one 1080p/63 round trip per core, measured ~4.7 s
    n = max(1, args.warmup + args.steps)
    rows = int(H * min(1.0, budget / n / per_frame_s)) // 8 * 8
    rows = max(8, min(H, rows))
    vals, last = [], None
    pool = mp.get_context("spawn").Pool(cores)
    try:
        for i in range(args.warmup + args.steps):
            last = cpu_port_baseline(rows=rows, pool=pool, step=i)
            if i >= args.warmup:
                vals.append(last["value"])
    finally:
        pool.close()
        pool.join()
    v = statistics.mean(vals)
    last["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * cores * rows / H / v,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(),
                       "step": "bounded sample: %d of 1080 pixel rows of one frame per host core" % rows},
            "cpu_baseline": last,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import svs_b200
    from svs_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # multi-GPU exchange of the extracted bits: "fused" = the extract kernel stores its rows straight
    # into every rank's gathered buffer over NVLink (symmetric memory; multicast when available),
    # "nccl" = chunked extract + ncclAllGather overlapped on a side stream, "nccl-seq" = plain all-gather
    gather_mode = os.environ.get("SVS_GATHER", "push") if world > 1 else "none"
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else 0     # (N = 1 keeps all cores for cpu_baseline)
    if world > 1:
        reserved = int(os.environ.get("SVS_RESERVED_SMS", "8")) if gather_mode == "nccl" else 0
        if gather_mode == "nccl":
            # the kernels are persistent and fill every SM: leave a few SMs to the all-gather so that
            # it really overlaps, and tell NCCL not to ask for more CTAs than that
            os.environ.setdefault("NCCL_MAX_NCHANNELS", str(max(1, reserved)))
        dist.init_process_group("nccl", device_id=dev)
        svs_b200.lib().svs_set_reserved_sms(reserved)
    args.warmup = max(args.warmup, 3)

    F = FRAMES_PER_GPU
    cap = svs_b200.capacity_bits(H, W, NUM_AC)
    nbytes = (cap + 7) // 8
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    frames = torch.randint(LO, HI, (F, H, W, CH), dtype=torch.uint8, device=dev, generator=gen)
    payload = torch.randint(0, 256, (F * nbytes,), dtype=torch.uint8, device=dev, generator=gen)
    total_bits = F * cap
    stego = torch.empty((F, H, W), dtype=torch.uint8, device=dev)
    pitch = svs_b200.bits_row_bytes(H, W, NUM_AC)
    bits = torch.empty((F, pitch), dtype=torch.uint8, device=dev)
    overlap = fused = None
    gathered = None
    if gather_mode == "fused":
        try:
            fused = sharding.FusedExtractGather(F, pitch, dev, use_multicast=os.environ.get("SVS_MULTICAST", "1") != "0")
            bits, gathered = fused.local, fused.gathered
        except Exception as exc:       # no peer mapping on this box: say so and use NCCL (every rank fails alike)
            sys.stderr.write("rank %d: symmetric memory unavailable (%r); falling back to NCCL all-gather\n" % (rank, exc))
            gather_mode = "nccl"
    if gather_mode == "push":
        try:
            overlap = sharding.CopyEngineGather(F, pitch, dev, n_streams=int(os.environ.get("SVS_PUSH_STREAMS", "4")),
                                                n_buffers=int(os.environ.get("SVS_PUSH_BUFFERS", "2")))
            bits, gathered = overlap.local, overlap.gathered
        except Exception as exc:
            sys.stderr.write("rank %d: symmetric memory unavailable (%r); falling back to NCCL all-gather\n" % (rank, exc))
            gather_mode = "nccl"
    if gather_mode == "nccl":
        overlap = sharding.OverlappedExtractGather(F, pitch, dev, chunks=int(os.environ.get("SVS_GATHER_CHUNKS", "2")))
        bits, gathered = overlap.local, overlap.gathered
    elif world > 1 and fused is None and overlap is None:
        gathered = torch.empty((world * F, pitch), dtype=torch.uint8, device=dev)
    L = svs_b200.lib()
    stream = torch.cuda.current_stream()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    marks = []

    def step(timed):
        e0, e1, e2, e3 = (ev(), ev(), ev(), ev()) if timed else (None,) * 4
        if timed:
            e0.record(stream)
        svs_b200.embed_frames(frames, payload, total_bits, DELTA, NUM_AC, out=stego)
        if timed:
            e1.record(stream)
        if fused is not None:     # one kernel: extract + stores into every rank's gathered buffer, then a barrier
            fused.run(stego, DELTA, NUM_AC)
        elif overlap is not None:  # chunked extract, each chunk all-gathered on a side stream (overlaps what follows)
            overlap.run(stego, DELTA, NUM_AC)
        else:
            svs_b200.extract_frames(stego, DELTA, NUM_AC, out=bits)
            if world > 1:
                sharding.all_gather_bits(bits, out=gathered)
        if timed:
            e2.record(stream)
        if timed:
            e3.record(stream)
            marks.append((e0, e1, e2, e3))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = L.svs_kernel_launch_count()
    barrier()
    t_wall0 = time.time()
    start, stop = ev(), ev()
    start.record(stream)
    for _ in range(args.steps):
        step(True)
    if overlap is not None:
        overlap.wait()              # every gather has landed before the clock stops
    stop.record(stream)
    barrier()
    t_wall1 = time.time()
    launches = L.svs_kernel_launch_count() - launches0
    ms_total = start.elapsed_time(stop)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    embed_ms = statistics.mean(a.elapsed_time(b) for a, b, _, _ in marks)
    extract_ms = statistics.mean(b.elapsed_time(c) for _, b, c, _ in marks)
    gather_ms = statistics.mean(c.elapsed_time(d) for _, _, c, d in marks)
    t = torch.tensor([ms_total, embed_ms, extract_ms, gather_ms, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, embed_ms, extract_ms, gather_ms = tmax[:4].tolist()
        launches = int(tsum[4].item())

    if gather_mode == "push":
        bits, gathered = overlap.local, overlap.gathered       # the buffers of the last step
    # correctness of what was timed: mid-range frames -> the round trip returns the payload
    ok = bool(torch.equal(bits[:, :nbytes].reshape(-1), payload))
    parity_detail = {"own_rows": ok}
    if world > 1:
        mine = gathered[rank * F:(rank + 1) * F, :nbytes].reshape(-1)
        parity_detail["own_rows_in_gathered"] = bool(torch.equal(mine, payload))
        # ... and every OTHER rank's rows arrived intact: compare per-rank checksums of the gathered
        # stream with the checksums the owners computed from their payloads
        w8 = torch.arange(1, 8192 + 1, device=dev, dtype=torch.int64)

        def checksum(rows):
            v = rows.reshape(-1).to(torch.int64)
            pad = (-v.numel()) % w8.numel()
            v = torch.nn.functional.pad(v, (0, pad)).reshape(-1, w8.numel())
            return (v * w8).sum()

        own = checksum(payload).reshape(1)
        sums = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sums, own)
        torch.cuda.synchronize()
        per_rank = [int(checksum(gathered[r * F:(r + 1) * F, :nbytes])) == int(sums[r]) for r in range(world)]
        parity_detail["rows_of_rank_ok"] = per_rank
        ok = ok and parity_detail["own_rows_in_gathered"] and all(per_rank)
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())

    # ---------------- e2e through the host-buffer C ABI (pinned host memory) ----------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, svs_b200, torch, dist, frames, payload, dev, local_rank, world, cap, nbytes)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = ms_total / args.steps
    value = world * F / (ms_per_step / 1000.0)
    eb, xb = algorithmic_bytes(F, cap)
    peak, peak_src = measured_peak()
    ach = eb / (embed_ms / 1000.0) / 1e9
    ach_x = xb / (extract_ms / 1000.0) / 1e9
    ach_rt = (eb + xb) / ((embed_ms + extract_ms) / 1000.0) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "frames_per_gpu": F, "height": H, "width": W, "num_ac": NUM_AC,
                   "delta": DELTA, "step": "embed (BGR->gray stego) + extract (gray stego->packed bits)"
                                           + {"none": "", "fused": " + all-gather of the bits fused into the extract kernel (stores to every rank's buffer over NVLink) + symmetric-memory barrier",
                                              "push": " + all-gather of the bits by DMA into every rank's symmetric buffer (copy engines, overlapped with the next batch) + symmetric-memory barrier",
                                              "nccl": " + NCCL all-gather of bits (chunked, overlapped on a side stream)",
                                              "nccl-seq": " + NCCL all-gather of bits"}[gather_mode],
                   "l2": "inputs (%.1f GB per step) far exceed the 126 MB L2; no flush needed" % ((eb + xb) / 1e9),
                   "parallelism": "frame-sharded x%d" % world,
                   "host": "rank bound to the %d CPUs local to its GPU" % numa_cpus if numa_cpus else "no CPU binding"},
        "mpixel_per_s": value * H * W / 1e6,
        "roofline": {"bound": "hbm", "kernel": "fast::embed_fast_kernel<3,1,true>", "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "traffic": recorded_traffic(), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": eb, "launch_ms": embed_ms},
        "roofline_extract": {"bound": "hbm", "kernel": "fast::extract_fast_kernel<1,true>", "achieved": ach_x, "peak": peak,
                             "unit": "GB/s", "frac": ach_x / peak, "algorithmic_bytes_per_launch": xb,
                             "launch_ms": extract_ms},
        "roofline_round_trip": {"achieved": ach_rt, "peak": peak, "unit": "GB/s", "frac": ach_rt / peak},
        # the co-limiter the HBM fraction has to be read with: reproducing scipy's float32 DCT bit for bit
        # takes 54 un-fused FP32 operations per 8-point transform (no FMA contraction), i.e.
        # 32 transforms x 54 + 63 x 2 + 64 per block for embed; peak = SMs x 128 lanes x SM clock
        "roofline_fp32_pipe": fp32_pipe_roofline(F, embed_ms, clocks),
        "allgather": None if world == 1 else (
            "fused into extract_kernel: %s over NVLink into symmetric memory, barrier after each launch" % fused.mode
            if fused is not None else
            "%s; overlaps the next batch, all complete inside the timed region" % overlap.mode if gather_mode == "push" else
            "NCCL, overlapped: %d chunks on a side stream, %s SMs left to NCCL, all complete inside the timed region"
            % (len(overlap.bounds), os.environ.get("SVS_RESERVED_SMS", "8")) if overlap is not None
            else "NCCL, sequential on the compute stream"),
        "gpu_launches": launches, "clocks": clocks, "parity_check": ok, "parity_detail": parity_detail,
        "e2e": e2e,
    }
    if world == 1 and not args.no_cpu:
        try:
            line["cpu_baseline"] = cpu_port_baseline()
            line["cpu_baseline_c_oracle"] = cpu_c_oracle_rate()
        except Exception as exc:                       # the GPU numbers stay valid without it
            line["cpu_baseline"] = {"error": repr(exc)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        raise SystemExit("round trip did not return the payload")


def run_e2e(args, svs_b200, torch, dist, frames, payload, dev, local_rank, world, cap, nbytes):
    """Same round trip through svs_*_frames_host: pinned host buffers, copies inside the timing."""
    import ctypes
    L = svs_b200.lib()
    # a third of the device batch per step (600 frames = 5.1 GB of pinned host memory per rank): the
    # link-bound rate does not depend on the batch size, the host memory of an 8-rank box does
    F = min(frames.shape[0], int(os.environ.get("SVS_BENCH_E2E_FRAMES", "600")))
    frames, payload = frames[:F], payload[:F * nbytes]
    steps = max(1, min(args.steps, int(os.environ.get("SVS_BENCH_E2E_STEPS", "3"))))
    try:
        h_frames = torch.empty(frames.shape, dtype=torch.uint8).pin_memory()
        h_frames.copy_(frames)
        h_payload = payload.cpu().pin_memory()
        h_stego = torch.empty((F, H, W), dtype=torch.uint8).pin_memory()
        h_bits = torch.empty((F, nbytes), dtype=torch.uint8).pin_memory()
    except Exception as exc:
        return {"error": "pinned allocation failed: %r" % (exc,)}
    ctx = ctypes.c_void_p()
    staging = int(os.environ.get("SVS_BENCH_E2E_STAGING_MB", "192")) << 20
    svs_b200._native.check(L.svs_ctx_create(local_rank, staging, ctypes.byref(ctx)), "svs_ctx_create")

    def one():
        rc = L.svs_embed_frames_host(ctx, h_frames.data_ptr(), CH, F, H, W, H * W * CH, W * CH,
                                     h_payload.data_ptr(), 0, F * cap, float(DELTA), NUM_AC,
                                     h_stego.data_ptr(), 1, None, None, None)
        svs_b200._native.check(rc, "svs_embed_frames_host")
        rc = L.svs_extract_frames_host(ctx, h_stego.data_ptr(), 1, F, H, W, H * W, W, float(DELTA), NUM_AC,
                                       h_bits.data_ptr(), nbytes)
        svs_b200._native.check(rc, "svs_extract_frames_host")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    one()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    barrier()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    ok = bool(torch.equal(h_bits.reshape(-1), h_payload))
    L.svs_ctx_destroy(ctx)
    px = H * W
    return {"value": world * F * steps / dt, "unit": UNIT,
            "h2d_bytes_per_step": F * (3 * px + nbytes + px), "d2h_bytes_per_step": F * (px + nbytes),
            "steps": steps, "frames_per_step_per_gpu": F, "ms_per_step": 1000.0 * dt / steps, "parity_check": ok,
            # what bounds it: the host<->device link, not the kernels (both directions run concurrently)
            "pcie_h2d_gbs": F * (3 * px + nbytes + px) * steps / dt / 1e9,
            "pcie_d2h_gbs": F * (px + nbytes) * steps / dt / 1e9,
            "api": "svs_embed_frames_host + svs_extract_frames_host (C ABI, pinned host buffers, 3-slot chunk pipeline, %d MB of device staging)" % (staging >> 20)}


if __name__ == "__main__":
    main()
