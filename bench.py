#!/usr/bin/env python
"""bench.py - frames/s of the DCT-QIM embed+extract round trip on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 1080p|4k] [--num-ac 63] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json):
  1080p (default, configs[2], the headline): 1800 synthetic 1920x1080 BGR frames PER GPU (weak
        scaling), 63 AC coefficients per block (maximum capacity), delta 20.
  4k    (configs[3]): 600 synthetic 3840x2160 BGR frames IN TOTAL, frame-sharded over the N ranks
        (strong scaling; 75 frames per rank at N = 8), --num-ac 63 or 10, delta 20.
Payload = random bits filling every frame.  One step = one pass of the hot path over the batch:
embed every frame (BGR in, gray stego out), then extract every stego frame (packed bits out);
with N > 1 each rank owns a contiguous frame range + its payload slice and the extracted
bitstreams are all-gathered.  Inputs are resident in HBM for `value`; `e2e` runs the same round
trip through the host-buffer C ABI (pinned host memory, H2D + D2H inside the timed region).
`parity` compares frames of the TIMED tensors with the CPU oracle (outside the timed region).
See DESIGN.md section 6.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CH = 3
DELTA = 20
LO, HI = 64, 192            # mid-range noise: the reference's own round trip is error-free here
UNIT = "frames/s"
WORKLOADS = {
    #          H     W     frames               scaling   metric
    "1080p": (1080, 1920, ("per_gpu", 1800), "weak", "1080p frames/s embed+extract (device-timed)"),
    "4k": (2160, 3840, ("total", 600), "strong", "4K frames/s embed+extract (device-timed)"),
}


class Workload:
    def __init__(self, name, num_ac, world):
        self.name = name
        self.H, self.W, (mode, frames), self.scaling, self.metric = WORKLOADS[name]
        env = os.environ.get("SVS_BENCH_FRAMES")
        if env:
            frames = int(env)
        self.frames_mode = mode
        self.frames_per_gpu = frames if mode == "per_gpu" else max(1, frames // world)
        self.num_ac = num_ac
        self.blocks = (self.H // 8) * (self.W // 8)
        self.cap = self.blocks * min(max(num_ac, 0), 63)
        self.nbytes = (self.cap + 7) // 8

    def describe(self):
        per = "%d frames per GPU" % self.frames_per_gpu if self.frames_mode == "per_gpu" else \
            "%d frames per GPU (a fixed total split over the ranks)" % self.frames_per_gpu
        return "%s (%dx%d) x %s, BGR u8, %d AC, delta %d, uniform[%d,%d) noise" % (
            self.name, self.W, self.H, per, self.num_ac, DELTA, LO, HI)

    def algorithmic_bytes(self, frames):
        """embed = 3HW + HW + cap/8, extract(gray) = HW + cap/8 per frame."""
        px = self.H * self.W
        return frames * (3 * px + px + self.nbytes), frames * (px + self.nbytes)


def fp32_lane_ops_per_block(num_ac):
    """FP32 lane operations the op-exact embed must issue per 8x8 block: 32 eight-point transforms
    of 54 un-fused operations, the bias removal of the axis-0 pass (8), two FMAs per quantised
    coefficient and the always-exact quantiser of the three tie-prone coefficients (3 x 9)."""
    n = min(max(num_ac, 0), 63)
    return 32 * 54 + 8 + 2 * n + 27


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


def recorded_profile():
    """Figures of the last committed `ncu --set full` capture (profiles/traffic.json): DRAM bytes
    per frame of the embed launch and the operand-delivery model of profiles/rf_model.py."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


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
def reference_function():
    """The reference's own proses_frame_qim_dct from the staged, SHA-256-verified copy under
    oracle/_ref (kind "reference"); the loop-structured port (kind "port") only when nothing is staged."""
    from oracle import stage_ref
    if stage_ref.available():
        return stage_ref.import_reference()["config_and_setup"].proses_frame_qim_dct, "reference"
    from oracle import ref_port
    return ref_port.proses_frame_qim_dct, "port"


def _cpu_round_trip(job):
    """One round trip over the top `rows` pixel rows of one synthetic frame, in one process."""
    seed, rows, H, W, num_ac = job
    import contextlib
    import io
    from tests.synth import synth_frames, synth_bits, bits_to_str
    fn, kind = reference_function()
    frame = synth_frames("cpu%d" % seed, (H, W, CH), LO, HI)[:rows]
    cap = (rows // 8) * (W // 8) * min(num_ac, 63)
    seg = bits_to_str(synth_bits("cpu%d" % seed, cap))
    with contextlib.redirect_stdout(io.StringIO()):
        t0 = time.perf_counter()
        _, stego, k = fn(frame, 'embed', DELTA, seg, num_ac_coeffs_to_use=num_ac)
        out = fn(stego, 'extract', DELTA, num_ac_coeffs_to_use=num_ac)
        dt = time.perf_counter() - t0
    assert k == cap and out == seg
    return dt, kind


def cpu_reference_baseline(wl, rows=None, pool=None, step=0):
    """The reference's CPU path on all host cores: one process per core, disjoint frames (`rows`
    < H times a strip of each frame; the per-block cost does not depend on the strip)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    rows = wl.H if rows is None else rows
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(cores)
    try:
        t0 = time.perf_counter()
        res = pool.map(_cpu_round_trip, [(step * cores + i, rows, wl.H, wl.W, wl.num_ac) for i in range(cores)])
        wall = time.perf_counter() - t0
    finally:
        if own:
            pool.close()
            pool.join()
    per, kind = [r[0] for r in res], res[0][1]
    frac = rows / float(wl.H)
    what = ("the unmodified reference function (oracle/_ref/config_and_setup.py, staged from /root/reference, SHA-256 verified)"
            if kind == "reference" else "oracle/ref_port.py (the reference's loop structure; nothing staged under oracle/_ref)")
    return {"value": cores * frac / max(per), "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d processes x %d of %d pixel rows of one %s/%d-AC/delta-%d frame each, embed+extract through %s; "
                      "rate = frames / slowest process; wall %.1fs" % (cores, rows, wl.H, wl.name, wl.num_ac, DELTA, what, wall),
            "per_core_frames_per_s": frac / statistics.mean(per)}


def cpu_c_oracle_rate(wl, frames=None):
    """The plain-C op-exact oracle, all host threads (a far stronger CPU figure than the reference)."""
    import numpy as np
    from oracle import c_oracle
    from tests.synth import synth_frames, synth_bits
    threads = c_oracle.max_threads()
    frames = frames or max(2, int(32 * (1080 * 1920) / (wl.H * wl.W)))
    f = synth_frames("cpuc", (frames, wl.H, wl.W, CH), LO, HI)
    packed = np.packbits(synth_bits("cpuc", frames * wl.cap))
    t0 = time.perf_counter()
    stego, _, _ = c_oracle.embed_frames(f, packed, frames * wl.cap, DELTA, wl.num_ac, threads=threads, want_gray=False)
    c_oracle.extract_frames(stego, DELTA, wl.num_ac, threads=threads)
    dt = time.perf_counter() - t0
    return {"value": frames / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d frames through oracle/dctqim_oracle.c (op-exact plain C, pthreads)" % frames}


def run_reference_arm(args, rank, wl):
    """--impl reference: the reference's CPU implementation of the path on the host cores.

    A step is a bounded sample of the workload: every host core runs the top `rows` pixel rows of
    one frame; `rows` is chosen so that warmup + steps fit SVS_REF_BUDGET_S (default 150 s)."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    budget = float(os.environ.get("SVS_REF_BUDGET_S", "150"))
    per_frame_s = 13.0 * (wl.H * wl.W) / (1080.0 * 1920.0) * (0.35 + 0.65 * min(wl.num_ac, 63) / 63.0)
    n = max(1, args.warmup + args.steps)
    rows = int(wl.H * min(1.0, budget / n / per_frame_s)) // 8 * 8
    rows = max(8, min(wl.H, rows))
    vals, last = [], None
    pool = mp.get_context("spawn").Pool(cores)
    try:
        for i in range(args.warmup + args.steps):
            last = cpu_reference_baseline(wl, rows=rows, pool=pool, step=i)
            if i >= args.warmup:
                vals.append(last["value"])
    finally:
        pool.close()
        pool.join()
    v = statistics.mean(vals)
    last["value"] = v
    line = {"impl": "reference", "metric": wl.metric, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * cores * rows / wl.H / v,
            "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.describe(),
                       "step": "bounded sample: %d of %d pixel rows of one frame per host core" % (rows, wl.H)},
            "cpu_baseline": last,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# oracle spot check of the timed tensors (outside the timed region)
# ----------------------------------------------------------------------------------------------
def psnr(a, b):
    import numpy as np
    mse = float(((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * math.log10(255.0 * 255.0 / mse)


def oracle_spot_check(wl, frames, payload, stego, bits, which):
    """Frames `which` of this rank's TIMED tensors through the C oracle: the fields the
    north_star asks for.  Bit-exactness is the bar; the LSB figures are reported for the record."""
    import numpy as np
    from oracle import c_oracle
    threads = min(len(which), c_oracle.max_threads())
    idx = sorted(set(int(i) for i in which))
    f_host = np.stack([frames[i].cpu().numpy() for i in idx])
    pay = payload.cpu().numpy()
    px_diff = max_diff = bits_diff = 0
    worst_psnr_delta = 0.0
    for k, i in enumerate(idx):
        want_s, gray, nb = c_oracle.embed_frames(f_host[k:k + 1], pay, wl.cap, DELTA, wl.num_ac, bit_offset=i * wl.cap,
                                                 threads=threads)
        got_s = stego[i].cpu().numpy()
        d = np.abs(want_s[0].astype(np.int16) - got_s.astype(np.int16))
        px_diff += int((d > 0).sum())
        max_diff = max(max_diff, int(d.max()))
        want_b = c_oracle.extract_frames(want_s, DELTA, wl.num_ac, threads=threads)[0]
        got_b = bits[i, :wl.nbytes].cpu().numpy()
        bits_diff += int(np.unpackbits(want_b ^ got_b).sum())
        pr, pg = psnr(gray[0], want_s[0]), psnr(gray[0], got_s)
        worst_psnr_delta = max(worst_psnr_delta, abs(pr - pg) if math.isfinite(pr) and math.isfinite(pg) else 0.0)
    n_px = len(idx) * wl.H * wl.W
    return {"checked_frames": idx, "against": "oracle/dctqim_oracle.c (op-exact CPU restatement, pinned to the reference's golden vectors)",
            "stego_px_diff": px_diff, "stego_max_abs_diff": max_diff, "frac_px_diff": px_diff / float(n_px),
            "psnr_delta_db": worst_psnr_delta, "bits_diff": bits_diff}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--num-ac", type=int, default=63)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = Workload(args.workload, args.num_ac, world)

    if args.impl == "reference":
        run_reference_arm(args, rank, wl)
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
    H, W, NUM_AC = wl.H, wl.W, wl.num_ac
    # multi-GPU exchange of the extracted bits: "push" = copy engines into every rank's symmetric buffer
    # (default), "fused" = the extract kernel stores its rows straight into every rank's gathered buffer
    # over NVLink (multicast when available), "nccl" = chunked extract + ncclAllGather overlapped on a
    # side stream, "nccl-seq" = plain all-gather
    gather_mode = os.environ.get("SVS_GATHER", "push") if world > 1 else "none"
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else 0     # (N = 1 keeps all cores for cpu_baseline)
    if world > 1:
        reserved = int(os.environ.get("SVS_RESERVED_SMS", "8")) if gather_mode == "nccl" else 0
        if gather_mode == "nccl":
            os.environ.setdefault("NCCL_MAX_NCHANNELS", str(max(1, reserved)))
        dist.init_process_group("nccl", device_id=dev)
        svs_b200.lib().svs_set_reserved_sms(reserved)
    args.warmup = max(args.warmup, 3)

    F = wl.frames_per_gpu
    cap, nbytes = wl.cap, wl.nbytes
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    frames = torch.randint(LO, HI, (F, H, W, CH), dtype=torch.uint8, device=dev, generator=gen)
    pay_len = (F * cap + 7) // 8
    pay_alloc = (pay_len + 3) // 4 * 4                    # the kernels read whole 32-bit words
    # two payloads: the LAST timed step embeds the second one, so that a stale double-buffered
    # result of an earlier step cannot pass the final comparison
    payloads = [torch.zeros(pay_alloc, dtype=torch.uint8, device=dev) for _ in range(2)]
    for p in payloads:
        p[:pay_len] = torch.randint(0, 256, (pay_len,), dtype=torch.uint8, device=dev, generator=gen)
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
                                                n_buffers=int(os.environ.get("SVS_PUSH_BUFFERS", "2")),
                                                use_multicast=os.environ.get("SVS_PUSH_MULTICAST", "0") == "1")
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

    def step(timed, payload):
        e0, e1, e2 = (ev(), ev(), ev()) if timed else (None,) * 3
        if timed:
            e0.record(stream)
        svs_b200.embed_frames(frames, payload, total_bits, DELTA, NUM_AC, out=stego)
        if timed:
            e1.record(stream)
        if fused is not None:     # one kernel: extract + stores into every rank's gathered buffer, then a barrier
            fused.run(stego, DELTA, NUM_AC)
        elif overlap is not None:  # extract, then the exchange on side streams (overlaps what follows)
            overlap.run(stego, DELTA, NUM_AC)
        else:
            svs_b200.extract_frames(stego, DELTA, NUM_AC, out=bits)
            if world > 1:
                sharding.all_gather_bits(bits, out=gathered)
        if timed:
            e2.record(stream)
            marks.append((e0, e1, e2))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False, payloads[0])
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
    for i in range(args.steps):
        step(True, payloads[1] if i == args.steps - 1 else payloads[0])
    if overlap is not None:
        overlap.wait()              # every gather has landed before the clock stops
    stop.record(stream)
    barrier()
    t_wall1 = time.time()
    launches = L.svs_kernel_launch_count() - launches0
    ms_total = start.elapsed_time(stop)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    payload = payloads[1]           # what the last timed step embedded

    embed_ms = statistics.mean(a.elapsed_time(b) for a, b, _ in marks)
    extract_ms = statistics.mean(b.elapsed_time(c) for _, b, c in marks)
    t = torch.tensor([ms_total, embed_ms, extract_ms, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, embed_ms, extract_ms = tmax[:3].tolist()
        launches = int(tsum[3].item())

    if gather_mode == "push":
        bits, gathered = overlap.local, overlap.gathered       # the buffers of the last step
    # correctness of what was timed, part 1: mid-range frames -> the round trip returns the payload
    ok = bool(torch.equal(bits[:, :nbytes].reshape(-1)[:pay_len], payload[:pay_len]))
    parity_detail = {"own_rows": ok}
    # part 2: frames 0 / F/2 / F-1 of the timed tensors against the CPU oracle
    try:
        parity = oracle_spot_check(wl, frames, payload, stego, bits, [0, F // 2, F - 1])
    except Exception as exc:
        parity = {"error": repr(exc)}
    if world > 1:
        mine = gathered[rank * F:(rank + 1) * F, :nbytes].reshape(-1)[:pay_len]
        parity_detail["own_rows_in_gathered"] = bool(torch.equal(mine, payload[:pay_len]))
        # ... every OTHER rank's rows arrived intact: per-rank checksums of the gathered stream
        # against the checksums the owners computed from their payloads
        w8 = torch.arange(1, 8192 + 1, device=dev, dtype=torch.int64)

        def checksum(rows):
            v = rows.reshape(-1).to(torch.int64)
            pad = (-v.numel()) % w8.numel()
            v = torch.nn.functional.pad(v, (0, pad)).reshape(-1, w8.numel())
            return (v * w8).sum()

        own = checksum(bits[:, :nbytes]).reshape(1)
        sums = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sums, own)
        torch.cuda.synchronize()
        per_rank = [int(checksum(gathered[r * F:(r + 1) * F, :nbytes])) == int(sums[r]) for r in range(world)]
        parity_detail["rows_of_rank_ok"] = per_rank
        # ... and one frame of ANOTHER rank through the oracle: every rank publishes its middle stego
        # frame; rank r extracts the one of rank r+1 on the CPU and compares it with the gathered row
        try:
            from oracle import c_oracle
            mid = F // 2
            shared = torch.empty((world, H, W), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(shared, stego[mid].contiguous())
            other = (rank + 1) % world
            want = c_oracle.extract_frames(shared[other].cpu().numpy()[None], DELTA, NUM_AC)[0]
            got = gathered[other * F + mid, :nbytes].cpu().numpy()
            diff = int(np.unpackbits(want ^ got).sum())
            t_diff = torch.tensor([diff], dtype=torch.int64, device=dev)
            dist.all_reduce(t_diff, op=dist.ReduceOp.SUM)
            parity["other_rank_frame"] = {"what": "rank r checks frame F/2 of rank r+1 in its gathered stream against the oracle",
                                          "bits_diff_all_ranks": int(t_diff.item())}
            ok = ok and int(t_diff.item()) == 0
        except Exception as exc:
            parity["other_rank_frame"] = {"error": repr(exc)}
        ok = ok and parity_detail["own_rows_in_gathered"] and all(per_rank)
    ok = ok and parity.get("stego_px_diff", 1) == 0 and parity.get("bits_diff", 1) == 0
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())

    # ---------------- e2e through the host-buffer C ABI (pinned host memory) ----------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, wl, svs_b200, torch, dist, frames, payload, dev, local_rank, world)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = ms_total / args.steps
    value = world * F / (ms_per_step / 1000.0)
    eb, xb = wl.algorithmic_bytes(F)
    peak, peak_src = measured_peak()
    ach = eb / (embed_ms / 1000.0) / 1e9
    ach_x = xb / (extract_ms / 1000.0) / 1e9
    ach_rt = (eb + xb) / ((embed_ms + extract_ms) / 1000.0) / 1e9
    prof = recorded_profile()
    traffic = prof["embed_dram_bytes"] / prof["frames"] * F * (H * W) / float(prof.get("pixels_per_frame", 1080 * 1920)) \
        if prof.get("frames") and prof.get("embed_dram_bytes") else None
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    lane_ops = F * wl.blocks * fp32_lane_ops_per_block(NUM_AC)
    fp32_peak = 148 * 128 * sm_mhz * 1e6
    np_pairs = 4 if NUM_AC >= 48 else (NUM_AC + 16) // 16
    line = {
        "metric": wl.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": wl.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.describe(), "frames_per_gpu": F, "height": H, "width": W, "num_ac": NUM_AC,
                   "delta": DELTA, "step": "embed (BGR->gray stego) + extract (gray stego->packed bits)"
                                           + {"none": "", "fused": " + all-gather of the bits fused into the extract kernel (stores to every rank's buffer over NVLink) + symmetric-memory barrier",
                                              "push": " + all-gather of the bits by DMA into every rank's symmetric buffer (copy engines, overlapped with the next batch) + symmetric-memory barrier",
                                              "nccl": " + NCCL all-gather of bits (chunked, overlapped on a side stream)",
                                              "nccl-seq": " + NCCL all-gather of bits"}[gather_mode],
                   "l2": "inputs (%.1f GB per step) far exceed the 126 MB L2; no flush needed" % ((eb + xb) / 1e9),
                   "parallelism": "frame-sharded x%d" % world,
                   "host": "rank bound to the %d CPUs local to its GPU" % numa_cpus if numa_cpus else "no CPU binding"},
        "mpixel_per_s": value * H * W / 1e6,
        "roofline": {"bound": "hbm", "kernel": "blk::embed_blk_kernel<3,1,%s,false>" % ("true" if NUM_AC >= 63 else "false"),
                     "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": eb, "launch_ms": embed_ms},
        "roofline_extract": {"bound": "hbm", "kernel": "blk::extract_blk_kernel<1,%d>" % np_pairs, "achieved": ach_x, "peak": peak,
                             "unit": "GB/s", "frac": ach_x / peak, "algorithmic_bytes_per_launch": xb,
                             "launch_ms": extract_ms},
        "roofline_round_trip": {"achieved": ach_rt, "peak": peak, "unit": "GB/s", "frac": ach_rt / peak},
        # the co-limiters the HBM fraction has to be read with: reproducing
        # scipy's float32 DCT bit for bit takes 54 un-fused FP32 operations per 8-point transform, and
        # every packed add reads two register pairs - the register file delivers one 32-bit operand per
        # bank, lane and clock (profiles/microbench/coissue.cu, profiles/rf_model.py)
        "roofline_fp32_pipe": {"bound": "fp32 issue (secondary; not an HBM or tensor bound)", "kernel": "blk::embed_blk_kernel",
                               "achieved": lane_ops / (embed_ms / 1000.0) / 1e12, "peak": fp32_peak / 1e12, "unit": "T lane-op/s",
                               "frac": lane_ops / (embed_ms / 1000.0) / fp32_peak,
                               "lane_ops_per_block": fp32_lane_ops_per_block(NUM_AC)},
        "roofline_operand_delivery": None if not prof.get("embed_rf_cycles_per_32_blocks") else {
            "bound": "register-file operand delivery (2 banks x one 32-bit read per lane and clock; secondary). "
                     "frac = modelled / measured cycles: ~1 means the kernel runs at the modelled operand-delivery rate "
                     "(the model is good to a few per cent, so slightly above 1 is possible)",
            "model_cycles_per_32_blocks": prof["embed_rf_cycles_per_32_blocks"],
            "measured_cycles_per_32_blocks": embed_ms / 1000.0 * sm_mhz * 1e6 / (F * wl.blocks / 32.0 / (148 * 4)),
            "frac": prof["embed_rf_cycles_per_32_blocks"] / (embed_ms / 1000.0 * sm_mhz * 1e6 / (F * wl.blocks / 32.0 / (148 * 4))),
            "source": "profiles/rf_model.py over the committed ncu capture (1080p / 63 AC)"} if (H, NUM_AC) == (1080, 63) else None,
        "allgather": None if world == 1 else (
            "fused into extract_kernel: %s over NVLink into symmetric memory, barrier after each launch" % fused.mode
            if fused is not None else
            "%s; overlaps the next batch, all complete inside the timed region" % overlap.mode if gather_mode == "push" else
            "NCCL, overlapped: %d chunks on a side stream, %s SMs left to NCCL, all complete inside the timed region"
            % (len(overlap.bounds), os.environ.get("SVS_RESERVED_SMS", "8")) if overlap is not None
            else "NCCL, sequential on the compute stream"),
        "phase_ms": {"embed": embed_ms, "extract_incl_exchange_launch": extract_ms,
                     "note": "the exchange of step i overlaps embed of step i+1; the last one is inside ms_per_step only"},
        "gpu_launches": launches, "clocks": clocks, "parity_check": ok, "parity": parity, "parity_detail": parity_detail,
        "e2e": e2e,
    }
    if world == 1 and not args.no_cpu:
        try:
            # ~25 s of CPU work: a strip of one frame per host core through the reference function
            rows = max(8, int(H * min(1.0, 25.0 / (13.0 * (H * W) / (1080.0 * 1920.0) * (0.35 + 0.65 * min(NUM_AC, 63) / 63.0)))) // 8 * 8)
            line["cpu_baseline"] = cpu_reference_baseline(wl, rows=rows)
            line["cpu_baseline_c_oracle"] = cpu_c_oracle_rate(wl)
        except Exception as exc:                       # the GPU numbers stay valid without it
            line["cpu_baseline"] = {"error": repr(exc)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        raise SystemExit("parity check failed: %r %r" % (parity, parity_detail))


def run_e2e(args, wl, svs_b200, torch, dist, frames, payload, dev, local_rank, world):
    """Same round trip through svs_*_frames_host: pinned host buffers, copies inside the timing."""
    import ctypes
    L = svs_b200.lib()
    H, W, NUM_AC, cap, nbytes = wl.H, wl.W, wl.num_ac, wl.cap, wl.nbytes
    # a bounded batch per step (default 600 x 1080p = 5.1 GB of pinned host memory per rank): the
    # link-bound rate does not depend on the batch size, the host memory of an 8-rank box does
    limit = int(os.environ.get("SVS_BENCH_E2E_FRAMES", "600"))
    F = max(1, min(frames.shape[0], int(limit * (1080 * 1920) / float(H * W))))
    pay_len = (F * cap + 7) // 8
    frames = frames[:F]
    steps = max(1, min(args.steps, int(os.environ.get("SVS_BENCH_E2E_STEPS", "3"))))
    try:
        h_frames = torch.empty(frames.shape, dtype=torch.uint8).pin_memory()
        h_frames.copy_(frames)
        h_payload = torch.zeros(pay_len + 16, dtype=torch.uint8).pin_memory()
        h_payload[:pay_len] = payload[:pay_len].cpu()
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
    ok = bool(torch.equal(h_bits.reshape(-1)[:pay_len], h_payload[:pay_len]))
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
