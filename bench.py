#!/usr/bin/env python
"""bench.py -- decoded frames/s of the MJPEG423 hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 1080p|4k|4k-q1|480p|1080p-8192]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ... [--gather]
    python bench.py --impl reference ...      # the reference's own C functions on the host cores
    python bench.py --workload enc-1080p      # the encoder (own metric: encoded frames/sec)

One "step" = one pass of the hot path (entropy decode -> dequantise -> IDCT -> YCbCr->BGRA) over the
rank's whole batch of frames.  `value` is timed with the compressed stream already resident in HBM
(CUDA events on the library's streams, max over ranks); `e2e` goes through the public host-buffer call
(mjpeg423_b200_decode_frames): pinned .mpg in, pinned BGRA frames out, H2D and D2H inside the timed
region.  Frames are independent intra frames, so ranks shard by frame range with no collective
("weak": every rank decodes its own `frames` frames; the 1080p-8192 workload is the fixed-total variant).
--gather additionally moves every rank's frames into ONE contiguous buffer on rank 0 (NCCL point-to-point over
NVLink, SURVEY.md 8e "optional single contiguous output"), timed on its own and reported under "gather".
The JSON line also carries "roofline" (dominant kernel: the ALGORITHMIC bytes of SURVEY.md 8d, C + 4P per frame, over
its measured time vs the measured HBM peak; the symbol lists and the block index it really moves are reported as
"intermediate_bytes"), "stages" (every kernel with the bytes it moves), "stages_staged" (the IDCT / colour stage kernels
of the staged modes: 9P / 7P / 10P), "extra" (the other BASELINE configurations, each verified on its own: 4K dense =
configs[3]; the 8192-frame 1080p batch sharded over the ranks = configs[4], strong scaling), "host_link" (plain pinned
D2H copy bandwidth, all ranks at once: the ceiling of e2e), "cpu_baseline" (N=1: the compiled reference, one pinned
process per host core, >= 3 s), "clocks" (nvidia-smi samples inside the timed region) and "gpu_launches".

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (W, H, frames per GPU (weak) or total (strong), unique pictures, noise amplitude, quant, strong?)
    "480p": (640, 480, 4800, 64, 16, "default", False),         # BASELINE configs[0]/[1] content, longer
    "1080p": (1920, 1080, 2000, 2000, 16, "default", False),    # BASELINE configs[2]: the headline config; every frame
                                                                # its own procedural picture (noise seeded per frame)
    "4k": (3840, 2160, 256, 16, 256, "default", False),         # BASELINE configs[3]: dense, entropy-bound
    "4k-q1": (3840, 2160, 64, 8, 256, "ones", False),           # configs[3] extreme: all-ones quant tables
    "1080p-8192": (1920, 1080, 8192, 64, 16, "default", True),  # BASELINE configs[4]: fixed total, sharded
    # SURVEY.md 8f3 (next row): the ENCODER, frames -> .mpg.  Not a BASELINE config; reported with its own metric.
    "enc-1080p": (1920, 1080, 256, 32, 16, "default", False),
    "enc-480p": (640, 480, 1024, 32, 16, "default", False),
}
HBM_FALLBACK_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def bind_near_gpu(dev: int) -> dict | None:
    """Multi-GPU boxes: run this rank on the CPUs next to its GPU (sysfs `local_cpulist` of the GPU's PCI function) so
    that the pinned host buffers it allocates afterwards land on that NUMA node -- with eight ranks copying 50 GB/s
    each, buffers on the far socket cap the box.  No effect where the GPU is local to every CPU (single-node VMs).
    MJPEG423_BENCH_NO_BIND=1 turns it off."""
    if os.environ.get("MJPEG423_BENCH_NO_BIND"):
        return None
    try:
        import torch
        pr = torch.cuda.get_device_properties(dev)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        text = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        cpus = set()
        for part in text.split(","):
            if part:
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        have = os.sched_getaffinity(0)
        cpus &= have
        if not cpus or cpus == have:
            return {"numa_node": node, "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "bound": True, "cpus": len(cpus)}
    except Exception as e:           # no sysfs / old torch: run unbound
        return {"bound": False, "why": str(e)[:80]}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU, sampled every 50 ms.  The sampler is started EARLY (nvidia-smi
    needs about a second to enumerate an 8-GPU box) and reads its lines with their arrival time; begin()/end() mark the
    timed region and stop() reports the samples that arrived inside it."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []                 # (arrival time, text)
        self.t_begin = self.t_end = None

    def start(self):
        import threading
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.lines.append((time.perf_counter(), line))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def begin(self):
        self.t_begin = time.perf_counter()

    def end(self):
        self.t_end = time.perf_counter()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.t_end is None:
            self.end()
        time.sleep(0.06)                # let the sample that covers the end of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        t0 = self.t_begin if self.t_begin is not None else 0.0
        inside = [l for t, l in self.lines if t0 <= t <= self.t_end + 0.06]
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in inside:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous frame range of `rank` when `total` frames are split over `world` ranks."""
    return total * rank // world, total * (rank + 1) // world


def reduce_over_ranks(dist, maxima, sums, device):
    """MAX-reduce the timings and SUM-reduce the counts over all ranks (no-op for a single process)."""
    import torch
    t = torch.tensor(list(maxima), dtype=torch.float64, device=device)
    c = torch.tensor(list(sums), dtype=torch.float64, device=device)
    if dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()], [float(x) for x in c.tolist()]


def gather_frames(dist, mine, counts, frame_bytes: int, root: int = 0):
    """SURVEY.md 8e, the OPTIONAL single contiguous output: every rank's slice (`mine`: a flat uint8 tensor of
    counts[rank] * frame_bytes) is sent to `root`, which receives them into ONE buffer in frame order -- grouped
    point-to-point transfers (ncclSend / ncclRecv over NVLink with the nccl backend; the slices differ in length, so it
    is not an all-gather).  Returns the contiguous tensor on root, None elsewhere."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank != root:
        reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, mine, root)])
        for r in reqs:
            r.wait()
        return None
    full = torch.empty(int(sum(counts)) * frame_bytes, dtype=torch.uint8, device=mine.device)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64) * frame_bytes
    ops = [dist.P2POp(dist.irecv, full[int(offs[r]):int(offs[r + 1])], r) for r in range(world) if r != root and counts[r]]
    full[int(offs[root]):int(offs[root + 1])].copy_(mine)
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    return full


def make_stream(wl: str, frames: int, nthreads: int):
    from mjpeg423_b200 import synth
    W, H, _, uniq, amp, quant, _ = WORKLOADS[wl]
    q = np.ones(64, np.int16) if quant == "ones" else None
    uniq = min(uniq, frames)
    mpg = synth.synth_mpg(W, H, frames, uniq, amp, 0, q, q, nthreads=nthreads)
    return mpg, uniq, q


def workload_config(wl: str, frames: int, world: int) -> dict:
    """The keys that NAME the workload: identical in the b200 arm and the reference arm."""
    W, H, _, uniq, amp, quant, strong = WORKLOADS[wl]
    return {"workload": wl, "width": W, "height": H, "frames_total" if strong else "frames_per_gpu": frames,
            "unique_pictures": min(uniq, frames), "noise_amp": amp, "quant": quant,
            "l2": "inputs larger than L2 (no flush needed)", "parallelism": f"frame-range x{world}, no collective"}


def _cpu_worker(job):
    """One host process of the CPU baseline: pinned to one core, decodes its contiguous frame range `reps` times."""
    cpu, lo, hi, reps, quant_ones, barrier, mpg = job
    try:
        os.sched_setaffinity(0, {cpu})
    except Exception:
        pass
    from oracle import oracle
    chk = oracle.best()
    q = np.ones(64, np.int16) if quant_ones else None
    secs = np.zeros(3)
    chk.decode_mpg(mpg, lo, 1, yq=q, cq=q)                 # page the library and the stream in
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(reps):
        chk.decode_mpg(mpg, lo, hi - lo, yq=q, cq=q, nthreads=1, stage_secs=secs)
    return t0, time.perf_counter(), (hi - lo) * reps, secs


_CPU_JOBS = None


def _cpu_worker_idx(k):
    return _cpu_worker(_CPU_JOBS[k])


def cpu_reference_fps(mpg: np.ndarray, n_frames: int, quant_ones: bool, min_seconds: float, per_core_fps: float):
    """the reference decode functions, one PROCESS per host core, each pinned
    (sched_setaffinity) and given a contiguous frame range of the same in-memory stream; wall time from the first start
    to the last finish.  Must run before this process touches CUDA (the workers are forked)."""
    global _CPU_JOBS
    import multiprocessing as mp
    cpus = sorted(os.sched_getaffinity(0))
    cores = len(cpus)
    per = max(1, n_frames // cores)
    reps = max(1, int(np.ceil(min_seconds * per_core_fps / per)))
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(cores)
    _CPU_JOBS = [(cpus[k], k * per, (k + 1) * per, reps, quant_ones, barrier, mpg) for k in range(cores)]   # inherited by fork
    procs_out = ctx.Queue()

    def run(k):
        procs_out.put(_cpu_worker(_CPU_JOBS[k]))
    procs = [ctx.Process(target=run, args=(k,)) for k in range(cores)]
    for pr in procs:
        pr.start()
    res = [procs_out.get() for _ in procs]
    for pr in procs:
        pr.join()
    t0, t1 = min(r[0] for r in res), max(r[1] for r in res)
    frames = sum(r[2] for r in res)
    secs = np.sum([r[3] for r in res], axis=0)
    from oracle import oracle
    return {"value": frames / (t1 - t0), "unit": "frames/s", "cores": cores, "kind": oracle.best().kind,
            "sample": f"{cores} pinned processes x {per} frames x {reps} passes of the same stream "
                      f"(contiguous frame range each), {t1 - t0:.2f} s wall",
            "per_core": frames / (t1 - t0) / cores,
            "stage_share": {k: float(v / secs.sum()) for k, v in zip(("entropy", "idct", "colour"), secs)}}


PER_CORE_FPS = {"480p": 250.0, "1080p": 40.0, "4k": 9.0, "4k-q1": 6.0, "1080p-8192": 40.0}   # measured on the B200 boxes' hosts


def reference_arm(args, rank: int, world: int):
    """The reference's own lossless_decode/idct/ycbcr_to_rgb (oracle/_ref, compiled from /root/reference)
    or, if that did not travel, the C restatement -- on all host cores, bounded sample per step."""
    if rank != 0:
        return
    W, H, frames, uniq, amp, quant, _ = WORKLOADS[args.workload]
    if args.frames:
        frames = args.frames
    cores = len(os.sched_getaffinity(0))
    per = max(2, int(np.ceil(PER_CORE_FPS[args.workload])))                 # frames per process and pass (~1 s)
    n = per * cores
    mpg, uniq_n, q = make_stream(args.workload, n, cores)
    runs = []
    for i in range(args.warmup + args.steps):                   # a step = every core decoding its range for ref_seconds
        r = cpu_reference_fps(mpg, n, q is not None, args.ref_seconds, PER_CORE_FPS[args.workload])
        if i >= args.warmup:
            runs.append(r)
    fps = float(np.mean([r["value"] for r in runs]))
    cb = dict(runs[-1])
    cb["value"] = fps
    line = {
        "impl": "reference", "metric": "decoded frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean([n / r["value"] for r in runs])) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args.workload, frames, args.gpus),
        "details": {"host": "cpu", "sample_frames": n, "sample_unique_pictures": uniq_n},
        "cpu_baseline": cb,
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def procedural_frames(W: int, H: int, n: int, uniq: int, amp: int) -> np.ndarray:
    """(n, H, W, 4) BGRA: `uniq` procedural pictures (gradients + LCG-free numpy noise) cycled."""
    rng = np.random.default_rng(0x423)
    y, x = np.mgrid[0:H, 0:W]
    base = np.zeros((uniq, H, W, 4), np.uint8)
    for f in range(uniq):
        g = np.stack([(255 * (x + y) // (W + H) + f) & 255, (255 * y // H + f) & 255, (255 * x // W + f) & 255], -1)
        base[f, ..., :3] = (g + rng.integers(0, max(amp, 1), size=(H, W, 3))) & 255
    return base[np.arange(n) % uniq]


def _encode_worker(job):
    """One host process of the encoder's reference arm: encode a `per`-frame clip `reps` times."""
    W, H, per, amp, max_i, reps = job
    from oracle import oracle
    chk = oracle.best()
    clip = procedural_frames(W, H, per, per, amp)
    for _ in range(reps):
        chk.encode_mpg(clip, max_i)
    return reps


def encoder_bench(args, rank: int, world: int, local_rank: int):
    """encoded frames/s of mjpeg423_b200_encode_frames (frames -> complete .mpg in host memory).
    `value`: frames already resident in HBM; `e2e`: pinned host frames in.  --impl reference: the reference's own
    encoder functions (oracle/_ref frame loop, else the restatement), one clip per host core."""
    W, H, frames, uniq, amp, _, _ = WORKLOADS[args.workload]
    if args.frames:
        frames = args.frames
    cores = os.cpu_count() or 1
    max_i = 24                                   # COMMON/config.h:54
    if args.impl == "reference":
        if rank != 0:
            return
        from concurrent.futures import ProcessPoolExecutor
        from oracle import oracle
        chk = oracle.best()
        per = 4 if H >= 1080 else 16
        # one PROCESS per core: the reference's lossless_encode keeps its bit buffer in file-scope globals
        # (LIB/encoder/lossless_encode.c:17-19) and is not thread-safe
        with ProcessPoolExecutor(max_workers=cores) as pool:
            list(pool.map(_encode_worker, [(W, H, per, amp, max_i, 0)] * cores))        # start-up + warm-up
            t0 = time.perf_counter()
            for _ in range(args.steps):
                list(pool.map(_encode_worker, [(W, H, per, amp, max_i, 1)] * cores))
            dt = time.perf_counter() - t0
        fps = per * cores * args.steps / dt
        print(json.dumps({"impl": "reference", "metric": "encoded frames/sec", "value": fps, "unit": "frames/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+int32",
                          "data": "synthetic", "config": {"workload": args.workload, "width": W, "height": H, "max_I_interval": max_i},
                          "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": chk.kind,
                                           "sample": f"{cores} processes x {per}-frame clip per step"},
                          "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return
    import torch
    import torch.distributed as dist
    import mjpeg423_b200
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    fr = procedural_frames(W, H, frames, uniq, amp)
    dec = mjpeg423_b200.Decoder(local_rank)
    sampler = ClockSampler(local_rank)
    sampler.start()
    d = dec.device_alloc(fr.nbytes)
    dec.to_device(d, fr)
    mpg = dec.encode_frames(None, max_i, d_frames=d, shape=(frames, H, W)).copy()
    pin_out = dec.pinned(mpg.size + (64 << 20))            # the .mpg lands in pinned host memory
    verified = None
    if not args.no_verify:
        from oracle import oracle
        chk = oracle.best()
        k = 6
        want = chk.encode_mpg(fr[:k], max_i)
        body = int(want[16:20].view("<u4")[0])           # payload bytes of the k-frame file = prefix of the long one
        if not np.array_equal(mpg[20:20 + body], want[20:20 + body]):
            raise SystemExit("encoder output differs from the oracle -- refusing to report a number")
        verified = f"frame records of the first {k} frames byte-identical to the {chk.kind} encoder"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        dec.encode_frames(None, max_i, d_frames=d, shape=(frames, H, W), out=pin_out)
    barrier()
    sampler.begin()
    ev_ms, launches = 0.0, 0
    for _ in range(args.steps):
        dec.encode_frames(None, max_i, d_frames=d, shape=(frames, H, W), out=pin_out)
        st = dec.stats()
        ev_ms += st["total_ms"]
        launches += st["kernel_launches"]
    barrier()
    sampler.end()
    clocks = sampler.stop()
    pin = dec.pinned(fr.nbytes)
    pin.array[:] = fr.reshape(-1)
    host_frames = pin.array.reshape(fr.shape)
    dec.encode_frames(host_frames, max_i, out=pin_out)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        out = dec.encode_frames(host_frames, max_i, out=pin_out)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    (ev_ms_max, e2e_s_max), (total_frames, total_launches) = reduce_over_ranks(dist, [ev_ms, e2e_s], [frames, launches], "cuda")
    if rank == 0:
        peak, peak_src = measured_peak()
        fps = total_frames * args.steps / (ev_ms_max / 1e3)
        P = W * H
        # algorithmic bytes per frame: read 4P pixels, write and re-read 6P of levels (size + emit passes), write C
        alg = 4 * P + 3 * 6 * P + mpg.size / frames
        line = {"metric": "encoded frames/sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64+int32", "data": "synthetic",
                "config": {"workload": args.workload, "width": W, "height": H, "frames_per_gpu": frames, "unique_pictures": uniq,
                           "noise_amp": amp, "max_I_interval": max_i, "compressed_bytes_per_frame": mpg.size / frames,
                           "p_frames": int(mjpeg423_b200.probe(mpg).num_pframes), "verified": verified,
                           "l2": "inputs larger than L2 (%.1f GB of frames per step)" % (fr.nbytes / 1e9),
                           "timing": "cuda events (whole call incl. the .mpg read-back), max over ranks"},
                "clocks": clocks,
                "e2e": {"value": total_frames * e2e_steps / e2e_s_max, "unit": "frames/s", "h2d_bytes_per_step": int(fr.nbytes),
                        "d2h_bytes_per_step": int(out.size), "api": "mjpeg423_b200_encode_frames (pinned host frames in, pinned host .mpg out)"},
                "gpu_launches": int(total_launches),
                "roofline": {"bound": "hbm", "kernel": "encoder pipeline (k_enc_transform/size/scan/emit)", "achieved": fps / world * alg / 1e9,
                             "peak": peak, "unit": "GB/s", "frac": fps / world * alg / 1e9 / peak, "traffic": None,
                             "peak_source": peak_src, "algorithmic_bytes_per_frame": alg},
                "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    pin.free()
    pin_out.free()
    dec.device_free(d)
    dec.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_config(dec, wl: str, my_frames: int, steps: int, warmup: int, verify: bool, gen_threads: int, barrier):
    """One extra BASELINE configuration on this rank: stream resident in HBM, every frame verified through the 64-bit
    position-mixed checksum against the oracle, then `steps` timed decode_resident() calls (CUDA events of the library)."""
    from mjpeg423_b200 import api
    W, H = WORKLOADS[wl][0], WORKLOADS[wl][1]
    mpg, uniq, q = make_stream(wl, my_frames, gen_threads)
    frame_bytes = W * H * 4
    if q is not None:
        dec.set_quant(q, q)
    dec.upload(mpg)
    d_out = dec.device_alloc(my_frames * frame_bytes)
    verified = None
    try:
        if verify:
            from oracle import oracle
            chk = oracle.best()
            dec.decode_resident(d_out)
            got = dec.hash_frames(d_out, frame_bytes, my_frames)
            want = api.frame_hash_host(chk.decode_mpg(mpg, 0, uniq, yq=q, cq=q, nthreads=gen_threads))[np.arange(my_frames) % uniq]
            if not np.array_equal(got, want):
                raise SystemExit(f"{wl}: frame {int(np.flatnonzero(got != want)[0])} differs from the {chk.kind} oracle")
            verified = f"{my_frames} frames/rank: per-frame 64-bit checksum == {chk.kind} oracle"
        for _ in range(warmup):
            dec.decode_resident(d_out)
        barrier()
        ev_ms = 0.0
        for _ in range(steps):
            dec.decode_resident(d_out)
            ev_ms += dec.stats()["total_ms"]
        barrier()
        st = dec.stats()
    finally:
        dec.device_free(d_out)
        if q is not None:
            dec.set_quant(None, None)
    return {"ev_ms": ev_ms, "frames": my_frames, "verified": verified, "payload_bytes": st["payload_bytes"],
            "fixups": st["fixups"], "W": W, "H": H}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="override frames per GPU (weak) / total (strong)")
    ap.add_argument("--e2e-frames", type=int, default=0, help="frames per end-to-end step (0 = auto, ~4 GB of output)")
    ap.add_argument("--ref-seconds", type=float, default=3.0, help="wall seconds per reference-arm step")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU-seconds of work for the cpu_baseline sample")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra BASELINE configurations (4K dense, 8192-frame strong)")
    ap.add_argument("--gather", action="store_true",
                    help="multi-GPU: also gather every rank's frames into one contiguous buffer on rank 0 over NCCL (timed separately)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload.startswith("enc-"):
        encoder_bench(args, rank, world, local_rank)
        return
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    W, H, frames, uniq, amp, quant, strong = WORKLOADS[args.workload]
    if args.frames:
        frames = args.frames
    cores = os.cpu_count() or 1
    # ---- CPU baseline first: it forks one pinned process per core, which must happen before CUDA is initialised ----
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        per = max(2, int(np.ceil(PER_CORE_FPS[args.workload])))
        ncores = len(os.sched_getaffinity(0))
        mpg_cpu, _, q_cpu = make_stream(args.workload, per * ncores, ncores)
        cpu_baseline = cpu_reference_fps(mpg_cpu, per * ncores, q_cpu is not None, max(3.0, args.cpu_seconds / 4), PER_CORE_FPS[args.workload])
        del mpg_cpu

    import torch
    import torch.distributed as dist

    import mjpeg423_b200
    from mjpeg423_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the host arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if strong:
        lo, hi = shard_range(frames, rank, world)
        my_frames = hi - lo
    else:
        my_frames = frames
    binding = bind_near_gpu(local_rank) if world > 1 else None
    mpg, uniq, q = make_stream(args.workload, my_frames, max(1, cores // world))
    frame_bytes = W * H * 4
    P = W * H

    dec = mjpeg423_b200.Decoder(local_rank)
    sampler = ClockSampler(local_rank)
    sampler.start()                          # early: nvidia-smi takes a while to deliver its first line
    if q is not None:
        dec.set_quant(q, q)
    info = dec.upload(mpg)
    out_t = None
    if args.gather and world > 1:
        out_t = torch.empty(my_frames * frame_bytes, dtype=torch.uint8, device="cuda")   # NCCL needs a torch tensor
        d_out = out_t.data_ptr()
    else:
        d_out = dec.device_alloc(my_frames * frame_bytes)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness gate (untimed): every frame of this rank against the CPU oracle ------------------
    verified = None
    if not args.no_verify:
        from oracle import oracle
        chk = oracle.best()
        dec.decode_resident(d_out)
        got = dec.hash_frames(d_out, frame_bytes, my_frames)
        want_u = np.concatenate([api.frame_hash_host(chk.decode_mpg(mpg, i, min(64, uniq - i), yq=q, cq=q, nthreads=max(1, cores // world)))
                                 for i in range(0, uniq, 64)])          # (in batches: 2000 decoded 1080p frames are 16.6 GB)
        want = want_u[np.arange(my_frames) % uniq]
        if not np.array_equal(got, want):
            bad = int(np.flatnonzero(got != want)[0])
            raise SystemExit(f"rank {rank}: frame {bad} differs from the {chk.kind} oracle -- refusing to report a number")
        first = dec.to_host(d_out, frame_bytes).reshape(H, W, 4)
        assert np.array_equal(first, chk.decode_mpg(mpg, 0, 1, yq=q, cq=q)[0])
        verified = f"{my_frames} frames/rank: per-frame 64-bit checksum == {chk.kind} oracle; frame 0 byte-compared"

    # ---- device-resident throughput -------------------------------------------------------------------------
    for _ in range(args.warmup):
        dec.decode_resident(d_out)
    barrier()
    sampler.begin()
    ev_ms, launches = 0.0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dec.decode_resident(d_out)
        st = dec.stats()
        ev_ms += st["total_ms"]
        launches += st["kernel_launches"]
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    sampler.end()
    clocks = sampler.stop()
    payload_bytes = st["payload_bytes"]

    # ---- per-stage times (profiled step: events between the kernels, stages serialised) ---------------------
    dec.set_option(api.OPT_PROFILE, 1)
    dec.decode_resident(d_out)
    dec.decode_resident(d_out)
    ps = dec.stats()
    dec.set_option(api.OPT_PROFILE, 0)

    # ---- end to end: pinned host .mpg -> pinned host frames through the public call ----------------------------
    e2e_frames = args.e2e_frames or int(max(1, min(my_frames, (4 << 30) // frame_bytes)))
    fr_off = 20
    for _ in range(e2e_frames):
        fr_off += int(mpg[fr_off:fr_off + 4].view("<u4")[0])
    h2d_bytes = fr_off - 20
    pin_in = dec.pinned(mpg.size + 64)
    pin_in.array[:mpg.size] = mpg
    pin_out = dec.pinned(e2e_frames * frame_bytes)
    e2e_steps = max(2, min(args.steps, 5))
    dec.decode_frames(pin_in.array[:mpg.size], 0, e2e_frames, out=pin_out)     # warm-up (allocates rings)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dec.decode_frames(pin_in.array[:mpg.size], 0, e2e_frames, out=pin_out)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if not args.no_verify:
        tail = pin_out.array[-frame_bytes:].reshape(1, H, W, 4)
        assert api.frame_hash_host(tail)[0] == want[e2e_frames - 1], "end-to-end output differs from the oracle"

    # ---- optional: one contiguous output on rank 0 (NCCL point-to-point over NVLink), timed on its own --------------------
    gather = None
    if out_t is not None:
        cnt_t = torch.zeros(world, dtype=torch.int64, device="cuda")
        cnt_t[rank] = my_frames
        dist.all_reduce(cnt_t)
        counts = [int(x) for x in cnt_t.tolist()]
        if sum(counts) * frame_bytes > 100e9:
            raise SystemExit("--gather: the contiguous output would not fit one GPU; use the 1080p-8192 workload")
        full = gather_frames(dist, out_t, counts, frame_bytes)        # warm-up (allocates, connects)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(3):
            full = gather_frames(dist, out_t, counts, frame_bytes)
        g1.record()
        barrier()
        g_ms = torch.tensor([g0.elapsed_time(g1) / 3], dtype=torch.float64, device="cuda")
        dist.all_reduce(g_ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            ok = None
            if not args.no_verify:
                lo0 = 0
                ok = True
                for r, c in enumerate(counts):        # every rank decoded frames [0, c) of its own copy of the clip cycle
                    got_r = dec.hash_frames(full.data_ptr() + lo0 * frame_bytes, frame_bytes, c)
                    ok = ok and bool(np.array_equal(got_r, want_u[np.arange(c) % uniq]))
                    lo0 += c
                if not ok:
                    raise SystemExit("gathered output differs from the oracle")
            moved = (sum(counts) - counts[0]) * frame_bytes
            gather = {"ms": float(g_ms.item()), "bytes": int(moved), "GBs": moved / float(g_ms.item()) / 1e6,
                      "how": "grouped ncclSend/ncclRecv (torch.distributed.batch_isend_irecv) into one buffer on rank 0",
                      "verified": ok}
        del full

    # ---- the stage kernels of the staged modes (IDCT 9P, colour 7P; fused IDCT + colour 10P) --------------
    n_st = int(min(my_frames, 256))
    staged = {}
    dec.upload(mpg, 0, n_st)
    dec.set_option(api.OPT_PROFILE, 1)
    for mode in (2, 1):
        dec.set_option(api.OPT_STAGED, mode)
        dec.decode_resident(d_out)
        dec.decode_resident(d_out)
        st2 = dec.stats()
        if mode == 2:
            staged["k_idct"] = {"ms": st2["idct_ms"], "bytes": 9 * P * n_st}
            staged["k_colour"] = {"ms": st2["colour_ms"], "bytes": 7 * P * n_st}
            staged["k_decode_coef"] = {"ms": st2["decode_ms"], "bytes": 6 * P * n_st + 4 * st2["list_entries"]}
        else:
            staged["k_idct_colour"] = {"ms": st2["idct_colour_ms"], "bytes": 10 * P * n_st}
    dec.set_option(api.OPT_STAGED, 0)
    dec.set_option(api.OPT_PROFILE, 0)
    if not args.no_verify:
        assert np.array_equal(dec.hash_frames(d_out, frame_bytes, n_st), want[:n_st]), "staged modes differ from the oracle"

    # ---- the host link: plain device -> pinned-host copies, every rank at once (what bounds e2e) ----------------------
    link_bytes = 1 << 30
    lt_d = torch.empty(link_bytes, dtype=torch.uint8, device="cuda")
    lt_h = torch.empty(link_bytes, dtype=torch.uint8).pin_memory()
    lt_h.copy_(lt_d, non_blocking=True)
    barrier()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    for _ in range(4):
        lt_h.copy_(lt_d, non_blocking=True)
    l1.record()
    barrier()
    link_ms = l0.elapsed_time(l1) / 4
    del lt_d, lt_h

    # ---- the other BASELINE configurations, each verified on its own ------------------------------------------------
    extras = {}
    if not args.no_extra and args.workload == "1080p" and not (args.gather and world > 1):
        dec.device_free(d_out)
        d_out = None
        gen_threads = max(1, cores // world)
        extras["4k"] = run_config(dec, "4k", 128, 5, 2, not args.no_verify, gen_threads, barrier)           # configs[3], weak
        lo8, hi8 = shard_range(WORKLOADS["1080p-8192"][2], rank, world)
        extras["1080p-8192"] = run_config(dec, "1080p-8192", hi8 - lo8, 3, 1, not args.no_verify, gen_threads, barrier)   # configs[4]

    # ---- reduce over ranks ----------------------------------------------------------------------------------------
    (ev_ms_max, wall_ms_max, e2e_s_max, link_ms_max, x4k_ms, x8k_ms), (total_frames, total_launches, total_e2e_frames, x4k_fr, x8k_fr) = reduce_over_ranks(
        dist, [ev_ms, wall_ms, e2e_s, link_ms, extras.get("4k", {}).get("ev_ms", 0.0), extras.get("1080p-8192", {}).get("ev_ms", 0.0)],
        [my_frames, launches, e2e_frames, extras.get("4k", {}).get("frames", 0), extras.get("1080p-8192", {}).get("frames", 0)], "cuda")

    if rank == 0:
        peak, peak_src = measured_peak()
        fps = total_frames * args.steps / (ev_ms_max / 1e3)
        Cbar = payload_bytes / my_frames
        n = my_frames
        blocks = (3 * P // 64) * n
        lists = 4 * ps["list_entries"]
        stage = {
            # ALGORITHMIC bytes per step: C = compressed bytes, P = pixels per frame; symbol lists are
            # 4 bytes per coded coefficient, the block index 8 bytes per block, per-segment state 16 bytes
            "entropy_sync": {"ms": ps["sync_ms"], "bytes": payload_bytes + 12 * ps["segments"], "kernels": "k_entropy_sync"},
            "entropy_chain": {"ms": ps["chain_ms"], "bytes": 16 * ps["segments"], "kernels": "k_entropy_chain"},
            "entropy_index": {"ms": ps["index_ms"], "bytes": payload_bytes + lists + 8 * blocks + 20 * ps["segments"],
                              "kernels": "k_entropy_index"},      # (+ the tiny k_entropy_dcscan, timed with it)
            "decode_fused": {"ms": ps["decode_ms"], "bytes": lists + 8 * blocks + 4 * P * n, "kernels": "k_decode_fused"},
        }
        for s in stage.values():
            s["GBs"] = s["bytes"] / max(s["ms"], 1e-9) / 1e6
            s["frac_of_hbm_peak"] = s["GBs"] / peak
        dom = max(stage, key=lambda k: stage[k]["ms"])
        # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (bytes per frame of
        # the same workload, scaled to this step's frames); null if no capture is on file for it.
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                cap = json.load(f).get(args.workload, {}).get(stage[dom]["kernels"])
            if cap:
                traffic = cap["dram_bytes_per_frame"] * n
        except Exception:
            pass
        # The roofline number uses the ALGORITHMIC bytes of SURVEY.md 8d -- read the bitstream once, write BGRA once:
        # (C + 4P) per frame -- over the dominant kernel's measured time; what the kernel really moves (symbol lists and
        # block index on top of its pixels) is reported beside it.
        alg = (Cbar + 4 * P) * n
        ach = alg / max(stage[dom]["ms"], 1e-9) / 1e6
        roofline = {"bound": "hbm", "kernel": stage[dom]["kernels"], "achieved": ach, "peak": peak,
                    "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_step": alg, "algorithmic_bytes_per_frame": Cbar + 4 * P,
                    "intermediate_bytes": {"moved_by_kernel_per_step": stage[dom]["bytes"], "symbol_lists_per_step": lists,
                                           "block_index_per_step": 8 * blocks},
                    "step_ms_in_kernel": stage[dom]["ms"],
                    "launches_per_step": ps["kernel_launches"] // 5,     # one per pipeline chunk
                    "pipeline_headline": {"bytes_per_frame": Cbar + 4 * P, "GBs": fps / world * (Cbar + 4 * P) / 1e9,
                                          "frac": fps / world * (Cbar + 4 * P) / 1e9 / peak}}
        for v in staged.values():
            v["frames"] = n_st
            v["GBs"] = v["bytes"] / max(v["ms"], 1e-9) / 1e6
            v["frac_of_hbm_peak"] = v["GBs"] / peak
        extra = {}
        for key, ms_max, fr_tot, steps_x in (("4k", x4k_ms, x4k_fr, 5), ("1080p-8192", x8k_ms, x8k_fr, 3)):
            if key not in extras:
                continue
            e = extras[key]
            cb = e["payload_bytes"] / e["frames"]
            px = e["W"] * e["H"]
            xfps = fr_tot * steps_x / (ms_max / 1e3)
            extra[key] = {"config": workload_config(key, int(fr_tot) if WORKLOADS[key][6] else e["frames"], world),
                          "metric": "decoded frames/sec", "value": xfps, "unit": "frames/s", "steps": steps_x,
                          "ms_per_step": ms_max / steps_x, "scaling": "strong" if WORKLOADS[key][6] else "weak",
                          "compressed_bytes_per_frame": cb, "bits_per_pixel": 8 * cb / px, "chain_fixups": e["fixups"],
                          "pipeline_headline": {"bytes_per_frame": cb + 4 * px, "GBs": xfps / world * (cb + 4 * px) / 1e9,
                                                "frac": xfps / world * (cb + 4 * px) / 1e9 / peak},
                          "verified": e["verified"]}
        line = {
            "metric": "decoded frames/sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(args.workload, frames, world),
            "details": {"frames_this_rank": my_frames, "compressed_bytes_per_frame": Cbar, "bits_per_pixel": 8 * Cbar / P,
                        "l2": "inputs larger than L2 (bitstream %.0f MB, output %.1f GB per GPU per step)" %
                              (payload_bytes / 1e6, my_frames * frame_bytes / 1e9),
                        "timing": "cuda events, max over ranks", "cpu_binding": binding,
                        "wall_ms_per_step": wall_ms_max / args.steps, "verified": verified},
            "clocks": clocks,
            "e2e": {"value": total_e2e_frames * e2e_steps / e2e_s_max, "unit": "frames/s", "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(e2e_frames * frame_bytes), "frames_per_step": e2e_frames, "steps": e2e_steps,
                    "GBs_d2h": total_e2e_frames / world * e2e_steps * frame_bytes / e2e_s_max / 1e9,
                    "api": "mjpeg423_b200_decode_frames (pinned host in/out)"},
            "host_link": {"what": "plain device -> pinned host copies of 1 GiB, all ranks at once (no kernels)",
                          "d2h_GBs_per_gpu": link_bytes / link_ms_max / 1e6, "d2h_GBs_aggregate": world * link_bytes / link_ms_max / 1e6,
                          "e2e_share_of_link": (total_e2e_frames / world * e2e_steps * frame_bytes / e2e_s_max / 1e9) / (link_bytes / link_ms_max / 1e6)},
            "gpu_launches": int(total_launches),
            "roofline": roofline,
            "stages": stage,
            "stages_staged": staged,
            "extra": extra,
            "segments": {"per_step": ps["segments"], "chain_fixups": ps["fixups"]},
            "cpu_baseline": cpu_baseline,
        }
        if gather is not None:
            line["gather"] = gather
        print(json.dumps(line), flush=True)
    pin_in.free()
    pin_out.free()
    if out_t is None and d_out is not None:
        dec.device_free(d_out)
    dec.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
