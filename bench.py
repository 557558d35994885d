#!/usr/bin/env python
"""Benchmark of the north-star path: ORB extraction of both eyes + Frame::ComputeStereoMatches on
KITTI-shape (1241x376, nFeatures=2000) synthetic stereo frames -- and, on the same JSON line, the three
other workloads BASELINE.json names: configs[2] (TUM frames: extraction + SearchByProjection against a 20k-point map), configs[3]
(batched offline extraction of 32768 frames) and configs[4] (brute-force keyframe-vs-keyframe Hamming matching with the NCCL all-gather of descriptor sets).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl ours|reference]

One "step" of the headline = one batch of F stereo frames per GPU through the whole hot path.  `value` is stereo
frames/s with the images already resident in HBM (CUDA events on the launching stream, max over ranks); the timed
region is R back-to-back repeats of the K steps the caller asked for, R chosen so that it lasts >= 1 s
(`config.timed_steps` = R * K; `ms_per_step` is per step).  `e2e` is the same through the host-buffer C ABI
(obs_stereo_frames_submit / _wait: page-locked host images in, keypoints, descriptors, uRight and depth out), driven
by ONE host thread that keeps three handle pairs in flight.  Multi-GPU: one process per GPU (torchrun), frames
sharded by rank, no data-path collective for configs[1]/[3]; configs[4] all-gathers the descriptor shards over NCCL
in chunks and matches while the gather is in flight.  `sub` carries configs[3] and configs[4] with their own value /
e2e / roofline and a result hash that does not depend on the number of GPUs (checked against the 1-GPU hash
committed in tests/golden/bench_hashes.json).

Only the `cpu_baseline` legs and `--impl reference` touch oracle/ (the reference's own ORBextractor.cc compiled in
place + the stereo matcher) -- as the thing timed beside the product, never inside it.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from object_slam_b200 import synth  # noqa: E402

W, H, NFEAT = 1241, 376, 2000
PITCH = 1280
METRIC = "stereo frames/s, ORB extract (both eyes) + stereo match, KITTI 1241x376, 2000 kp"
WORKLOAD = ("configs[1]: KITTI-shape 1241x376 synthetic stereo pairs, nFeatures=2000, nLevels=8, scale 1.2, FAST 20/7: "
            "ORB extraction of both eyes + Frame::ComputeStereoMatches")
LEVEL_PX = [1241 * 376, 1034 * 313, 862 * 261, 718 * 218, 598 * 181, 499 * 151, 416 * 126, 346 * 105]
MIN_TIMED_S = 1.0
HASH_FILE = os.path.join(ROOT, "tests", "golden", "bench_hashes.json")


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def expected_hash(key):
    try:
        with open(HASH_FILE) as f:
            return json.load(f).get(key)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
            try:
                pw.append(float(f[6]))
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "power_w_max": max(pw) if pw else None}


class Ctx:
    """Rank bookkeeping, process group, device, host CPU affinity: shared by the headline and the sub-workloads."""

    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.torch = self.dist = self.dev = None
        self.all_cpus = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(os.cpu_count() or 1))
        self.affinity_note = "unchanged"

    def init_gpu(self):
        import torch
        import torch.distributed as dist
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: object_slam_b200 has no CPU path")
        self.torch, self.dist = torch, dist
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.pin_near_gpu()
        if self.world > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def pin_near_gpu(self):
        """Run this rank's host threads -- and therefore allocate its page-locked buffers -- on the CPUs NVML reports as local
        to its GPU (the NUMA node of the PCIe root the GPU hangs off), and never on more than cores / ranks of them."""
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = self.torch.cuda.get_device_properties(self.local_rank).uuid
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.local_rank)
            words = (max(self.all_cpus) + 64) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            near = {64 * i + b for i, wd in enumerate(mask) for b in range(64) if (int(wd) >> b) & 1}
            use = sorted(near & set(self.all_cpus))
            if use and self.world > 1:
                # ranks that share a NUMA node split its CPUs
                share = max(len(self.all_cpus) // self.world, 1)
                k = (self.local_rank * share) % max(len(use), 1)
                use = (use[k:] + use[:k])[:max(share, 1)]
            if use:
                os.sched_setaffinity(0, use)
                self.affinity_note = f"{len(use)} CPUs local to GPU {self.local_rank} (NVML affinity): {use[0]}..{use[-1]}"
        except Exception as ex:     # NVML or the affinity call unavailable: keep the inherited mask
            self.affinity_note = f"unchanged ({type(ex).__name__})"

    def release_cpus(self):
        try:
            os.sched_setaffinity(0, self.all_cpus)
        except Exception:
            pass

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op="max"):
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def sum_u64(self, value):
        """Sum of the ranks' 64-bit hashes modulo 2^64 (two 32-bit halves through an int64 all-reduce)."""
        lo, hi = value & 0xffffffff, (value >> 32) & 0xffffffff
        t = self.torch.tensor([lo, hi], dtype=self.torch.int64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        lo, hi = int(t[0]), int(t[1])
        return (lo + (hi << 32)) & 0xffffffffffffffff

    def repeats(self, ms_for_k_steps):
        """Back-to-back repeats of the K steps that make the timed region last MIN_TIMED_S (same on every rank)."""
        ms = self.reduce([ms_for_k_steps])[0]
        return int(min(max(math.ceil(MIN_TIMED_S * 1e3 / max(ms, 1e-3)), 1), 4096))

    def done(self):
        if self.world > 1 and self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's own ORBextractor.cc (oracle/_ref) + the stereo matcher
# --------------------------------------------------------------------------------------------
def cpu_stereo_frames(pairs, cores):
    """Runs extraction of both eyes + stereo matching for every pair on `cores` host threads.
    Returns (seconds, kind)."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    use_ref = oracle.ref_available()
    tl = threading.local()

    def eye(img):
        if not hasattr(tl, "ex"):
            tl.ex = oracle.ReferenceExtractor(NFEAT) if use_ref else oracle.OracleExtractor(NFEAT)
        k, d = tl.ex(img)
        return k, d, [tl.ex.level(l) for l in range(8)]

    tables = oracle.OracleExtractor(NFEAT).tables()

    def stereo(lr):
        (kL, dL, pL), (kR, dR, pR) = lr
        return oracle.stereo_match(kL, dL, kR, dR, pL, pR, tables["scale"], tables["inv_scale"],
                                   synth.KITTI_BF, 0.0, synth.KITTI_FX)

    with ThreadPoolExecutor(cores) as pool:
        list(pool.map(eye, [pairs[0][0]] * cores))            # per-thread extractor construction, untimed
        t0 = time.perf_counter()
        eyes = list(pool.map(eye, [im for p in pairs for im in p]))
        list(pool.map(stereo, [(eyes[2 * i], eyes[2 * i + 1]) for i in range(len(pairs))]))
        dt = time.perf_counter() - t0
    return dt, ("reference" if use_ref else "port")


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nper = max(cores // 2, 1) * 2          # frames per step: one bounded sample, two eyes per core pair
    pairs = [synth.stereo_pair((H, W), i) for i in range(nper)]
    for _ in range(args.warmup):
        cpu_stereo_frames(pairs[:max(cores // 2, 1)], cores)
    total = 0.0
    kind = "port"
    for _ in range(args.steps):
        dt, kind = cpu_stereo_frames(pairs, cores)
        total += dt
    fps = nper * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": nper},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": f"{nper} stereo frames per step x {args.steps} steps; reference src/ORBextractor.cc compiled "
                                   f"in place (oracle/_ref) per eye + ComputeStereoMatches, {cores} host threads"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm: the headline (configs[1])
# --------------------------------------------------------------------------------------------
def load_pipe_counters():
    """ncu instruction / pipe counts of the kernels (profiles/traffic.json, written by tools/ncu_summary.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def measure_stereo(ctx, args):
    torch = ctx.torch
    import ctypes as C
    from object_slam_b200._capi import check, lib
    from object_slam_b200.extractor import ORBextractor, StereoFrames, stereo_match_device
    rank, local_rank, world = ctx.rank, ctx.local_rank, ctx.world
    F = args.frames
    # synthetic stereo frames, distinct per rank and per slot (seed = global frame index)
    pairs = [synth.stereo_pair((H, W), rank * F + i) for i in range(F)]
    hostL = np.zeros((F, H, PITCH), np.uint8)
    hostR = np.zeros((F, H, PITCH), np.uint8)
    for i, (l, r) in enumerate(pairs):
        hostL[i, :, :W] = l
        hostR[i, :, :W] = r
    dL = torch.from_numpy(hostL).cuda()
    dR = torch.from_numpy(hostR).cuda()

    mk = lambda: ORBextractor(NFEAT, 1.2, 8, 20, 7, max_size=(W, H), max_batch=F, device=local_rank)
    exL, exR = mk(), mk()
    tstream, tstream2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream, stream2 = tstream.cuda_stream, tstream2.cuda_stream

    # Device-resident pipelines: each owns an extractor pair and two streams; consecutive steps go to alternate
    # pipelines, so the tail of one batch overlaps the head of the next (all kernels here are latency / issue
    # bound, none fills the machine alone).
    NPIPES = int(os.environ.get("OBS_BENCH_PIPES", "2"))
    dpipes = [(exL, exR, tstream, tstream2)]
    for _ in range(NPIPES - 1):
        dpipes.append((mk(), mk(), torch.cuda.Stream(), torch.cuda.Stream()))
    step_no = [0]

    def step_device():
        # the two eyes on two streams, like the reference's two extraction threads (Frame.cc:78-81); the stereo match
        # (left stream) waits for the right eye inside the library, and the right eye's next extraction waits for it
        eL, eR, sA, sB = dpipes[step_no[0] % NPIPES]
        step_no[0] += 1
        eL.extract_device(dL.data_ptr(), F, W, H, PITCH, H * PITCH, sA.cuda_stream)
        eR.extract_device(dR.data_ptr(), F, W, H, PITCH, H * PITCH, sB.cuda_stream)
        stereo_match_device(eL, eR, synth.KITTI_BF, 0.0, synth.KITTI_FX, sA.cuda_stream)

    def timed(nsteps):
        step_no[0] = 0
        ctx.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _, _, sA, sB in dpipes:            # nothing of the timed region starts before ev0
            for st_ in (sA, sB):
                if st_ is not tstream:
                    st_.wait_stream(tstream)
        for _ in range(nsteps):
            step_device()
        for _, _, sA, sB in dpipes:            # ev1 follows the last kernel of every pipeline
            for st_ in (sA, sB):
                if st_ is not tstream:
                    tstream.wait_stream(st_)
        ev1.record()
        ctx.barrier()
        return ev0.elapsed_time(ev1)

    # ---- device-resident throughput
    for _ in range(max(args.warmup, NPIPES)):
        step_device()
    reps = ctx.repeats(timed(args.steps))          # a first pass of the K steps sizes the timed region (and warms up further)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total = ctx.reduce([timed(reps * args.steps)])[0]
    clocks = sampler.stop() if sampler else None
    timed_steps = reps * args.steps
    value = world * F * timed_steps / (ms_total * 1e-3)

    # per-kernel durations: a profiling pass of the same run directly after the timed region, one eye at a time with
    # the stages serialised (in the timed region four streams overlap, so per-kernel brackets there include waiting)
    exL.set_profiling(True)
    exR.set_profiling(True)
    for _ in range(min(args.steps, 10)):
        exL.extract_device(dL.data_ptr(), F, W, H, PITCH, H * PITCH, stream)
        torch.cuda.synchronize()
        exR.extract_device(dR.data_ptr(), F, W, H, PITCH, H * PITCH, stream2)
        torch.cuda.synchronize()
        stereo_match_device(exL, exR, synth.KITTI_BF, 0.0, synth.KITTI_FX, stream)
        torch.cuda.synchronize()
    stL, ncL, nsL = exL.stage_ms()
    stR, ncR, _ = exR.stage_ms()
    exL.set_profiling(False)
    exR.set_profiling(False)
    counts = exL.fetch_counts()
    countsR = exR.fetch_counts()

    # ---- the white-noise stress case (SURVEY 8d input 2: ~42k level-0 candidates per image, the minTh fallback never fires,
    # the quadtree leaves shared memory for its global-memory path): same step, same batch size
    stress = None
    if not args.no_stress:
        for i in range(F):
            l = synth.noise_image((H, W), 7_000_000 + rank * F + i)
            hostL[i, :, :W] = l
            hostR[i, :, :W - 24] = l[:, 24:]
            hostR[i, :, W - 24:W] = l[:, W - 1:W]
        dL.copy_(torch.from_numpy(hostL)); dR.copy_(torch.from_numpy(hostR))
        for _ in range(3):
            step_device()
        ms_s = ctx.reduce([timed(max(args.steps, 4))])[0]
        stress = {"value": world * F * max(args.steps, 4) / (ms_s * 1e-3), "unit": "frames/s",
                  "input": "white-noise stereo pairs (right = left shifted by 24 px)",
                  "mean_keypoints_left": float(exL.fetch_counts().mean()),
                  "fraction_of_headline": world * F * max(args.steps, 4) / (ms_s * 1e-3) / value}
        for i, (l, r) in enumerate(pairs):
            hostL[i, :, :W] = l
            hostR[i, :, :W] = r
        dL.copy_(torch.from_numpy(hostL)); dR.copy_(torch.from_numpy(hostR))
    del dpipes[1:]

    # ---- end to end through the host-buffer C ABI: obs_stereo_frames_submit / _wait (both eyes + ComputeStereoMatches in one
    # call).  Inputs and outputs live in page-locked host memory allocated on the CPUs next to this GPU; every step moves its
    # images host->device and its keypoints, descriptors, uRight and depth device->host.  ONE host thread keeps WORKERS
    # handle pairs in flight round-robin (wait for the oldest, resubmit it), so one step's transfers overlap the others'
    # kernels and no thread spins: the wait sleeps on a blocking-sync event.
    WORKERS = args.e2e_pipelines
    sf = [StereoFrames(NFEAT, 1.2, 8, 20, 7, (W, H), F, device=local_rank) for _ in range(WORKERS)]
    for p in sf:
        for i, (l, r) in enumerate(pairs):
            p.left[i] = l
            p.right[i] = r

    def run_e2e(nsteps):
        live = 0
        for i in range(nsteps + WORKERS):
            p = sf[i % WORKERS]
            if i >= WORKERS:
                p.wait()
            if i < nsteps:
                p.submit(synth.KITTI_BF, 0.0, synth.KITTI_FX)
        return live

    run_e2e(2 * WORKERS)
    ctx.barrier()
    t0 = time.perf_counter()
    run_e2e(max(args.steps, WORKERS))
    e2e_reps = ctx.repeats(1e3 * (time.perf_counter() - t0))
    e2e_steps = e2e_reps * max(args.steps, WORKERS)
    ctx.barrier()
    t0 = time.perf_counter()
    run_e2e(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = ctx.reduce([time.perf_counter() - t0])[0]
    e2e_value = world * F * e2e_steps / e2e_s
    # the e2e results equal the device-resident ones
    k0 = sf[0].results(1)[0]
    assert len(k0[0]) == int(counts[0]), "host-buffer path and device path disagree on the keypoint count of frame 0"
    h2d, d2h = sf[0].h2d_bytes, sf[0].d2h_bytes

    if rank != 0:
        return None

    # ---- roofline of the dominant kernel (per-launch CUDA-event time of the profiling pass above)
    peak, peak_src = measured_peaks()
    nkp = float(counts.mean())
    alg = {   # algorithmic bytes per launch (one eye, F images): each stage reads its input once, writes its output once
        "pyramid": F * (sum(LEVEL_PX[:-1]) + sum(LEVEL_PX[1:])),
        "fast": F * sum(LEVEL_PX),
        "quadtree": None,
        "blur": F * 2 * sum(LEVEL_PX),
        "describe": F * nkp * (749 + 512 + 60),
        "stereo": F * (2 * nkp * 32 + nkp * (121 + 11 * 21) + nkp * 8),
    }
    per_launch_ms = {k: (stL[k] + stR[k]) / max(ncL + ncR, 1) for k in ORBextractor.STAGES}
    per_launch_ms["stereo"] = stL["stereo"] / max(nsL, 1)
    step_ms = ms_total / timed_steps
    serial_ms = sum((2 * v if k != "stereo" else v) for k, v in per_launch_ms.items())
    shares = {k: (2 * v if k != "stereo" else v) / serial_ms for k, v in per_launch_ms.items()}
    dominant = max(shares, key=shares.get)
    tj = load_pipe_counters()
    per_img = float(tj.get("_images_per_launch", F))
    traffic = tj.get(dominant)
    if traffic is not None:
        traffic = traffic * F / per_img
    # Per stage: achieved HBM fraction and achieved fraction of the pipes the stage is bound by.  Numerators: warp-instruction /
    # ALU-instruction / shared-memory-wavefront counts of the committed ncu capture (profiles/traffic.json `_pipes`, written by
    # tools/ncu_pipes.py, per 64 images) over the LIVE launch time.  Denominators: measured here on this device by
    # obs_microbench_pipes (issue rate with both integer pipes fed, ALU-pipe rate, conflict-free LDS wavefront rate).
    gi, ga, gl = C.c_double(), C.c_double(), C.c_double()
    check(lib().obs_microbench_pipes(local_rank, C.byref(gi), C.byref(ga), C.byref(gl)))
    pipe_peaks = {"issue_ginst_per_s": gi.value, "alu_ginst_per_s": ga.value, "lds_gwavefronts_per_s": gl.value,
                  "source": "obs_microbench_pipes on this device (csrc/microbench.cu), best of 3"}
    stage_fracs = {}
    for k, ms in per_launch_ms.items():
        ent = {"ms_per_launch": ms}
        if alg.get(k):
            ent["hbm_gbs"] = alg[k] / (ms * 1e-3) / 1e9
            ent["hbm_frac"] = ent["hbm_gbs"] / peak
        pp = tj.get("_pipes", {}).get(k)
        if pp:
            sc = F / per_img
            t = ms * 1e-3
            ent["issue_frac"] = pp["warp_instructions"] * sc / t / (gi.value * 1e9)
            if pp.get("alu_instructions"):
                ent["alu_pipe_frac"] = pp["alu_instructions"] * sc / t / (ga.value * 1e9)
            if pp.get("lsu_wavefronts"):
                ent["lds_wavefront_frac"] = pp["lsu_wavefronts"] * sc / t / (gl.value * 1e9)
            fr = {"issue": ent["issue_frac"], "alu": ent.get("alu_pipe_frac", 0.0), "lds_wavefronts": ent.get("lds_wavefront_frac", 0.0),
                  "hbm": ent.get("hbm_frac", 0.0)}
            ent["binding_pipe"] = max(fr, key=fr.get)
            ent["binding_frac"] = fr[ent["binding_pipe"]]
        stage_fracs[k] = ent
    achieved = alg[dominant] / (per_launch_ms[dominant] * 1e-3) / 1e9 if alg[dominant] is not None else 0.0
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "stages": stage_fracs, "pipe_peaks": pipe_peaks,
                "bound_note": "every stage is bound inside the SM (instruction issue / ALU pipe / shared-memory wavefronts) before it is HBM "
                              "bound (DESIGN.md section 4): `stages` holds, per stage, the HBM fraction and the fractions of the measured "
                              "pipe peaks (`pipe_peaks`); counts from the committed ncu capture (profiles/traffic.json) over the live launch time",
                "per_launch_ms": per_launch_ms, "share_of_step": shares,
                "serialised_step_ms": serial_ms,
                "note": "per_launch_ms: CUDA-event brackets of a profiling pass run directly after the timed region with every kernel "
                        "serialised (shares are of that serial sum); in the timed region the eyes, the blur and consecutive batches "
                        "overlap on four streams, which is why ms_per_step is below serialised_step_ms",
                "whole_step": {"algorithmic_bytes": F * 19.5e6, "achieved": F * 19.5e6 / (step_ms * 1e-3) / 1e9,
                               "frac": F * 19.5e6 / (step_ms * 1e-3) / 1e9 / peak}}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        ctx.release_cpus()
        cores = os.cpu_count() or 1
        sample = [pairs[i % F] for i in range(args.cpu_frames)]
        dt, kind = cpu_stereo_frames(sample, cores)
        cpu_baseline = {"value": len(sample) / dt, "unit": "frames/s", "cores": cores, "kind": kind,
                        "sample": f"{len(sample)} stereo frames (the step's {F}, cycled); reference src/ORBextractor.cc compiled in place "
                                  f"(oracle/_ref) per eye + ComputeStereoMatches on {cores} host threads, {dt:.1f} s"}

    return {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "frames_per_step_per_gpu": F, "parallelism": f"frames sharded over {world} GPU(s), no collective",
                   "timed_steps": timed_steps, "repeats": reps, "timed_region_s": ms_total * 1e-3,
                   "l2": f"working set per step {2 * F * 7.3:.0f} MB (> 126 MB L2): inputs and intermediates larger than L2",
                   "mean_keypoints_left": nkp, "mean_keypoints_right": float(countsR.mean()),
                   "host_affinity": ctx.affinity_note},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "seconds": e2e_s,
                "api": "obs_stereo_frames_submit / obs_stereo_frames_wait (both eyes + ComputeStereoMatches per call), page-locked host "
                       f"buffers in and out; one host thread keeps {WORKERS} handle pairs in flight and sleeps on a blocking-sync event"},
        # per eye: 3 tiled resize levels + the cluster-chained tail, FAST, quadtree, blur, describe; per frame batch: stereo rows / match / filter
        "gpu_launches": (2 * 8 + 3) * timed_steps,
        "roofline": roofline,
        "stress": stress,
        "cpu_baseline": cpu_baseline,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=128, help="stereo frames per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-frames", type=int, default=256, help="stereo frames of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stress", action="store_true", help="skip the white-noise stress case")
    ap.add_argument("--no-sub", action="store_true", help="headline only: skip the configs[2] / configs[3] / configs[4] sub-results")
    ap.add_argument("--e2e-pipelines", type=int, default=3, help="handle pairs the e2e leg keeps in flight from its one host thread")
    ap.add_argument("--workload", default="stereo", choices=["stereo", "knn2", "projection", "bow", "extract"],
                    help="stereo = configs[1] (the headline, with configs[3] and configs[4] as sub-results); knn2 = configs[4] alone; "
                         "extract = configs[3] alone; projection = configs[2] TUM-shape extraction + SearchByProjection against a "
                         "20k-point map; bow = SearchByBoW keyframe pairs")
    ap.add_argument("--keyframes", type=int, default=4096, help="knn2: keyframes in total (sharded over the GPUs)")
    ap.add_argument("--window", type=int, default=8, help="knn2: every keyframe is matched against the +-window neighbours")
    ap.add_argument("--gather-chunks", type=int, default=4, help="knn2: chunks of the all-gather (matching overlaps the later chunks)")
    ap.add_argument("--knn2-engine", default="auto", choices=["auto", "tensor", "popc"])
    ap.add_argument("--knn2-cta-pair", type=int, default=1, choices=[0, 1], help="knn2 tensor engine: CTA pairs (cta_group::2) or one CTA per SM")
    ap.add_argument("--map-points", type=int, default=20000, help="projection: map points per frame")
    ap.add_argument("--total-frames", type=int, default=32768, help="extract: frames of the whole offline job (configs[3]), sharded over the GPUs")
    ap.add_argument("--pool", type=int, default=512, help="extract: distinct synthetic frames (seed = global frame index mod pool)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    ctx = Ctx(args)
    import bench_match
    if args.impl == "reference":
        if args.workload == "stereo":
            run_reference_arm(args, ctx.rank)
        else:
            bench_match.run_reference(args, ctx)
        return
    ctx.init_gpu()
    if args.workload != "stereo":
        line = bench_match.run(args, ctx, ClockSampler)
        if ctx.rank == 0 and line is not None:
            print(json.dumps(line), flush=True)
        ctx.done()
        return

    sub = {}
    if not args.no_sub:
        # the sharded batch workloads north_star names, measured in the same process group
        sub["configs[2]"] = bench_match.run_projection(args, ctx.rank, ctx.local_rank, ctx.world, ClockSampler, as_sub=True)
        sub["configs[3]"] = bench_match.measure_extract(args, ctx, ClockSampler)
        sub["configs[4]"] = bench_match.measure_knn2(args, ctx, ClockSampler)
    line = measure_stereo(ctx, args)
    if ctx.rank == 0:
        if sub:
            line["sub"] = sub
        print(json.dumps(line), flush=True)
    ctx.done()


if __name__ == "__main__":
    main()
