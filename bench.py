#!/usr/bin/env python
"""Benchmark of the north-star path: ORB extraction of both eyes + Frame::ComputeStereoMatches on
KITTI-shape (1241x376, nFeatures=2000) synthetic stereo frames.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl ours|reference]

One "step" = one batch of F stereo frames through the whole hot path.  `value` is stereo frames/s
with the images already resident in HBM (CUDA-event timed on the launching stream); `e2e` is the same
through the host-buffer C ABI (host images in, keypoints/descriptors/uRight/depth out), wall clock
around synchronous calls.  Multi-GPU: one process per GPU (torchrun), frames sharded by rank, no
data-path collective; timing is the max over ranks.

Only the `cpu_baseline` leg and `--impl reference` touch oracle/ (the reference's own
ORBextractor.cc compiled in place + the restated stereo matcher) -- as the thing timed beside the
product, never inside it.
"""
import argparse
import json
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
LEVEL_PX = [1241 * 376, 1034 * 313, 862 * 261, 718 * 218, 598 * 181, 499 * 151, 416 * 126, 346 * 105]


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

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
        sm, mx, reasons = [], [], set()
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
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's own ORBextractor.cc (oracle/_ref) + the restated stereo matcher
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
        "config": {"workload": "configs[1]: KITTI-shape 1241x376 synthetic stereo pairs, nFeatures=2000, "
                               "ORB extraction of both eyes + ComputeStereoMatches", "frames_per_step": nper},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": f"{nper} stereo frames per step x {args.steps} steps; reference src/ORBextractor.cc compiled "
                                   f"in place (oracle/_ref) per eye + restated ComputeStereoMatches, {cores} host threads"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=128, help="stereo frames per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-frames", type=int, default=256, help="stereo frames of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-pipelines", type=int, default=3, help="host-buffer pipelines of the e2e leg (each: an extractor pair + pinned buffers)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the e2e leg (0 = min(steps, 12))")
    ap.add_argument("--workload", default="stereo", choices=["stereo", "knn2", "projection", "bow", "extract"],
                    help="stereo = configs[1] (the headline); knn2 = configs[4] keyframe-vs-keyframe Hamming matching; "
                         "projection = configs[2] TUM-shape extraction + SearchByProjection against a 20k-point map; "
                         "extract = configs[3] batched offline extraction of --total-frames KITTI-shape frames")
    ap.add_argument("--keyframes", type=int, default=4096, help="knn2: keyframes in total (sharded over the GPUs)")
    ap.add_argument("--window", type=int, default=8, help="knn2: every keyframe is matched against the +-window neighbours")
    ap.add_argument("--map-points", type=int, default=20000, help="projection: map points per frame")
    ap.add_argument("--total-frames", type=int, default=32768, help="extract: frames of the whole offline job (configs[3]), sharded over the GPUs")
    ap.add_argument("--pool-batches", type=int, default=4, help="extract: batches of distinct synthetic frames resident per GPU (cycled)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.workload != "stereo":
        import bench_match
        return bench_match.run(args, rank, local_rank, world, ClockSampler)
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from object_slam_b200.extractor import ORBextractor, ComputeStereoMatches, stereo_match_device

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: object_slam_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

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

    exL = ORBextractor(NFEAT, 1.2, 8, 20, 7, max_size=(W, H), max_batch=F, device=local_rank)
    exR = ORBextractor(NFEAT, 1.2, 8, 20, 7, max_size=(W, H), max_batch=F, device=local_rank)
    # a side stream of torch's: the ABI treats a NULL stream as "the handle's own stream"
    tstream = torch.cuda.Stream()
    tstream2 = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream, stream2 = tstream.cuda_stream, tstream2.cuda_stream

    # Device-resident pipelines: each owns an extractor pair and two streams; consecutive steps go to alternate
    # pipelines, so the tail of one batch overlaps the head of the next (all kernels here are latency / issue
    # bound, none fills the machine alone).
    NPIPES = int(os.environ.get("OBS_BENCH_PIPES", "2"))
    dpipes = [(exL, exR, tstream, tstream2)]
    for _ in range(NPIPES - 1):
        dpipes.append((ORBextractor(NFEAT, 1.2, 8, 20, 7, max_size=(W, H), max_batch=F, device=local_rank),
                       ORBextractor(NFEAT, 1.2, 8, 20, 7, max_size=(W, H), max_batch=F, device=local_rank),
                       torch.cuda.Stream(), torch.cuda.Stream()))
    step_no = [0]

    def step_device():
        # the two eyes on two streams, like the reference's two extraction threads (Frame.cc:78-81); the stereo match
        # (left stream) waits for the right eye inside the library, and the right eye's next extraction waits for it
        eL, eR, sA, sB = dpipes[step_no[0] % NPIPES]
        step_no[0] += 1
        eL.extract_device(dL.data_ptr(), F, W, H, PITCH, H * PITCH, sA.cuda_stream)
        eR.extract_device(dR.data_ptr(), F, W, H, PITCH, H * PITCH, sB.cuda_stream)
        stereo_match_device(eL, eR, synth.KITTI_BF, 0.0, synth.KITTI_FX, sA.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    for _ in range(max(args.warmup, NPIPES)):
        step_device()
    step_no[0] = 0
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _, _, sA, sB in dpipes:            # nothing of the timed region starts before ev0
        for st_ in (sA, sB):
            if st_ is not tstream:
                st_.wait_stream(tstream)
    for _ in range(args.steps):
        step_device()
    for _, _, sA, sB in dpipes:            # ev1 follows the last kernel of every pipeline
        for st_ in (sA, sB):
            if st_ is not tstream:
                tstream.wait_stream(st_)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
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

    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * F * args.steps / (ms_total * 1e-3)

    # ---- end to end through the host-buffer C ABI (two threads for the two eyes, like Frame.cc:78-81).
    # Inputs and outputs live in page-locked host memory (obs_host_alloc); every step moves its images
    # host->device and its keypoints, descriptors, uRight and depth device->host.  Two independent pipelines
    # (each with its own extractor pair and buffers) take alternate steps from two worker threads, so one step's
    # transfers overlap the other's kernels (double buffering across steps); throughput = frames / wall clock.
    from object_slam_b200._capi import pinned_empty, KEYPOINT_DTYPE
    cap = exL.capacity
    WORKERS = args.e2e_pipelines

    class Pipe:
        def __init__(self, eL, eR):
            self.eL, self.eR = eL, eR
            self.pinL = pinned_empty((F, H, W), np.uint8)
            self.pinR = pinned_empty((F, H, W), np.uint8)
            for i, (l, r) in enumerate(pairs):
                self.pinL[i] = l
                self.pinR[i] = r
            mk = lambda: (pinned_empty((F, cap), KEYPOINT_DTYPE), pinned_empty((F, cap, 32), np.uint8), pinned_empty((F,), np.int32))
            self.outL, self.outR = mk(), mk()
            self.outS = (pinned_empty((F, cap), np.float32), pinned_empty((F, cap), np.float32))

        def step(self):
            res = [None, None]
            th = threading.Thread(target=lambda: res.__setitem__(1, self.eR.extract_batch(self.pinR, out=self.outR, copy=False)))
            th.start()
            res[0] = self.eL.extract_batch(self.pinL, out=self.outL, copy=False)
            th.join()
            return res, ComputeStereoMatches(self.eL, self.eR, synth.KITTI_BF, 0.0, synth.KITTI_FX, out=self.outS)

    pipes = [Pipe(exL, exR)]
    for _ in range(WORKERS - 1):
        pipes.append(Pipe(ORBextractor(NFEAT, 1.2, 8, 20, 7, max_size=(W, H), max_batch=F, device=local_rank),
                          ORBextractor(NFEAT, 1.2, 8, 20, 7, max_size=(W, H), max_batch=F, device=local_rank)))
    e2e_steps = max(WORKERS, (args.e2e_steps or max(4, min(args.steps, 12))) // WORKERS * WORKERS)

    def run_pipes(nsteps):
        ths = [threading.Thread(target=lambda p=p: [p.step() for _ in range(nsteps // WORKERS)]) for p in pipes[1:]]
        for th in ths:
            th.start()
        for _ in range(nsteps // WORKERS):
            pipes[0].step()
        for th in ths:
            th.join()

    run_pipes(2 * WORKERS)
    barrier()
    t0 = time.perf_counter()
    run_pipes(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * F * e2e_steps / float(t.item())
    _, rec_bytes, kp_cap = exL.results_device()
    h2d = 2 * F * H * W
    d2h = 2 * F * (kp_cap * 60 + 4) + 2 * F * kp_cap * 4 + F * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (per-launch CUDA-event time from the timed region above)
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
    step_ms = ms_total / args.steps
    serial_ms = sum((2 * v if k != "stereo" else v) for k, v in per_launch_ms.items())
    shares = {k: (2 * v if k != "stereo" else v) / serial_ms for k, v in per_launch_ms.items()}
    dominant = max(shares, key=shares.get)
    traffic, pipes = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        # the ncu capture ran with tj["_images_per_launch"] images per launch; DRAM traffic scales with the batch
        traffic = tj.get(dominant)
        if traffic is not None:
            traffic = traffic * F / float(tj.get("_images_per_launch", F))
        pipes = tj.get("_pipes", {}).get(dominant)
    except Exception:
        pass
    issue = None
    if pipes and pipes.get("warp_instructions"):
        # warp instructions of one launch (ncu count at 64 images, scaled to this batch) over the live launch time, against
        # 148 SMs x 4 schedulers x SM clock
        wi = pipes["warp_instructions"] * F / 64.0
        sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or 1965.0)
        pk = 148 * 4 * sm_hz
        issue = {"achieved": wi / (per_launch_ms[dominant] * 1e-3) / 1e12, "peak": pk / 1e12, "unit": "T warp-instructions/s",
                 "frac": wi / (per_launch_ms[dominant] * 1e-3) / pk,
                 "source": "instruction count from profiles/traffic.json (_pipes, ncu smsp__inst_executed.sum), time measured live"}
    if alg[dominant] is not None:
        achieved = alg[dominant] / (per_launch_ms[dominant] * 1e-3) / 1e9
    else:
        achieved = 0.0
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "ncu_pipes": pipes,
                "issue": issue,
                "bound_note": "every stage is integer-issue bound before it is HBM bound (DESIGN.md section 4): ncu_pipes holds the ncu "
                              "pipe utilisation of the dominant kernel (ALU pipe: one warp instruction per 2 cycles per SM sub-partition)",
                "per_launch_ms": per_launch_ms, "share_of_step": shares,
                "serialised_step_ms": serial_ms,
                "note": "per_launch_ms: CUDA-event brackets of a profiling pass run directly after the timed region with every kernel "
                        "serialised (shares are of that serial sum); in the timed region the eyes, the blur and consecutive batches "
                        "overlap on four streams, which is why ms_per_step is below serialised_step_ms",
                "whole_step": {"algorithmic_bytes": F * 19.5e6, "achieved": F * 19.5e6 / (step_ms * 1e-3) / 1e9,
                               "frac": F * 19.5e6 / (step_ms * 1e-3) / 1e9 / peak}}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        sample = [pairs[i % F] for i in range(args.cpu_frames)]
        dt, kind = cpu_stereo_frames(sample, cores)
        cpu_baseline = {"value": len(sample) / dt, "unit": "frames/s", "cores": cores, "kind": kind,
                        "sample": f"{len(sample)} stereo frames (the step's {F}, cycled); reference src/ORBextractor.cc compiled in place "
                                  f"(oracle/_ref) per eye + restated ComputeStereoMatches on {cores} host threads, {dt:.1f} s"}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "configs[1]: KITTI-shape 1241x376 synthetic stereo pairs, nFeatures=2000, nLevels=8, scale 1.2, "
                               "FAST 20/7: ORB extraction of both eyes + Frame::ComputeStereoMatches",
                   "frames_per_step_per_gpu": F, "parallelism": f"frames sharded over {world} GPU(s), no collective",
                   "l2": f"working set per step {2 * F * 7.3:.0f} MB (> 126 MB L2): inputs and intermediates larger than L2",
                   "mean_keypoints_left": nkp, "mean_keypoints_right": float(countsR.mean())},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "obs_extract_batch x2 (two host threads) + obs_stereo_match, page-locked host buffers in and out; "
                                            f"{WORKERS} pipelines take alternate steps (transfers of one overlap kernels of the other)"},
        # per eye: 3 tiled resize levels + the cluster-chained tail, FAST, quadtree, blur, describe; per frame batch: stereo rows / match / filter
        "gpu_launches": (2 * 8 + 3) * args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
