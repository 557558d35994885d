"""The other workloads of bench.py (`--workload knn2|projection|bow|extract`); same JSON contract as the headline line.

knn2        configs[4]: K keyframes x 2000 descriptors, every keyframe matched against its +-window neighbours
            (best / second best / ratio test).  Query keyframes are sharded over the ranks (strong scaling: the
            total is fixed); every rank needs all descriptor sets, so each step starts with the all-gather of the
            descriptor shards over NCCL (csrc/comm.cu) and matches the pairs whose database keyframe is local
            while it is in flight.
projection  configs[2]: TUM-shape frames: extraction + frame-set (grid) build + SearchByProjection against a
            20k-point map per frame.  Frames are sharded over the ranks, no collective (weak scaling).
extract     configs[3]: batched offline extraction of 32768 KITTI-shape frames, contiguous frame shards per rank, no
            collective; a step is one batch of F frames per rank.
"""
import ctypes as C
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from object_slam_b200 import sharding, synth  # noqa: E402

NDESC = 2000
TH_LOW, RATIO = 50, 0.6


def _device_keyframes(torch, K, seed, dev):
    """cfg5 descriptors generated on the device: uniform random, 30 % of keyframe i+1 are noisy copies of keyframe i."""
    g = torch.Generator(device=dev); g.manual_seed(seed)
    d = torch.randint(0, 256, (K, NDESC, 32), dtype=torch.uint8, device=dev, generator=g)
    m = int(NDESC * 0.3)
    noise = torch.randint(0, 256, (K, m, 32), dtype=torch.uint8, device=dev, generator=g)
    noise &= torch.randint(0, 256, (K, m, 32), dtype=torch.uint8, device=dev, generator=g)
    noise &= torch.randint(0, 256, (K, m, 32), dtype=torch.uint8, device=dev, generator=g)      # ~1/8 of the bits set
    for k in range(1, K):
        d[k, NDESC - m:] = d[k - 1, :m] ^ noise[k]
    return d


def _popc_peak(mode=2):
    from object_slam_b200._capi import check, lib
    g = C.c_double()
    check(lib().obs_microbench_popc(int(os.environ.get("LOCAL_RANK", "0")), mode, C.byref(g)))
    return g.value


def cpu_knn2(D, pairs, cores):
    import oracle
    def one(p):
        return oracle.hamming_knn2(D[p[0]], D[p[1]], TH_LOW, RATIO)[0]
    with ThreadPoolExecutor(cores) as pool:
        list(pool.map(one, pairs[:cores]))
        t0 = time.perf_counter()
        list(pool.map(one, pairs))
        return time.perf_counter() - t0


def run_knn2(args, rank, local_rank, world, ClockSampler):
    K, Wn = args.keyframes, args.window
    metric = "Hamming distance evaluations/s, brute-force keyframe-vs-keyframe matching (best + second best + ratio), 2000 descriptors per keyframe"
    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        D = synth.keyframe_descriptors(8, NDESC, 0)
        pairs = sharding.window_pairs(0, 8, 8, 8)[:max(cores * 2, 8)]
        total = 0.0
        for _ in range(args.warmup):
            cpu_knn2(D, pairs[:cores], cores)
        for _ in range(args.steps):
            total += cpu_knn2(D, pairs, cores)
        v = len(pairs) * NDESC * NDESC * args.steps / total
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": "distances/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "u32 popcount", "data": "synthetic",
                          "config": {"workload": "configs[4] brute-force keyframe matching", "pairs_per_step": len(pairs)},
                          "cpu_baseline": {"value": v, "unit": "distances/s", "cores": cores, "kind": "port",
                                           "sample": f"{len(pairs)} keyframe pairs per step, restated SearchByBoW candidate loop with the reference's SWAR DescriptorDistance"},
                          "e2e": {"value": v, "unit": "distances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    import torch
    import torch.distributed as dist
    from object_slam_b200._capi import check, lib, pinned_empty
    from object_slam_b200.matcher import ORBmatcher
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    per = sharding.padded_shard(K, world)
    lo, hi = rank * per, min(K, (rank + 1) * per)
    full = _device_keyframes(torch, K, 1234, dev)              # every rank generates the same set, keeps only its shard
    local = torch.zeros((per, NDESC, 32), dtype=torch.uint8, device=dev)
    local[:hi - lo] = full[lo:hi]
    del full
    allD = torch.zeros((world * per, NDESC, 32), dtype=torch.uint8, device=dev)
    pairs = sharding.window_pairs(lo, hi, K, Wn)
    p_loc, p_rem = sharding.split_by_locality(pairs, lo, hi)
    P = len(pairs)
    dp_loc = torch.from_numpy(p_loc).to(dev); dp_rem = torch.from_numpy(p_rem).to(dev)
    best = torch.empty((max(P, 1), NDESC), dtype=torch.int32, device=dev)
    M = ORBmatcher(RATIO, True, device=local_rank)
    mstream = M.stream
    comm = None
    if world > 1:
        def bcast(raw):
            t = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
            dist.broadcast(t, 0)
            return bytes(t.cpu().tolist())
        comm = sharding.Comm(rank, world, local_rank, bcast)
    local_bytes = per * NDESC * 32

    def knn(dpairs, n, out_off):
        if n:
            check(lib().obs_hamming_knn2(M._h, C.c_void_p(allD.data_ptr()), world * per, NDESC, C.c_void_p(dpairs.data_ptr()), n,
                                         TH_LOW, RATIO, C.c_void_p(best.data_ptr() + out_off * NDESC * 4), None, None))

    mine = allD[rank * per:(rank + 1) * per]                  # this rank's shard lives in its slot of the gathered set
    mine.copy_(local)
    torch.cuda.synchronize()

    def step():
        # the gather is ordered after everything queued on the matcher stream (the producer of `mine` and the
        # previous step's readers of the remote slots) and runs on the communicator's stream; pairs whose
        # database keyframe is local are matched meanwhile, the others wait for the gather
        if comm:
            comm.allgather(mine.data_ptr(), local_bytes, allD.data_ptr(), 1, mstream)
        knn(dp_loc, len(p_loc), 0)
        if comm:
            comm.wait(0, mstream)
        knn(dp_rem, len(p_rem), len(p_loc))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ms_stream = torch.cuda.ExternalStream(mstream)
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ms_stream):
        e0.record()
    for _ in range(args.steps):
        step()
    with torch.cuda.stream(ms_stream):
        e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total, float(P)], dtype=torch.float64, device=dev)
    if world > 1:
        tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        ms_total, P_all = float(tm[0]), int(ts[1])
    else:
        P_all = P
    dist_per_step = P_all * NDESC * NDESC
    value = dist_per_step * args.steps / (ms_total * 1e-3)
    matches = int((best[:P] >= 0).sum())

    # ---- end to end: descriptor shard from page-locked host memory in, best indices out, every step
    h_local = pinned_empty((per, NDESC, 32), np.uint8)
    h_local[:] = local.cpu().numpy()
    h_best = pinned_empty((max(P, 1), NDESC), np.int32)
    e2e_steps = max(3, min(args.steps, 5))

    def step_host():
        with torch.cuda.stream(ms_stream):
            mine.copy_(torch.from_numpy(h_local), non_blocking=True)
        step()
        with torch.cuda.stream(ms_stream):
            torch.from_numpy(h_best).copy_(best, non_blocking=True)
        M.sync()

    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = dist_per_step * e2e_steps / float(t.item())
    if rank == 0:
        popc_only = _popc_peak(2)          # Gdist/s at 8 popc per distance == popc pipe ceiling / 8
        kernel_ms = ms_total / args.steps
        peak_popc = popc_only * 8e9        # popc/s of one GPU
        ach_popc = 5.0 * (dist_per_step / world) / (kernel_ms * 1e-3)      # the kernel issues 5 popc per distance (carry-save folding)
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            Dh = synth.keyframe_descriptors(8, NDESC, 0)
            sp = sharding.window_pairs(0, 8, 8, 8)[:max(2 * cores, 8)]
            dt = cpu_knn2(Dh, sp, cores)
            cpu = {"value": len(sp) * NDESC * NDESC / dt, "unit": "distances/s", "cores": cores, "kind": "port",
                   "sample": f"{len(sp)} keyframe pairs (2000 x 2000 distances each) on {cores} host threads, {dt:.1f} s; restated "
                             "SearchByBoW candidate loop with the reference's SWAR DescriptorDistance"}
        print(json.dumps({
            "metric": metric, "value": value, "unit": "distances/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 popcount", "data": "synthetic (generated on the device)",
            "config": {"workload": f"configs[4]: {K} keyframes x {NDESC} descriptors, window +-{Wn} ({P_all} ordered keyframe pairs per step), "
                                   "ratio 0.6, TH_LOW 50", "parallelism": f"query keyframes sharded over {world} GPU(s); NCCL all-gather of the "
                                   "descriptor shards every step" if world > 1 else "1 GPU, no collective",
                       "l2": f"descriptor set {K * NDESC * 32 / 1e6:.0f} MB + {P_all * NDESC * 4 / 1e6:.0f} MB of results per step (> 126 MB L2)",
                       "queries_per_s": value / NDESC, "accepted_matches_rank0": matches},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "distances/s", "h2d_bytes_per_step": world * local_bytes, "d2h_bytes_per_step": P_all * NDESC * 4,
                    "steps": e2e_steps, "api": "descriptor shard H2D from page-locked memory, obs_comm_allgather, obs_hamming_knn2, best indices D2H"},
            "gpu_launches": (2 if world > 1 else 1) * args.steps,
            "roofline": {"bound": "int-popc", "kernel": "k_knn2", "achieved": ach_popc / 1e12, "peak": peak_popc / 1e12, "unit": "Tpopc/s",
                         "frac": ach_popc / peak_popc, "traffic": None,
                         "peak_source": "obs_microbench_popc mode 2 measured in this run (POPC pipe: 16 lanes/clk/SM)",
                         "distances_per_s_vs_plain_ceiling": (value / world) / (popc_only * 1e9)},
            "cpu_baseline": cpu}), flush=True)
    if comm:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ projection
def cpu_projection(imgs, mps, cores, nfeat, th):
    import oracle
    import threading
    import matcher_cases as mc
    tl = threading.local()
    use_ref = oracle.ref_available()
    sf = synth.scale_factors()

    def one(i):
        if not hasattr(tl, "ex"):
            tl.ex = oracle.ReferenceExtractor(nfeat) if use_ref else oracle.OracleExtractor(nfeat)
        k, d = tl.ex(imgs[i])
        F = oracle.OracleFrame(k, d, None, mc.bounds(synth.TUM_SHAPE))
        mp = mps[i % len(mps)]
        return oracle.search_by_projection_map(F, sf, *[mp[key] for key in mc.MP_KEYS], th, 0.8)[0]

    with ThreadPoolExecutor(cores) as pool:
        list(pool.map(one, range(min(cores, len(imgs)))))
        t0 = time.perf_counter()
        list(pool.map(one, range(len(imgs))))
        return time.perf_counter() - t0, ("reference" if use_ref else "port")


def run_projection(args, rank, local_rank, world, ClockSampler):
    import matcher_cases as mc
    Himg, Wimg = synth.TUM_SHAPE
    NF, NM, TH = 1000, args.map_points, 3.0
    F = args.frames
    metric = "frames/s, ORB extraction + SearchByProjection against a 20k-point local map, TUM 640x480, 1000 kp"
    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        imgs = [synth.blocky_image(synth.TUM_SHAPE, i) for i in range(2 * cores)]
        import oracle
        ex = oracle.OracleExtractor(NF)
        mps = []
        for i in range(4):
            k, d = ex(imgs[i])
            mps.append(synth.map_points_for_frame(k, d, synth.TUM_SHAPE, NM, 100 + i))
        for _ in range(args.warmup):
            cpu_projection(imgs[:cores], mps, cores, NF, TH)
        total, kind = 0.0, "port"
        for _ in range(args.steps):
            dt, kind = cpu_projection(imgs, mps, cores, NF, TH)
            total += dt
        v = len(imgs) * args.steps / total
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": {"workload": "configs[2] TUM-shape extraction + SearchByProjection", "frames_per_step": len(imgs)},
                          "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": kind,
                                           "sample": f"{len(imgs)} frames per step; reference ORBextractor.cc compiled in place + restated SearchByProjection"},
                          "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    import torch
    import torch.distributed as dist
    from object_slam_b200._capi import pinned_empty, KEYPOINT_DTYPE
    from object_slam_b200.extractor import ORBextractor
    from object_slam_b200.matcher import ORBmatcher
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    imgs = [synth.blocky_image(synth.TUM_SHAPE, rank * F + i) for i in range(F)]
    host = np.stack(imgs)
    dimg = torch.from_numpy(host).to(dev)
    ex = ORBextractor(NF, 1.2, 8, 20, 7, max_size=(Wimg, Himg), max_batch=F, device=local_rank)
    M = ORBmatcher(0.8, True, device=local_rank)
    fs = M.frame_set(ex.GetScaleFactors(), mc.bounds(synth.TUM_SHAPE), synth.camera_for(synth.TUM_SHAPE), max_frames=F, max_keypoints=ex.capacity)
    # map points per frame, derived from the frame's own keypoints (seeded); resident in HBM
    res = ex.extract_batch(host)
    mps = [synth.map_points_for_frame(k, d, synth.TUM_SHAPE, NM, 100 + rank * F + i) for i, (k, d) in enumerate(res[:8])]
    arr = [torch.from_numpy(np.stack([mps[i % len(mps)][k] for i in range(F)])).to(dev) for k in mc.MP_KEYS]
    kpm = torch.empty((F, fs.cap), dtype=torch.int32, device=dev)
    nm = torch.empty(F, dtype=torch.int32, device=dev)
    mstream = M.stream

    def step_device():
        ex.extract_device(dimg.data_ptr(), F, Wimg, Himg, Wimg, Himg * Wimg, mstream)
        fs.from_extractor(ex)
        M.SearchByProjection(fs, *[a.data_ptr() for a in arr], th=TH, n_points=NM, per_frame=True,
                             kp_match=kpm.data_ptr(), n_matches=nm.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ms_stream = torch.cuda.ExternalStream(mstream)
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    em0, em1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ms_stream):
        e0.record()
    for _ in range(args.steps):
        step_device()
    with torch.cuda.stream(ms_stream):
        e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    # the search alone (frame set already built)
    with torch.cuda.stream(ms_stream):
        em0.record()
    for _ in range(args.steps):
        M.SearchByProjection(fs, *[a.data_ptr() for a in arr], th=TH, n_points=NM, per_frame=True, kp_match=kpm.data_ptr(), n_matches=nm.data_ptr())
    with torch.cuda.stream(ms_stream):
        em1.record()
    barrier()
    ms_search = em0.elapsed_time(em1) / args.steps
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * F * args.steps / (ms_total * 1e-3)
    mean_matches = float(nm.float().mean())

    # ---- end to end: host images in; keypoints, descriptors and the keypoint -> map point assignment out
    cap = ex.capacity
    pin = pinned_empty((F, Himg, Wimg), np.uint8); pin[:] = host
    out = (pinned_empty((F, cap), KEYPOINT_DTYPE), pinned_empty((F, cap, 32), np.uint8), pinned_empty((F,), np.int32))
    hk = pinned_empty((F, fs.cap), np.int32); hn = pinned_empty((F,), np.int32)

    def step_host():
        ex.extract_batch(pin, out=out, copy=False)
        fs.from_extractor(ex)
        M.SearchByProjection(fs, *[a.data_ptr() for a in arr], th=TH, n_points=NM, per_frame=True, kp_match=hk, n_matches=hn)

    e2e_steps = max(3, min(args.steps, 10))
    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * F * e2e_steps / float(t.item())
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            sample = [imgs[i % F] for i in range(4 * cores)]
            dt, kind = cpu_projection(sample, mps, cores, NF, TH)
            cpu = {"value": len(sample) / dt, "unit": "frames/s", "cores": cores, "kind": kind,
                   "sample": f"{len(sample)} frames: reference ORBextractor.cc compiled in place + restated SearchByProjection on {cores} host threads, {dt:.1f} s"}
        print(json.dumps({
            "metric": metric, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": f"configs[2]: TUM-shape 640x480 frames, nFeatures=1000: extraction + keypoint grid + SearchByProjection "
                                   f"(th=3, nnratio=0.8) against {NM} map points per frame", "frames_per_step_per_gpu": F,
                       "parallelism": f"frames sharded over {world} GPU(s), no collective",
                       "l2": f"working set per step ~{F * (5.7 + NM * 56e-6):.0f} MB (> 126 MB L2)" , "mean_matches_per_frame": mean_matches,
                       "search_only_ms_per_step": ms_search, "search_only_point_queries_per_s": F * NM / (ms_search * 1e-3)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": F * Himg * Wimg,
                    "d2h_bytes_per_step": F * (cap * 60 + 4) + F * fs.cap * 4 + F * 4, "steps": e2e_steps,
                    "api": "obs_extract_batch (page-locked host images in, keypoints + descriptors out) + obs_frame_set_from_extractor + "
                           "obs_search_by_projection (assignment out)"},
            "gpu_launches": (12 + 1 + 2) * args.steps,
            "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": F * 5.742474e6 / (ms_total / args.steps * 1e-3) / 1e9,
                         "peak": _hbm_peak(), "unit": "GB/s", "frac": F * 5.742474e6 / (ms_total / args.steps * 1e-3) / 1e9 / _hbm_peak(),
                         "traffic": None, "note": "extraction dominates; the search alone is reported in config.search_only_*"},
            "cpu_baseline": cpu}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


# ------------------------------------------------------------------------------------------------ bow
def cpu_bow(pairs, cores, ratio):
    import oracle
    def one(p):
        return oracle.search_by_bow(p[0], p[1], TH_LOW, True, ratio, True)[0]
    with ThreadPoolExecutor(cores) as pool:
        list(pool.map(one, pairs[:cores]))
        t0 = time.perf_counter()
        list(pool.map(one, pairs))
        return time.perf_counter() - t0


def run_bow(args, rank, local_rank, world, ClockSampler):
    """SearchByBoW(KeyFrame*, KeyFrame*) (ORBmatcher.cc:522-655) over batches of synthetic keyframe pairs: 2000 keypoints, 100
    vocabulary nodes, ratio 0.8; pairs sharded over the ranks, no collective (weak scaling)."""
    NKP, NODES, RAT = 2000, 100, 0.8
    Bp = args.frames                                           # keyframe pairs per step per GPU
    metric = "keyframe pairs/s, ORBmatcher::SearchByBoW(KF, KF), 2000 keypoints per keyframe, 100 vocabulary nodes"
    uniq = [synth.bow_pair(synth.KITTI_SHAPE, NKP, 500 + i, n_nodes=NODES)[:2] for i in range(8)]
    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        sp = [uniq[i % 8] for i in range(max(512 * cores, 2048))]
        total = 0.0
        for _ in range(args.warmup):
            cpu_bow(sp[:cores], cores, RAT)
        for _ in range(args.steps):
            total += cpu_bow(sp, cores, RAT)
        v = len(sp) * args.steps / total
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "u32 popcount", "data": "synthetic",
                          "config": {"workload": "SearchByBoW keyframe pairs", "pairs_per_step": len(sp)},
                          "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port",
                                           "sample": f"{len(sp)} keyframe pairs per step, restated SearchByBoW (oracle/match_oracle.cpp)"},
                          "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    import torch
    import torch.distributed as dist
    from object_slam_b200._capi import BowSide, check, lib
    from object_slam_b200.matcher import ORBmatcher
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    M = ORBmatcher(RAT, True, device=local_rank)
    s1, k1 = M._pack_side([uniq[(rank * Bp + i) % 8][0] for i in range(Bp)])
    s2, k2 = M._pack_side([uniq[(rank * Bp + i) % 8][1] for i in range(Bp)])

    def to_dev(side, pk):
        keep = {k: torch.from_numpy(np.ascontiguousarray(v).view(np.uint8).reshape(-1)).to(dev) for k, v in pk.items() if v is not None}
        d = BowSide(side.cap, side.node_cap, *[keep[k].data_ptr() if k in keep else None
                                               for k in ("n", "descriptors", "keys_un", "valid", "u_right", "n_nodes", "node_id", "node_start", "node_idx")])
        return d, keep
    d1, keep1 = to_dev(s1, k1)
    d2, keep2 = to_dev(s2, k2)
    m12 = torch.empty((Bp, s1.cap), dtype=torch.int32, device=dev); m21 = torch.empty((Bp, s2.cap), dtype=torch.int32, device=dev)
    nm = torch.empty(Bp, dtype=torch.int32, device=dev)
    h12 = np.empty((Bp, s1.cap), np.int32); h21 = np.empty((Bp, s2.cap), np.int32); hn = np.empty(Bp, np.int32)

    def step():
        check(lib().obs_search_by_bow(M._h, C.byref(d1), C.byref(d2), Bp, TH_LOW, 1, RAT, 1, C.c_void_p(m12.data_ptr()),
                                      C.c_void_p(m21.data_ptr()), C.c_void_p(nm.data_ptr())))

    def step_host():
        check(lib().obs_search_by_bow(M._h, C.byref(s1), C.byref(s2), Bp, TH_LOW, 1, RAT, 1, h12.ctypes.data_as(C.c_void_p),
                                      h21.ctypes.data_as(C.c_void_p), hn.ctypes.data_as(C.c_void_p)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ms_stream = torch.cuda.ExternalStream(M.stream)
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ms_stream):
        e0.record()
    for _ in range(args.steps):
        step()
    with torch.cuda.stream(ms_stream):
        e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t[0])
    value = world * Bp * args.steps / (ms_total * 1e-3)
    step_host()
    barrier()
    e2e_steps = max(3, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * Bp * e2e_steps / float(t.item())
    if rank == 0:
        # distance evaluations of one step: per common node, valid keypoints of side 1 x keypoints of side 2 (an upper bound on what the
        # kernel evaluates: matched keypoints of side 2 are skipped)
        cand = 0
        for a, b in uniq:
            na = {int(a["node_id"][k]): (int(a["node_start"][k]), int(a["node_start"][k + 1])) for k in range(len(a["node_id"]))}
            for k in range(len(b["node_id"])):
                r = na.get(int(b["node_id"][k]))
                if r:
                    cand += int(a["valid"][a["node_idx"][r[0]:r[1]]].sum()) * int(b["valid"][b["node_idx"][b["node_start"][k]:b["node_start"][k + 1]]].sum())
        cand_per_pair = cand / len(uniq)
        popc_only = _popc_peak(2)
        ach = 8.0 * cand_per_pair * Bp / (ms_total / args.steps * 1e-3)
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            sp = [uniq[i % 8] for i in range(max(1024 * cores, 4096))]
            dt = cpu_bow(sp, cores, RAT)
            cpu = {"value": len(sp) / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
                   "sample": f"{len(sp)} keyframe pairs on {cores} host threads, {dt:.2f} s; restated SearchByBoW (oracle/match_oracle.cpp)"}
        print(json.dumps({
            "metric": metric, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 popcount",
            "data": "synthetic",
            "config": {"workload": f"SearchByBoW(KF, KF): {Bp} keyframe pairs per step per GPU, {NKP} keypoints, {NODES} nodes, ratio {RAT}, TH_LOW 50, "
                                   "orientation check", "parallelism": f"pairs sharded over {world} GPU(s), no collective",
                       "candidate_distances_per_pair": cand_per_pair, "accepted_matches_first_pair": int(nm[0].item())},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(sum(v.nbytes for v in list(k1.values()) + list(k2.values()) if v is not None)),
                    "d2h_bytes_per_step": int(h12.nbytes + h21.nbytes + hn.nbytes), "steps": e2e_steps,
                    "api": "obs_search_by_bow with host arrays in and out (pageable numpy buffers)"},
            "gpu_launches": args.steps,
            "roofline": {"bound": "int-popc", "kernel": "k_bow_search", "achieved": ach / 1e12, "peak": popc_only * 8e9 / 1e12, "unit": "Tpopc/s",
                         "frac": ach / (popc_only * 8e9), "traffic": None,
                         "peak_source": "obs_microbench_popc mode 2 measured in this run; the kernel is latency bound (one warp per vocabulary "
                                        "node, sequential over the node's keypoints as the reference's order demands), not popc bound"},
            "cpu_baseline": cpu}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_extract(imgs, cores, nfeat):
    """The reference extractor (oracle/_ref) on every image, `cores` host threads.  Returns (seconds, kind)."""
    import oracle
    import threading
    use_ref = oracle.ref_available()
    tl = threading.local()

    def one(img):
        if not hasattr(tl, "ex"):
            tl.ex = oracle.ReferenceExtractor(nfeat) if use_ref else oracle.OracleExtractor(nfeat)
        return len(tl.ex(img)[0])

    with ThreadPoolExecutor(cores) as pool:
        list(pool.map(one, imgs[:cores]))
        t0 = time.perf_counter()
        list(pool.map(one, imgs))
        return time.perf_counter() - t0, ("reference" if use_ref else "port")


def run_extract(args, rank, local_rank, world, ClockSampler):
    """configs[3]: batched offline ORB extraction of `--total-frames` (32768) KITTI-shape frames.  The frame range is cut into
    contiguous shards, one per rank (sharding.shard_range), no collective; a step is one batch of F frames per rank, so the job
    is total / (F * world) steps per rank.  The frames are a resident pool of distinct synthetic images (seed = global frame
    index of the first `pool` frames of the shard) cycled over the shard: 32768 KITTI frames are 15 GB of numpy generation."""
    Himg, Wimg, NF, PITCH = 376, 1241, 2000, 1280
    F = args.frames
    T = args.total_frames
    metric = "frames/s, batched offline ORB extraction, KITTI 1241x376, 2000 kp"
    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        imgs = [synth.blocky_image((Himg, Wimg), i) for i in range(2 * cores)]
        for _ in range(args.warmup):
            cpu_extract(imgs[:cores], cores, NF)
        total, kind = 0.0, "port"
        for _ in range(args.steps):
            dt, kind = cpu_extract(imgs, cores, NF)
            total += dt
        v = len(imgs) * args.steps / total
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": {"workload": "configs[3] batched offline ORB extraction, KITTI shape", "frames_per_step": len(imgs)},
                          "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": kind,
                                           "sample": f"{len(imgs)} frames per step; reference ORBextractor.cc compiled in place, {cores} host threads"},
                          "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    import threading
    import torch
    import torch.distributed as dist
    from object_slam_b200._capi import pinned_empty, KEYPOINT_DTYPE
    from object_slam_b200.extractor import ORBextractor
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: object_slam_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = sharding.shard_range(T, rank, world)
    POOL = min(args.pool_batches * F, hi - lo)
    POOL = max(POOL // F, 1) * F
    with ThreadPoolExecutor(os.cpu_count() or 1) as tp:
        imgs = list(tp.map(lambda i: synth.blocky_image((Himg, Wimg), lo + i), range(POOL)))
    host = np.zeros((POOL, Himg, PITCH), np.uint8)
    for i, im in enumerate(imgs):
        host[i, :, :Wimg] = im
    dimg = torch.from_numpy(host).to(dev)
    NB = POOL // F
    mk = lambda: ORBextractor(NF, 1.2, 8, 20, 7, max_size=(Wimg, Himg), max_batch=F, device=local_rank)
    NPIPES = int(os.environ.get("OBS_BENCH_PIPES", "3"))
    dpipes = [(mk(), torch.cuda.Stream()) for _ in range(NPIPES)]
    main_stream = torch.cuda.Stream()
    torch.cuda.set_stream(main_stream)
    step_no = [0]

    def step_device():
        ex, st = dpipes[step_no[0] % NPIPES]
        b = step_no[0] % NB
        step_no[0] += 1
        ex.extract_device(dimg.data_ptr() + b * F * Himg * PITCH, F, Wimg, Himg, PITCH, Himg * PITCH, st.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, NPIPES, 3)):
        step_device()
    step_no[0] = 0
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _, st in dpipes:
        st.wait_stream(main_stream)
    for _ in range(args.steps):
        step_device()
    for _, st in dpipes:
        main_stream.wait_stream(st)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    counts = dpipes[0][0].fetch_counts()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * F * args.steps / (ms_total * 1e-3)

    # ---- end to end: page-locked host images in, keypoints + descriptors out; WORKERS pipelines take alternate batches
    WORKERS = args.e2e_pipelines
    cap = dpipes[0][0].capacity

    class Pipe:
        def __init__(self, ex, b):
            self.ex = ex
            self.pin = pinned_empty((F, Himg, Wimg), np.uint8)
            for i in range(F):
                self.pin[i] = imgs[(b * F + i) % POOL]
            self.out = (pinned_empty((F, cap), KEYPOINT_DTYPE), pinned_empty((F, cap, 32), np.uint8), pinned_empty((F,), np.int32))

        def step(self):
            return self.ex.extract_batch(self.pin, out=self.out, copy=False)

    pipes = [Pipe(dpipes[i % NPIPES][0] if i < NPIPES else mk(), i) for i in range(WORKERS)]
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
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * F * e2e_steps / float(t.item())
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            sample = [imgs[i % POOL] for i in range(16 * cores)]
            dt, kind = cpu_extract(sample, cores, NF)
            cpu = {"value": len(sample) / dt, "unit": "frames/s", "cores": cores, "kind": kind,
                   "sample": f"{len(sample)} frames of the pool: reference src/ORBextractor.cc compiled in place (oracle/_ref) on {cores} host threads, {dt:.1f} s"}
        step_ms = ms_total / args.steps
        eye_bytes = 9359539.0
        peak = _hbm_peak()
        print(json.dumps({
            "metric": metric, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"configs[3]: batched offline ORB extraction of {T} synthetic KITTI-shape 1241x376 frames, nFeatures=2000, "
                                   f"nLevels=8, scale 1.2, FAST 20/7, contiguous frame shards over the ranks",
                       "frames_per_step_per_gpu": F, "job_steps_per_gpu": (hi - lo + F - 1) // F,
                       "job_seconds_at_this_rate": T / value,
                       "pool": f"{POOL} distinct frames per rank resident in HBM ({POOL * Himg * PITCH / 1e6:.0f} MB), cycled over the shard",
                       "parallelism": f"frames sharded over {world} GPU(s), no collective",
                       "l2": f"working set per step {F * 4.0:.0f} MB of images, pyramids and blurred levels (> 126 MB L2), pool of inputs larger than L2",
                       "mean_keypoints": float(counts.mean())},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": F * Himg * Wimg, "d2h_bytes_per_step": F * (cap * 60 + 4) + F * 4,
                    "steps": e2e_steps, "api": f"obs_extract_batch, page-locked host images in, keypoints + descriptors out; {WORKERS} pipelines "
                                               "take alternate batches from their own host threads"},
            "gpu_launches": 8 * args.steps,
            "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": F * eye_bytes / (step_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": F * eye_bytes / (step_ms * 1e-3) / 1e9 / peak, "traffic": None,
                         "note": "9,359,539 algorithmic bytes per KITTI image (SURVEY 8d); the kernels are integer-issue bound (DESIGN.md section 4), "
                                 "the per-kernel roofline is on the headline line (--workload stereo)"},
            "cpu_baseline": cpu}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run(args, rank, local_rank, world, ClockSampler):
    if args.workload == "bow":
        return run_bow(args, rank, local_rank, world, ClockSampler)
    if args.workload == "extract":
        return run_extract(args, rank, local_rank, world, ClockSampler)
    if args.workload == "knn2":
        return run_knn2(args, rank, local_rank, world, ClockSampler)
    return run_projection(args, rank, local_rank, world, ClockSampler)
