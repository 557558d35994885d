"""The other workloads of bench.py: configs[3] (extract) and configs[4] (knn2) -- measured as sub-results of the default line and
available alone with `--workload` -- plus configs[2] (projection) and the BoW search; same JSON contract as the headline line.

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
import math
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


def run_projection(args, rank, local_rank, world, ClockSampler, as_sub=False):
    """configs[2].  as_sub: return the result line (rank 0; None elsewhere) instead of printing it -- bench.py's default line carries it
    as sub["configs[2]"]."""
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
    if world > 1 and not dist.is_initialized():
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
    # the timed region is R back-to-back repeats of the K steps asked for, R chosen so that it lasts >= 1 s (calibrated on one repeat)
    with torch.cuda.stream(ms_stream):
        e0.record()
    for _ in range(args.steps):
        step_device()
    with torch.cuda.stream(ms_stream):
        e1.record()
    barrier()
    tcal = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tcal, op=dist.ReduceOp.MAX)
    reps = max(1, int(math.ceil(1000.0 / max(float(tcal.item()), 1e-3))))
    timed_steps = reps * args.steps
    with torch.cuda.stream(ms_stream):
        e0.record()
    for _ in range(timed_steps):
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
    value = world * F * timed_steps / (ms_total * 1e-3)
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
        line = {
            "metric": metric, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / timed_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": f"configs[2]: TUM-shape 640x480 frames, nFeatures=1000: extraction + keypoint grid + SearchByProjection "
                                   f"(th=3, nnratio=0.8) against {NM} map points per frame", "frames_per_step_per_gpu": F,
                       "parallelism": f"frames sharded over {world} GPU(s), no collective",
                       "timed_steps": timed_steps, "repeats": reps, "timed_region_s": ms_total * 1e-3,
                       "l2": f"working set per step ~{F * (5.7 + NM * 56e-6):.0f} MB (> 126 MB L2)" , "mean_matches_per_frame": mean_matches,
                       "search_only_ms_per_step": ms_search, "search_only_point_queries_per_s": F * NM / (ms_search * 1e-3)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": F * Himg * Wimg,
                    "d2h_bytes_per_step": F * (cap * 60 + 4) + F * fs.cap * 4 + F * 4, "steps": e2e_steps,
                    "api": "obs_extract_batch (page-locked host images in, keypoints + descriptors out) + obs_frame_set_from_extractor + "
                           "obs_search_by_projection (assignment out)"},
            "gpu_launches": (12 + 1 + 2) * timed_steps,
            "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": F * 5.742474e6 / (ms_total / timed_steps * 1e-3) / 1e9,
                         "peak": _hbm_peak(), "unit": "GB/s", "frac": F * 5.742474e6 / (ms_total / timed_steps * 1e-3) / 1e9 / _hbm_peak(),
                         "traffic": None, "note": "extraction dominates; the search alone is reported in config.search_only_*"},
            "cpu_baseline": cpu}
        if as_sub:
            return line
        print(json.dumps(line), flush=True)
    return None


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
    if world > 1 and not dist.is_initialized():
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


def _expected_hash(key):
    try:
        with open(os.path.join(ROOT, "tests", "golden", "bench_hashes.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


def _hash_check(key, value):
    exp = _expected_hash(key)
    return {"value": f"{value:016x}", "n1_hash": exp, "equals_n1_hash": (exp == f"{value:016x}") if exp else None,
            "source": "tests/golden/bench_hashes.json (written from a 1-GPU run by tools/update_bench_hashes.py)"}


# ------------------------------------------------------------------------------------------------ configs[3]
EXTRACT_METRIC = "frames/s, batched offline ORB extraction, KITTI 1241x376, 2000 kp"


def extract_workload(T):
    return (f"configs[3]: batched offline ORB extraction of {T} synthetic KITTI-shape 1241x376 frames, nFeatures=2000, nLevels=8, "
            "scale 1.2, FAST 20/7, contiguous frame shards over the ranks")


def _record_hashes(torch, ex, F, dev):
    """One 64-bit hash per image of the last extraction on `ex`, computed on the device from the result records: every valid
    keypoint field and descriptor word times a position-dependent odd multiplier, summed modulo 2^64 (int64 wrap-around)."""
    d_rec, rec_bytes, cap = ex.results_device()
    ints = rec_bytes // 4
    # view the library's record buffer without copying: a tensor over foreign device memory
    rec = _foreign_i32(torch, d_rec, F * ints, dev).view(F, ints)
    n = rec[:, 0].to(torch.int64)
    kp = rec[:, 16:16 + cap * 7].view(F, cap, 7).to(torch.int64)
    ds = rec[:, 16 + cap * 7:16 + cap * 15].view(F, cap, 8).to(torch.int64)
    idx = torch.arange(cap, device=dev, dtype=torch.int64)
    valid = (idx[None, :] < n[:, None]).to(torch.int64)
    wk = (idx[:, None] * 7 + torch.arange(7, device=dev, dtype=torch.int64)[None, :]) * 2654435761 + 1
    wd = (idx[:, None] * 8 + torch.arange(8, device=dev, dtype=torch.int64)[None, :]) * 40503 + 12345678901
    h = ((kp * wk[None]).sum(2) * valid).sum(1) + ((ds * wd[None]).sum(2) * valid).sum(1) + n * 1099511628211
    return h


class _CudaArray:
    """__cuda_array_interface__ over a raw device pointer (no ownership)."""
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (int(ptr), False), "version": 3}


def _foreign_i32(torch, ptr, n, dev):
    return torch.as_tensor(_CudaArray(ptr, n), device=dev)


def measure_extract(args, ctx, ClockSampler):
    """configs[3]: batched offline ORB extraction of `--total-frames` (32768) KITTI-shape frames.  The frame range is cut into
    contiguous shards, one per rank (sharding.shard_range), no data-path collective; a step is one batch of F frames and the job is
    shard / F steps per rank.  Global frame g is the synthetic image of seed g mod POOL (POOL = 512 distinct images, generated once --
    each rank makes POOL / world of them and the pool is all-gathered over NCCL -- and kept resident; 32768 distinct KITTI frames
    would be 15 GB of numpy generation).  `value` times the WHOLE job (repeated to fill >= 1 s); the job hash is the sum over all
    global frames of a per-frame hash of keypoints and descriptors, so it does not depend on the number of GPUs."""
    torch, dist = ctx.torch, ctx.dist
    from object_slam_b200._capi import pinned_empty, KEYPOINT_DTYPE
    from object_slam_b200.extractor import ORBextractor
    rank, local_rank, world, dev = ctx.rank, ctx.local_rank, ctx.world, ctx.dev
    Himg, Wimg, NF, PITCH = 376, 1241, 2000, 1280
    F, T = args.frames, args.total_frames
    POOL = max(args.pool // F, 1) * F
    lo, hi = sharding.shard_range(T, rank, world)
    # the pool: every rank generates its slice, NCCL spreads it
    own_lo, own_hi = sharding.shard_range(POOL, rank, world)
    with ThreadPoolExecutor(len(ctx.all_cpus) // max(world, 1) or 1) as tp:
        mine = list(tp.map(lambda i: synth.blocky_image((Himg, Wimg), i), range(own_lo, own_hi)))
    dimg = torch.zeros((POOL, Himg, PITCH), dtype=torch.uint8, device=dev)
    if mine:
        dimg[own_lo:own_hi, :, :Wimg] = torch.from_numpy(np.stack(mine)).to(dev)
    if world > 1:
        dist.all_reduce(dimg.view(torch.int32), op=dist.ReduceOp.SUM)        # slices are disjoint, the rest is zero
    NB = POOL // F
    steps_job = max((hi - lo) // F, 1)
    mk = lambda: ORBextractor(NF, 1.2, 8, 20, 7, max_size=(Wimg, Himg), max_batch=F, device=local_rank)
    NPIPES = int(os.environ.get("OBS_BENCH_PIPES", "3"))
    dpipes = [(mk(), torch.cuda.Stream()) for _ in range(NPIPES)]
    main_stream = torch.cuda.current_stream()
    first_batch = (lo // F) % NB

    def job_device(nrep):
        for s in range(nrep * steps_job):
            ex, st = dpipes[s % NPIPES]
            b = (first_batch + s % steps_job) % NB
            ex.extract_device(dimg.data_ptr() + b * F * Himg * PITCH, F, Wimg, Himg, PITCH, Himg * PITCH, st.cuda_stream)

    def timed(nrep):
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _, st in dpipes:
            st.wait_stream(main_stream)
        job_device(nrep)
        for _, st in dpipes:
            main_stream.wait_stream(st)
        e1.record()
        ctx.barrier()
        return e0.elapsed_time(e1)

    for s in range(max(args.warmup, NPIPES, 3)):
        dpipes[s % NPIPES][0].extract_device(dimg.data_ptr(), F, Wimg, Himg, PITCH, Himg * PITCH, dpipes[s % NPIPES][1].cuda_stream)
    reps = ctx.repeats(timed(1))
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total = ctx.reduce([timed(reps)])[0]
    clocks = sampler.stop() if sampler else None
    value = T * reps / (ms_total * 1e-3)

    # ---- the job hash: per-image hashes of the POOL images, each weighted by how often the shard visits it
    ex0, st0 = dpipes[0]
    hsum = 0
    kp_total = 0
    with torch.cuda.stream(st0):
        for b in range(NB):
            ex0.extract_device(dimg.data_ptr() + b * F * Himg * PITCH, F, Wimg, Himg, PITCH, Himg * PITCH, st0.cuda_stream)
            h = _record_hashes(torch, ex0, F, dev)
            visits = sum(1 for s in range(steps_job) if (first_batch + s) % NB == b)
            hsum = (hsum + visits * int(h.sum().item())) & 0xffffffffffffffff
            kp_total += int(ex0.fetch_counts().sum())
    job_hash = ctx.sum_u64(hsum)

    # ---- end to end: the shard's frames from page-locked host memory in, keypoints + descriptors out to page-locked memory, the
    # whole job; one host thread keeps WORKERS handles in flight (obs_extract_batch_submit / _wait)
    WORKERS = args.e2e_pipelines
    cap = ex0.capacity
    pool_host = pinned_empty((POOL, Himg, Wimg), np.uint8)
    pool_host[:] = dimg[:, :, :Wimg].cpu().numpy()
    outs = [(pinned_empty((F, cap), KEYPOINT_DTYPE), pinned_empty((F, cap, 32), np.uint8), pinned_empty((F,), np.int32)) for _ in range(WORKERS)]
    exs = [dpipes[i % NPIPES][0] if i < NPIPES else mk() for i in range(WORKERS)]

    def job_host(nrep):
        n = nrep * steps_job
        for s in range(n + WORKERS):
            k = s % WORKERS
            if s >= WORKERS:
                exs[k].wait_batch()
            if s < n:
                b = (first_batch + s % steps_job) % NB
                exs[k].submit_batch(pool_host[b * F:(b + 1) * F], *outs[k])

    torch.cuda.synchronize()
    job_host(1) if steps_job >= 2 * WORKERS else job_host(2)
    ctx.barrier()
    t0 = time.perf_counter()
    job_host(1)
    e2e_reps = ctx.repeats(1e3 * (time.perf_counter() - t0))
    ctx.barrier()
    t0 = time.perf_counter()
    job_host(e2e_reps)
    torch.cuda.synchronize()
    e2e_s = ctx.reduce([time.perf_counter() - t0])[0]
    e2e_value = T * e2e_reps / e2e_s
    if rank != 0:
        return None
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        ctx.release_cpus()
        cores = os.cpu_count() or 1
        sample = [np.ascontiguousarray(pool_host[i % POOL]) for i in range(8 * cores)]
        dt, kind = cpu_extract(sample, cores, NF)
        cpu = {"value": len(sample) / dt, "unit": "frames/s", "cores": cores, "kind": kind,
               "sample": f"{len(sample)} frames of the pool: reference src/ORBextractor.cc compiled in place (oracle/_ref) on {cores} host threads, {dt:.1f} s"}
    job_ms = ms_total / reps
    eye_bytes = 9359539.0
    peak = _hbm_peak()
    return {
        "metric": EXTRACT_METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps_job, "warmup": max(args.warmup, 3),
        "ms_per_step": job_ms / steps_job, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": extract_workload(T), "frames_per_step_per_gpu": F, "job_steps_per_gpu": steps_job,
                   "job_ms": job_ms, "repeats": reps, "timed_region_s": ms_total * 1e-3,
                   "pool": f"{POOL} distinct frames (seed = global frame index mod {POOL}) resident in HBM ({POOL * Himg * PITCH / 1e6:.0f} MB), "
                           "generated once across the ranks and spread over NCCL",
                   "parallelism": f"frames sharded over {world} GPU(s), no data-path collective",
                   "l2": f"working set per step {F * 4.0:.0f} MB of images, pyramids and blurred levels (> 126 MB L2), pool of inputs larger than L2",
                   "mean_keypoints": kp_total / POOL},
        "result_hash": _hash_check(f"extract:T={T}:pool={POOL}", job_hash),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": F * Himg * Wimg, "d2h_bytes_per_step": F * (cap * 60 + 4),
                "steps": e2e_reps * steps_job, "seconds": e2e_s,
                "api": f"obs_extract_batch_submit / _wait, page-locked host images in, keypoints + descriptors out; one host thread keeps "
                       f"{WORKERS} handles in flight"},
        "gpu_launches": 8 * reps * steps_job,
        "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": T * eye_bytes / (job_ms * 1e-3) / 1e9 / world, "peak": peak, "unit": "GB/s",
                     "frac": T * eye_bytes / (job_ms * 1e-3) / 1e9 / world / peak, "traffic": None,
                     "note": "per GPU; 9,359,539 algorithmic bytes per KITTI image (SURVEY 8d); the kernels are integer-issue bound (DESIGN.md "
                             "section 4), the per-kernel fractions are on the headline line"},
        "cpu_baseline": cpu}


# ------------------------------------------------------------------------------------------------ configs[4]
KNN2_METRIC = "Hamming distance evaluations/s, brute-force keyframe-vs-keyframe matching (best + second best + ratio), 2000 descriptors per keyframe"


def knn2_workload(K, Wn):
    return (f"configs[4]: {K} keyframes x {NDESC} descriptors, every keyframe against its +-{Wn} neighbours, ratio 0.6, TH_LOW 50, "
            "query keyframes sharded over the ranks, NCCL all-gather of the descriptor sets")


def _imma_peak(local_rank):
    from object_slam_b200._capi import check, lib
    g = C.c_double()
    check(lib().obs_microbench_imma(int(local_rank), C.byref(g)))
    return g.value


def measure_knn2(args, ctx, ClockSampler):
    """configs[4]: K keyframes x 2000 descriptors, every keyframe matched against its +-window neighbours (best / second best / ratio
    test).  Query keyframes are sharded over the ranks (strong scaling: the total is fixed).  Every rank needs all descriptor sets, so
    each step starts with the all-gather of the descriptor shards over NCCL (csrc/comm.cu) in `--gather-chunks` chunks; pairs whose
    database keyframe is local are matched while the gather is in flight, the others as soon as their chunk has arrived.  The result
    hash weights every best index with (query keyframe, database keyframe, row), so it does not depend on the sharding."""
    torch, dist = ctx.torch, ctx.dist
    from object_slam_b200._capi import check, lib, pinned_empty
    from object_slam_b200.matcher import ORBmatcher
    rank, local_rank, world, dev = ctx.rank, ctx.local_rank, ctx.world, ctx.dev
    K, Wn = args.keyframes, args.window
    NC = max(args.gather_chunks, 1) if world > 1 else 1
    per = sharding.padded_shard(K, world)
    while NC > 1 and per % NC:
        NC -= 1                                                # chunk boundaries on whole keyframes
    lo, hi = rank * per, min(K, (rank + 1) * per)
    full = _device_keyframes(torch, K, 1234, dev)              # every rank generates the same set, keeps only its shard
    local = torch.zeros((per, NDESC, 32), dtype=torch.uint8, device=dev)
    local[:hi - lo] = full[lo:hi]
    del full
    allD = torch.zeros((world * per, NDESC, 32), dtype=torch.uint8, device=dev)
    pairs = sharding.window_pairs(lo, hi, K, Wn)
    groups = sharding.split_by_chunk(pairs, lo, hi, per, NC)   # [local, chunk 0, chunk 1, ...]
    order = np.concatenate([g for g in groups if len(g)]) if len(pairs) else pairs
    P = len(pairs)
    dgroups = [torch.from_numpy(g).to(dev) if len(g) else None for g in groups]
    best = torch.empty((max(P, 1), NDESC), dtype=torch.int32, device=dev)
    M = ORBmatcher(RATIO, True, device=local_rank)
    M.set_knn2_engine({"auto": M.KNN2_AUTO, "tensor": M.KNN2_TENSOR, "popc": M.KNN2_POPC}[args.knn2_engine])
    check(lib().obs_set_option(b"knn2_cta_pair", int(getattr(args, "knn2_cta_pair", 1))))
    tensor = args.knn2_engine != "popc"
    mstream = M.stream
    comm = None
    if world > 1:
        def bcast(raw):
            t = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
            dist.broadcast(t, 0)
            return bytes(t.cpu().tolist())
        comm = sharding.Comm(rank, world, local_rank, bcast)
    local_bytes = per * NDESC * 32

    def knn(dpairs, out_off):
        n = 0 if dpairs is None else dpairs.shape[0]
        if n:
            check(lib().obs_hamming_knn2(M._h, C.c_void_p(allD.data_ptr()), world * per, NDESC, C.c_void_p(dpairs.data_ptr()), n,
                                         TH_LOW, RATIO, C.c_void_p(best.data_ptr() + out_off * NDESC * 4), None, None))
        return n

    mine = allD[rank * per:(rank + 1) * per]                  # this rank's shard lives in its slot of the gathered set
    mine.copy_(local)
    torch.cuda.synchronize()
    launches = [0]

    def step():
        # the gather is ordered after everything queued on the matcher stream (the producer of `mine` and the previous step's
        # readers of the remote slots) and runs on the communicator's stream; local pairs are matched meanwhile, the pairs of a
        # remote chunk wait for that chunk only
        if comm:
            comm.allgather(mine.data_ptr(), local_bytes, allD.data_ptr(), NC, mstream)
        off = knn(dgroups[0], 0)
        for c in range(NC):
            if comm:
                comm.wait(c, mstream)
            off += knn(dgroups[1 + c], off) if len(dgroups) > 1 + c else 0

    ms_stream = torch.cuda.ExternalStream(mstream)

    def timed(n):
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ms_stream):
            e0.record()
        for _ in range(n):
            step()
        with torch.cuda.stream(ms_stream):
            e1.record()
        ctx.barrier()
        return e0.elapsed_time(e1)

    for _ in range(max(args.warmup, 3)):
        step()
    steps = max(min(args.steps, 10), 1)
    reps = ctx.repeats(timed(steps))
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total = ctx.reduce([timed(reps * steps)])[0]
    clocks = sampler.stop() if sampler else None
    timed_steps = reps * steps
    P_all = int(ctx.reduce([P], "sum")[0])
    dist_per_step = P_all * NDESC * NDESC
    value = dist_per_step * timed_steps / (ms_total * 1e-3)

    # ---- result hash over (query keyframe, database keyframe, row, best index)
    hsum = 0
    if P:
        dorder = torch.from_numpy(order.astype(np.int64)).to(dev)
        pw = dorder[:, 0] * 1000003 + dorder[:, 1] * 7919 + 1
        cw = torch.arange(NDESC, device=dev, dtype=torch.int64) * 2654435761 + 12345
        for a in range(0, P, 2048):
            b = min(P, a + 2048)
            hsum = (hsum + int((((best[a:b].to(torch.int64) + 2) * cw[None, :]).sum(1) * pw[a:b]).sum().item())) & 0xffffffffffffffff
    matches = int(ctx.reduce([int((best[:P] >= 0).sum())], "sum")[0])
    job_hash = ctx.sum_u64(hsum)

    # ---- end to end: descriptor shard from page-locked host memory in, best indices out to page-locked memory, every step
    h_local = pinned_empty((per, NDESC, 32), np.uint8)
    h_local[:] = local.cpu().numpy()
    h_best = pinned_empty((max(P, 1), NDESC), np.int32)

    def step_host():
        with torch.cuda.stream(ms_stream):
            mine.copy_(torch.from_numpy(h_local), non_blocking=True)
        step()
        with torch.cuda.stream(ms_stream):
            torch.from_numpy(h_best).copy_(best, non_blocking=True)
        M.sync()

    step_host()
    ctx.barrier()
    t0 = time.perf_counter()
    step_host()
    e2e_steps = ctx.repeats(1e3 * (time.perf_counter() - t0))
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    e2e_s = ctx.reduce([time.perf_counter() - t0])[0]
    e2e_value = dist_per_step * e2e_steps / e2e_s
    if comm:
        comm.close()
    if rank != 0:
        return None
    kernel_ms = ms_total / timed_steps
    if tensor:
        peak_tops = _imma_peak(local_rank)
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                bf16 = float(json.load(f)["bf16_tflops"])
        except Exception:
            bf16 = 1590.0
        ach = 512.0 * (dist_per_step / world) / (kernel_ms * 1e-3) / 1e12     # 256 multiply-adds = 512 operations per distance
        clk = (clocks or {})
        clk_ratio = (clk.get("sm_mhz") or 0) / (clk.get("sm_max_mhz") or 1) if clk.get("sm_mhz") else None
        roof = {"bound": "tensor", "kernel": "k_knn2_tc_pair (tcgen05.mma.cta_group::2)", "achieved": ach, "peak": peak_tops, "unit": "TOP/s (int8)",
                "frac": ach / peak_tops,
                # the run is power-capped (clocks.reasons: sw_power_cap): the same peak scaled to the SM clock the timed region ran at
                "frac_at_sustained_clock": (ach / (peak_tops * clk_ratio)) if clk_ratio else None, "sustained_clock_ratio": clk_ratio,
                "traffic": None,
                "peak_source": "obs_microbench_imma measured in this run: tcgen05.mma kind::i8 M128 N256 K32 from shared memory issued back to "
                               "back on all SMs (MEASURED_PEAKS.json holds no int8 figure)",
                "vs_2x_measured_bf16": ach / (2 * bf16), "measured_bf16_tflops": bf16,
                "per_gpu_distances_per_s": value / world,
                "popc_pipe_ceiling_distances_per_s": _popc_peak(2) * 1e9,
                "note": "hamming = (256 - a.b) / 2 on +-1 int8 operands: 512 tensor operations per distance; the POPC pipe ceiling "
                        "(8 popc per distance) is what the non-tensor kernel is bound by"}
    else:
        popc_only = _popc_peak(2)
        peak_popc = popc_only * 8e9
        ach_popc = 5.0 * (dist_per_step / world) / (kernel_ms * 1e-3)
        roof = {"bound": "int-popc", "kernel": "k_knn2", "achieved": ach_popc / 1e12, "peak": peak_popc / 1e12, "unit": "Tpopc/s",
                "frac": ach_popc / peak_popc, "traffic": None,
                "peak_source": "obs_microbench_popc mode 2 measured in this run (POPC pipe: 16 lanes/clk/SM)"}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        ctx.release_cpus()
        cores = os.cpu_count() or 1
        Dh = synth.keyframe_descriptors(8, NDESC, 0)
        sp = sharding.window_pairs(0, 8, 8, 8)[:max(2 * cores, 8)]
        dt = cpu_knn2(Dh, sp, cores)
        cpu = {"value": len(sp) * NDESC * NDESC / dt, "unit": "distances/s", "cores": cores, "kind": "port",
               "sample": f"{len(sp)} keyframe pairs (2000 x 2000 distances each) on {cores} host threads, {dt:.1f} s; "
                         "SearchByBoW candidate loop with the reference's SWAR DescriptorDistance"}
    return {
        "metric": KNN2_METRIC, "value": value, "unit": "distances/s", "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
        "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int8 (+-1 per descriptor bit), int32 accumulate" if tensor else "u32 popcount", "data": "synthetic (generated on the device)",
        "config": {"workload": knn2_workload(K, Wn), "pairs_per_step": P_all, "gather_chunks": NC,
                   "timed_steps": timed_steps, "repeats": reps, "timed_region_s": ms_total * 1e-3,
                   "parallelism": (f"query keyframes sharded over {world} GPU(s); obs_comm_allgather of the descriptor shards every step in {NC} "
                                   f"chunks ({local_bytes * world / 1e6:.0f} MB gathered per rank per step), remote pairs wait per chunk")
                                  if world > 1 else "1 GPU, no collective",
                   "l2": f"descriptor set {K * NDESC * 32 / 1e6:.0f} MB + {P_all * NDESC * 4 / 1e6:.0f} MB of results per step (> 126 MB L2)",
                   "queries_per_s": value / NDESC, "accepted_matches": matches},
        "result_hash": _hash_check(f"knn2:K={K}:W={Wn}", job_hash),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "distances/s", "h2d_bytes_per_step": world * local_bytes, "d2h_bytes_per_step": P_all * NDESC * 4,
                "steps": e2e_steps, "seconds": e2e_s,
                "api": "descriptor shard H2D from page-locked memory, obs_comm_allgather, obs_hamming_knn2, best indices D2H to page-locked memory"},
        "gpu_launches": (3 * (1 + NC) if tensor else (1 + NC)) * timed_steps,
        "roofline": roof,
        "cpu_baseline": cpu}


def run_reference(args, ctx):
    """`--impl reference` of the non-headline workloads: the CPU arm of that workload on all host cores, rank 0 only."""
    if ctx.rank != 0:
        return
    cores = os.cpu_count() or 1
    if args.workload == "knn2":
        D = synth.keyframe_descriptors(8, NDESC, 0)
        pairs = sharding.window_pairs(0, 8, 8, 8)[:max(cores * 2, 8)]
        total = 0.0
        for _ in range(args.warmup):
            cpu_knn2(D, pairs[:cores], cores)
        for _ in range(args.steps):
            total += cpu_knn2(D, pairs, cores)
        v = len(pairs) * NDESC * NDESC * args.steps / total
        line = {"metric": KNN2_METRIC, "unit": "distances/s", "scaling": "strong", "dtype": "u32 popcount",
                "config": {"workload": knn2_workload(args.keyframes, args.window), "pairs_per_step": len(pairs)},
                "kind": "port", "sample": f"{len(pairs)} keyframe pairs per step, SearchByBoW candidate loop with the reference's SWAR DescriptorDistance"}
    elif args.workload == "extract":
        imgs = [synth.blocky_image((376, 1241), i) for i in range(2 * cores)]
        for _ in range(args.warmup):
            cpu_extract(imgs[:cores], cores, 2000)
        total, kind = 0.0, "port"
        for _ in range(args.steps):
            dt, kind = cpu_extract(imgs, cores, 2000)
            total += dt
        v = len(imgs) * args.steps / total
        line = {"metric": EXTRACT_METRIC, "unit": "frames/s", "scaling": "strong", "dtype": "u8",
                "config": {"workload": extract_workload(args.total_frames), "frames_per_step": len(imgs)},
                "kind": kind, "sample": f"{len(imgs)} frames per step; reference ORBextractor.cc compiled in place, {cores} host threads"}
    else:
        return (run_projection if args.workload == "projection" else run_bow)(args, ctx.rank, ctx.local_rank, ctx.world, None)
    out = {"impl": "reference", "metric": line["metric"], "value": v, "unit": line["unit"], "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": line["scaling"],
           "vs_baseline": None, "dtype": line["dtype"], "data": "synthetic", "config": line["config"],
           "cpu_baseline": {"value": v, "unit": line["unit"], "cores": cores, "kind": line["kind"], "sample": line["sample"]},
           "e2e": {"value": v, "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def run(args, ctx, ClockSampler):
    """One of the non-headline workloads as its own JSON line (returned on rank 0; projection / bow print theirs)."""
    if args.workload == "extract":
        return measure_extract(args, ctx, ClockSampler)
    if args.workload == "knn2":
        return measure_knn2(args, ctx, ClockSampler)
    (run_bow if args.workload == "bow" else run_projection)(args, ctx.rank, ctx.local_rank, ctx.world, ClockSampler)
    return None
