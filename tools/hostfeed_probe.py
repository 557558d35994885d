#!/usr/bin/env python
"""Raw host-feed probe: N concurrent processes, one per GPU, do nothing but cudaMemcpyAsync between page-locked host memory and
their GPU -- the transfers of one e2e step of bench.py (119.5 MB of images in, 33.6 MB of results out), no kernels -- for N = 1, 2, 4, 8
(as many as the box has).  Reports the aggregate GB/s per direction and the stereo frames/s that feed could carry, with the host
threads left where the OS puts them and pinned to the CPUs NVML reports as local to each GPU (page-locked buffers allocated after
pinning, so they land on that NUMA node).  This is the ceiling of the e2e number at each N.

    python tools/hostfeed_probe.py [--seconds 2] > profiles/hostfeed_rXX.json
"""
import argparse
import json
import multiprocessing as mp
import os
import time

H2D_BYTES = 2 * 128 * 376 * 1241        # one e2e step: 128 stereo frames in
D2H_BYTES = 33_555_968                  # ... and their keypoints, descriptors, uRight, depth out


def worker(idx, n, affinity, seconds, barrier, q):
    import torch
    torch.cuda.set_device(idx)
    note = "os default"
    if affinity:
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            allowed = sorted(os.sched_getaffinity(0))
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, (max(allowed) + 64) // 64)
            near = sorted({64 * i + b for i, wd in enumerate(mask) for b in range(64) if (int(wd) >> b) & 1} & set(allowed))
            if near:
                os.sched_setaffinity(0, near)
                note = f"cpus {near[0]}..{near[-1]} ({len(near)})"
        except Exception as ex:
            note = f"unavailable ({type(ex).__name__})"
    hin = torch.empty(H2D_BYTES, dtype=torch.uint8).pin_memory()
    hin.fill_(1)
    hout = torch.empty(D2H_BYTES, dtype=torch.uint8).pin_memory()
    din = torch.empty(H2D_BYTES, dtype=torch.uint8, device="cuda")
    dout = torch.zeros(D2H_BYTES, dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "d2h", "both"):
        for _ in range(3):
            din.copy_(hin, non_blocking=True); hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
        barrier.wait()
        t0 = time.perf_counter()
        steps = 0
        while time.perf_counter() - t0 < seconds:
            for _ in range(4):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s_in):
                        din.copy_(hin, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s_out):
                        hout.copy_(dout, non_blocking=True)
                steps += 1
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res[mode] = (steps, dt)
        barrier.wait()
    q.put((idx, note, res))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=2.0)
    args = ap.parse_args()
    import torch
    ngpu = torch.cuda.device_count()
    ctx = mp.get_context("spawn")
    out = {"gpus_on_box": ngpu, "host_cpus": len(os.sched_getaffinity(0)), "h2d_bytes_per_step": H2D_BYTES, "d2h_bytes_per_step": D2H_BYTES, "runs": []}
    for n in [k for k in (1, 2, 4, 8) if k <= ngpu]:
        for affinity in (False, True):
            barrier = ctx.Barrier(n)
            q = ctx.Queue()
            ps = [ctx.Process(target=worker, args=(i, n, affinity, args.seconds, barrier, q)) for i in range(n)]
            for p in ps:
                p.start()
            got = [q.get() for _ in range(n)]
            for p in ps:
                p.join()
            run = {"n": n, "numa_affinity": affinity, "placement": sorted(set(g[1] for g in got))}
            for mode in ("h2d", "d2h", "both"):
                steps_per_s = sum(g[2][mode][0] / g[2][mode][1] for g in got)
                if mode in ("h2d", "both"):
                    run[f"{mode}_h2d_gbs"] = steps_per_s * H2D_BYTES / 1e9
                if mode in ("d2h", "both"):
                    run[f"{mode}_d2h_gbs"] = steps_per_s * D2H_BYTES / 1e9
                if mode == "both":
                    run["stereo_frames_per_s_ceiling"] = steps_per_s * 128
            out["runs"].append(run)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
