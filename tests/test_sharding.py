"""Host-side multi-GPU logic on CPU: shard arithmetic, and a world_size-2 gloo run of the keyframe-matching
schedule (shard queries, all-gather descriptor shards, match local queries) with the oracle standing in for the
kernels -- the union of the ranks' results must equal the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest

from object_slam_b200 import sharding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 4096, 32768):
        for world in (1, 2, 3, 4, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
            assert sharding.padded_shard(n, world) == max(sizes) or n == 0


def test_window_pairs_cover_all_pairs_once():
    K = 13
    allp = sharding.window_pairs(0, K, K, K)
    assert len(allp) == K * (K - 1) and len({tuple(p) for p in allp}) == len(allp)
    for world in (2, 4):
        parts = [sharding.window_pairs(*sharding.shard_range(K, r, world), K, 3) for r in range(world)]
        got = np.concatenate(parts)
        want = sharding.window_pairs(0, K, K, 3)
        assert np.array_equal(got, want)
        for r, p in enumerate(parts):
            lo, hi = sharding.shard_range(K, r, world)
            loc, rem = sharding.split_by_locality(p, lo, hi)
            assert len(loc) + len(rem) == len(p)
            assert np.all((loc[:, 1] >= lo) & (loc[:, 1] < hi)) and np.all((rem[:, 1] < lo) | (rem[:, 1] >= hi))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K, n, window = 8, 120, 2
    D = synth.keyframe_descriptors(K, n, 9)
    per = sharding.padded_shard(K, world)
    lo, hi = rank * per, min(K, (rank + 1) * per)
    local = torch.zeros((per, n, 32), dtype=torch.uint8)
    local[:hi - lo] = torch.from_numpy(D[lo:hi])
    gathered = torch.zeros((world * per, n, 32), dtype=torch.uint8)
    dist.all_gather_into_tensor(gathered, local)           # the exchange step (NCCL on the GPU box)
    allD = gathered.numpy()
    pairs = sharding.window_pairs(lo, hi, K, window)
    res = {}
    for a, b in pairs:
        bi, bd, sd = oracle.hamming_knn2(allD[a], allD[b], 50, 0.6)
        res[(int(a), int(b))] = (bi.copy(), bd.copy(), sd.copy())
    q.put((rank, np.array_equal(allD[:K], D), res))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_keyframe_matching_equals_single_process():
    import multiprocessing as mp
    import oracle
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    K, n, window = 8, 120, 2
    D = synth.keyframe_descriptors(K, n, 9)
    merged = {}
    for rank, gathered_ok, res in outs:
        assert gathered_ok
        assert not (set(res) & set(merged))
        merged.update(res)
    want = sharding.window_pairs(0, K, K, window)
    assert set(merged) == {tuple(int(v) for v in p) for p in want}
    for a, b in want:
        bi, bd, sd = oracle.hamming_knn2(D[a], D[b], 50, 0.6)
        g = merged[(int(a), int(b))]
        assert np.array_equal(g[0], bi) and np.array_equal(g[1], bd) and np.array_equal(g[2], sd)


def _frame_worker(rank, world, port, q):
    """The frame-sharded schedule of bench.py (stereo / extract workloads): contiguous shard per rank, no data-path
    collective, the step time is the max over the ranks (all_reduce MAX) -- with the oracle standing in for the kernels."""
    import time
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T, shape = 7, (120, 160)
    lo, hi = sharding.shard_range(T, rank, world)
    ex = oracle.OracleExtractor(200)
    dist.barrier()
    t0 = time.perf_counter()
    res = {i: tuple(a.copy() for a in ex(synth.blocky_image(shape, i))) for i in range(lo, hi)}
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
    mine = float(t.item())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, res, mine, float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_frame_sharding_equals_single_process():
    import multiprocessing as mp
    import oracle
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_frame_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    merged = {}
    for rank, res, mine, mx in outs:
        assert not (set(res) & set(merged))
        merged.update(res)
        assert mx >= mine
    assert len({mx for _, _, _, mx in outs}) == 1          # every rank reports the same (max) time
    assert sorted(merged) == list(range(7))
    ex = oracle.OracleExtractor(200)
    for i in range(7):
        k, d = ex(synth.blocky_image((120, 160), i))
        assert merged[i][0].tobytes() == k.tobytes() and np.array_equal(merged[i][1], d)


def test_split_by_chunk_partitions_the_pairs():
    K, world, W, nc = 64, 4, 9, 4
    per = sharding.padded_shard(K, world)
    for rank in range(world):
        lo, hi = rank * per, (rank + 1) * per
        pairs = sharding.window_pairs(lo, hi, K, W)
        groups = sharding.split_by_chunk(pairs, lo, hi, per, nc)
        assert len(groups) == 1 + nc
        assert sum(len(g) for g in groups) == len(pairs)
        assert sorted(map(tuple, np.concatenate(groups))) == sorted(map(tuple, pairs))
        assert all(lo <= d < hi for _, d in groups[0])
        for c in range(nc):
            for _, d in groups[1 + c]:
                assert not (lo <= d < hi) and (d % per) * nc // per == c
    # one rank: everything is local
    pairs = sharding.window_pairs(0, K, K, W)
    groups = sharding.split_by_chunk(pairs, 0, K, K, 1)
    assert len(groups[0]) == len(pairs) and len(groups[1]) == 0
