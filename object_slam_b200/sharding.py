"""Host-side partitioning of the batched workloads over the GPUs of one box.

Frames (extraction, stereo, projection search) are independent: rank r takes a contiguous range and there
is no data-path collective.  Keyframe-vs-keyframe matching shards the *query* keyframes; every rank needs all
descriptor sets as database, which is the one exchange step (``Comm.allgather``, NCCL over NVLink).
Pure Python + numpy here; nothing in this module computes a match.
"""
import ctypes as C

import numpy as np


def shard_range(n, rank, world):
    """Contiguous range [lo, hi) of n items owned by `rank`; the first n % world ranks hold one more."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def padded_shard(n, world):
    """Items per rank when every rank must hold the same count (all-gather): ceil(n / world)."""
    return (n + world - 1) // world


def window_pairs(q_lo, q_hi, n_keyframes, window):
    """Ordered (query, database) keyframe pairs for queries [q_lo, q_hi): every other keyframe within
    `window` positions (window >= n_keyframes - 1 gives all ordered pairs)."""
    out = []
    for q in range(q_lo, min(q_hi, n_keyframes)):
        lo, hi = max(0, q - window), min(n_keyframes - 1, q + window)
        out.extend((q, d) for d in range(lo, hi + 1) if d != q)
    return np.asarray(out, np.int32).reshape(-1, 2)


def split_by_locality(pairs, lo, hi):
    """(pairs whose database keyframe lies in [lo, hi), the others): the first group needs no remote data,
    so it is matched while the all-gather is still in flight."""
    pairs = np.asarray(pairs, np.int32).reshape(-1, 2)
    local = (pairs[:, 1] >= lo) & (pairs[:, 1] < hi)
    return pairs[local], pairs[~local]


def split_by_chunk(pairs, lo, hi, per, n_chunks):
    """[pairs with a local database keyframe, pairs of gather chunk 0, chunk 1, ...]: the chunked all-gather
    (``Comm.allgather(..., n_chunks)``) delivers keyframes [per * c / n_chunks, per * (c + 1) / n_chunks) of EVERY rank's shard
    with chunk c, so a remote pair can run as soon as the chunk of its database keyframe has arrived.  Needs per % n_chunks == 0."""
    pairs = np.asarray(pairs, np.int32).reshape(-1, 2)
    assert n_chunks >= 1 and per % n_chunks == 0
    local = (pairs[:, 1] >= lo) & (pairs[:, 1] < hi)
    chunk = (pairs[:, 1] % per) * n_chunks // per
    return [pairs[local]] + [pairs[~local & (chunk == c)] for c in range(n_chunks)]


class Comm:
    """NCCL communicator of the library (csrc/comm.cu); the 128-byte id travels through `broadcast_id`,
    a callable that takes rank 0's bytes and returns them on every rank (e.g. via torch.distributed)."""

    def __init__(self, rank, world, device, broadcast_id):
        from ._capi import check, lib
        self._lib, self._check = lib(), check
        ident = (C.c_uint8 * 128)()
        if rank == 0:
            check(self._lib.obs_comm_unique_id(ident))
        raw = broadcast_id(bytes(ident))
        ident = (C.c_uint8 * 128).from_buffer_copy(raw)
        self._h = C.c_void_p()
        check(self._lib.obs_comm_create(ident, int(rank), int(world), int(device), C.byref(self._h)))
        self.rank, self.world = rank, world

    def allgather(self, d_local, local_bytes, d_all, n_chunks=1, producer_stream=None):
        self._check(self._lib.obs_comm_allgather(self._h, C.c_void_p(d_local), int(local_bytes), C.c_void_p(d_all),
                                                 int(n_chunks), C.c_void_p(producer_stream or 0)))

    def wait(self, chunk, consumer_stream):
        self._check(self._lib.obs_comm_wait(self._h, int(chunk), C.c_void_p(consumer_stream or 0)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.obs_comm_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close
