"""torchrun --nproc-per-node N tools/check_allgather.py : obs_comm_allgather (1 and 4 chunks) against the expected
layout, and sharded keyframe matching against the single-GPU result of the same pairs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from object_slam_b200 import sharding, synth  # noqa: E402
from object_slam_b200.matcher import ORBmatcher  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)


def bcast(raw):
    t = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())


comm = sharding.Comm(rank, world, lr, bcast)
K, n = 8 * world, 500
D = synth.keyframe_descriptors(K, n, 77)
per = K // world
ok = True
for chunks in (1, 4, 7):
    local = torch.from_numpy(D[rank * per:(rank + 1) * per]).to(dev)
    allD = torch.zeros((K, n, 32), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    comm.allgather(local.data_ptr(), local.numel(), allD.data_ptr(), chunks, torch.cuda.current_stream().cuda_stream)
    for c in range(chunks):
        comm.wait(c, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    good = np.array_equal(allD.cpu().numpy(), D)
    ok &= good
    print(f"rank {rank} chunks {chunks}: {'ok' if good else 'MISMATCH'}", flush=True)
M = ORBmatcher(0.6, True, device=lr)
pairs = sharding.window_pairs(rank * per, (rank + 1) * per, K, 3)
bi, bd, sd = M.knn2(allD.data_ptr(), pairs, n_keyframes=K, n_desc=n)
ref_bi, ref_bd, ref_sd = M.knn2(D, pairs)
good = np.array_equal(bi, ref_bi) and np.array_equal(bd, ref_bd) and np.array_equal(sd, ref_sd)
ok &= good
print(f"rank {rank} sharded knn2 on the gathered set: {'ok' if good else 'MISMATCH'} ({(bi >= 0).sum()} matches)", flush=True)
t = torch.tensor([int(ok)], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
comm.close()
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
