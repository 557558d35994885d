"""Times obs_comm_allgather alone (no matching kernel) under torchrun: in-place ncclAllGather (1 chunk) against the chunked form
(grouped ncclBroadcast per chunk), for the descriptor shards of configs[4] (4096 x 2000 x 32 B in total).
    python -m torch.distributed.run --nproc-per-node N tools/allgather_probe.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from object_slam_b200 import sharding  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
per = 4096 // world
local_bytes = per * 2000 * 32
allD = torch.zeros(world * local_bytes, dtype=torch.uint8, device=dev)
mine = allD[rank * local_bytes:(rank + 1) * local_bytes]
mine.fill_(rank + 1)


def bcast(raw):
    t = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())


comm = sharding.Comm(rank, world, local, bcast)
st = torch.cuda.Stream()
out = {}
for nc in (1, 2, 4, 8):
    for it in range(3):
        comm.allgather(mine.data_ptr(), local_bytes, allD.data_ptr(), nc, st.cuda_stream)
        comm.wait(nc - 1, st.cuda_stream)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record()
    reps = 10
    for it in range(reps):
        comm.allgather(mine.data_ptr(), local_bytes, allD.data_ptr(), nc, st.cuda_stream)
        comm.wait(nc - 1, st.cuda_stream)
    with torch.cuda.stream(st):
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[f"chunks_{nc}"] = {"ms": float(t), "recv_gbs_per_rank": (world - 1) * local_bytes / (float(t) * 1e-3) / 1e9}
ok = all(int(allD[r * local_bytes]) == r + 1 and int(allD[(r + 1) * local_bytes - 1]) == r + 1 for r in range(world))
if rank == 0:
    print(json.dumps({"world": world, "gathered_mb": world * local_bytes / 1e6, "correct": ok, **out}))
comm.close()
dist.destroy_process_group()
