#!/usr/bin/env python
"""Under torchrun on N GPUs: the sharded + gathered frame must equal the single-GPU frame bit for bit,
for both row partitions.  Prints one line per check on rank 0; exits non-zero on a mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtiow_rust_b200 as R  # noqa: E402
from rtiow_rust_b200 import dist as rdist  # noqa: E402

rank, local_rank, ws = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
bad = 0
for name, nx, ny, ns, bvh in (("book1", 300, 203, 6, True), ("final", 96, 64, 4, False), ("cornell", 64, 64, 4, False)):
    world, cam = R.build_scene(name, nx, ny, use_bvh=bvh)
    want = R.par_cast(nx, ny, ns, cam, world, device=local_rank).rgb
    for inter in (True, False):
        got = rdist.par_cast_distributed(nx, ny, ns, cam, world, interleaved=inter).rgb
        ok = np.array_equal(got.view(np.uint32), want.view(np.uint32))
        flags = torch.tensor([0 if ok else 1], device="cuda")
        dist.all_reduce(flags)
        if rank == 0:
            print(f"dist_check {name} {nx}x{ny}x{ns} world={ws} interleaved={inter}: mismatching ranks = {int(flags)}", flush=True)
        bad += int(flags)
    world.close()
dist.destroy_process_group()
sys.exit(1 if bad else 0)
