#!/usr/bin/env python
"""Under torchrun on N GPUs: the sharded frame every rank ends up with must equal the single-GPU frame bit for bit —
peer-store exchange (rows folded straight into every rank's frame over NVLink) and both NCCL partitions.  Rank 0 also
checks rtiow_b200_render_multi (one host thread, all GPUs).  Prints one line per check on rank 0; exits non-zero on a mismatch."""
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
    for exchange, inter in (("peer", True), ("nccl", True), ("nccl", False)):
        shard = rdist.RowShard(ny, rank, ws, inter)
        bufs = rdist.ShardBuffers(nx, ny, shard, f"cuda:{local_rank}", world=world, exchange=exchange)
        oks = []
        for rep in range(3):       # repeated calls: the peer hand-shake must also protect the frames between steps
            seed = 0xDEADBEEF + rep
            w = want if rep == 0 else R.par_cast(nx, ny, ns, cam, world, device=local_rank, seed=seed).rgb
            rdist.render_sharded_device(nx, ny, ns, cam, world, bufs, seed=seed)
            torch.cuda.synchronize()
            got = bufs.frame.cpu().numpy()
            oks.append(np.array_equal(got.view(np.uint32), w.view(np.uint32)))
        flags = torch.tensor([0 if all(oks) else 1], device="cuda")
        dist.all_reduce(flags)
        if rank == 0:
            print(f"dist_check {name} {nx}x{ny}x{ns} world={ws} exchange={'peer stores' if bufs.peer is not None else 'nccl'} "
                  f"interleaved={inter}: mismatching ranks = {int(flags)}", flush=True)
        bad += int(flags)
        bufs.close()
    world.close()
dist.barrier()
if rank == 0:   # the single-process entry point, over all GPUs of the box
    for name, nx, ny, ns, bvh in (("book1", 300, 203, 6, True), ("final", 96, 64, 4, False)):
        worlds = [R.build_scene(name, nx, ny, use_bvh=bvh) for _ in range(ws)]
        want = R.par_cast(nx, ny, ns, worlds[0][1], worlds[0][0]).rgb
        got = [R.par_cast_multi(nx, ny, ns, worlds[0][1], [w for w, _ in worlds]).rgb for _ in range(2)]
        ok = all(np.array_equal(g.view(np.uint32), want.view(np.uint32)) for g in got)
        print(f"dist_check render_multi {name} {nx}x{ny}x{ns} ngpus={ws}: {'ok' if ok else 'MISMATCH'}", flush=True)
        bad += 0 if ok else 1
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if bad else 0)
