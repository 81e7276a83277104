"""Multi-GPU sharding of par_cast: one process per GPU (torchrun), each rank renders a block of
scanlines straight into its slice of the full framebuffer on its own device, then ONE in-place
NCCL all-gather makes the frame whole on every rank (skipped when world_size == 1).

Scanlines are independent (src/lib.rs:324-332 parallelises exactly there) and the RNG is keyed by
global pixel and sample index, so the gathered image is bit-identical for every world size.
"""
import ctypes as C

import numpy as np

from . import _native as N
from . import api


class RowShard:
    """Rows [begin, end) of rank `rank` out of `world_size`: contiguous blocks, the first
    ny % world_size ranks get one extra row."""

    def __init__(self, ny, rank=0, world_size=1):
        self.ny, self.rank, self.world_size = ny, rank, world_size
        base, extra = divmod(ny, world_size)
        self.counts = [base + (1 if r < extra else 0) for r in range(world_size)]
        self.begins = [sum(self.counts[:r]) for r in range(world_size)]
        self.begin = self.begins[rank]
        self.end = self.begin + self.counts[rank]
        self.uniform = extra == 0

    def describe(self):
        return f"{self.counts[0]} rows x {self.world_size} rank(s), contiguous blocks"


def render_sharded_device(nx, ny, ns, camera, world, frame, shard, seed=api.DEFAULT_SEED):
    """Device-resident step: render this rank's rows into `frame` ([ny, nx, 3] CUDA float32), then
    all-gather.  Everything is enqueued on the current stream; nothing is synchronised."""
    import torch
    import torch.distributed as dist
    if shard.end > shard.begin:
        api.render_rows_device(nx, ny, ns, camera, world, frame[shard.begin:shard.end], (shard.begin, shard.end), seed=seed)
    if shard.world_size > 1:
        if shard.uniform:
            dist.all_gather_into_tensor(frame, frame[shard.begin:shard.end])  # in place: send = recv + rank*count
        else:  # ragged split: gather equal-sized padded blocks, then place each rank's rows
            mx = max(shard.counts)
            mine = torch.zeros((mx, nx, 3), dtype=frame.dtype, device=frame.device)
            mine[:shard.end - shard.begin] = frame[shard.begin:shard.end]
            allb = torch.empty((shard.world_size, mx, nx, 3), dtype=frame.dtype, device=frame.device)
            dist.all_gather_into_tensor(allb, mine)
            for r, (b, c) in enumerate(zip(shard.begins, shard.counts)):
                frame[b:b + c] = allb[r, :c]
    return frame


def par_cast_e2e(nx, ny, ns, camera, world, frame, pinned_out, shard, seed=api.DEFAULT_SEED):
    """End-to-end step with host buffers: scene descriptor H2D (fresh upload), render, gather, frame D2H
    into pinned host memory on rank 0.  Returns (h2d_bytes, d2h_bytes) of this rank."""
    import torch
    dev = frame.device.index or 0
    world.upload_fresh(dev)
    render_sharded_device(nx, ny, ns, camera, world, frame, shard, seed=seed)
    h2d = world.stats(dev)["scene_bytes"] + C.sizeof(N.CameraRec)
    d2h = 0
    if shard.rank == 0:
        pinned_out.copy_(frame, non_blocking=True)
        d2h = frame.numel() * 4
    torch.cuda.current_stream(dev).synchronize()
    return h2d, d2h


def par_cast_distributed(nx, ny, ns, camera, world, seed=api.DEFAULT_SEED):
    """par_cast across the ranks of the default process group (NCCL on GPUs).  Every rank returns the
    whole Image."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    ws = dist.get_world_size() if dist.is_initialized() else 1
    dev = torch.cuda.current_device()
    frame = torch.empty((ny, nx, 3), dtype=torch.float32, device=f"cuda:{dev}")
    render_sharded_device(nx, ny, ns, camera, world, frame, RowShard(ny, rank, ws), seed=seed)
    return api.Image(frame.cpu().numpy())


def gather_rows_cpu(local_rows, shard, nx):
    """The same assembly over any backend (used by the gloo tests on CPU): concatenates the ranks'
    row blocks in rank order.  `local_rows` is a [rows, nx, 3] float32 numpy array."""
    import torch
    import torch.distributed as dist
    mx = max(shard.counts)
    mine = torch.zeros((mx, nx, 3), dtype=torch.float32)
    mine[:local_rows.shape[0]] = torch.from_numpy(np.ascontiguousarray(local_rows, np.float32))
    parts = [torch.empty((mx, nx, 3), dtype=torch.float32) for _ in shard.counts]
    dist.all_gather(parts, mine)
    return torch.cat([p[:c] for p, c in zip(parts, shard.counts)]).numpy()
