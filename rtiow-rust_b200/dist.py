"""Multi-GPU sharding of par_cast: one process per GPU (torchrun).  Scanlines are independent
(src/lib.rs:324-332 parallelises exactly there) and the RNG is keyed by global pixel and sample index,
so the assembled image is bit-identical for every world size.

Two partitions of the rows:

  interleaved (default)  the frame is cut into bands of 4 scanlines (the height of the kernel's 8x4-pixel
                         work tiles) and rank r renders bands r, r + G, r + 2G, ...  Every GPU gets the same
                         mix of cheap (sky: one segment) and expensive (ground, spheres) scanlines.  The
                         ranks' packed row blocks are exchanged with ONE NCCL all-gather and de-interleaved
                         by a single strided device copy.
  contiguous             rank r renders rows [r*ny/G, (r+1)*ny/G) straight into its slice of the frame and
                         the all-gather is in place (BASELINE.json's "scanline block" wording); simpler, but
                         on the book-1 scene the top ranks finish early.

No other collective exists on the path; with one rank nothing is exchanged at all.
"""
import ctypes as C

import numpy as np

from . import _native as N
from . import api


BAND_ROWS = 4  # height of the megakernel's work tiles (render_kernel.cuh)


class RowShard:
    """Which output rows (row 0 = top) rank `rank` of `world_size` renders."""

    def __init__(self, ny, rank=0, world_size=1, interleaved=True, band=None):
        self.ny, self.rank, self.world_size, self.interleaved = ny, rank, world_size, bool(interleaved)
        if self.interleaved:
            # bands of `band` rows dealt round-robin; single rows when the frame is too small for that to balance
            self.band = band if band is not None else (BAND_ROWS if ny >= 8 * BAND_ROWS * world_size else 1)
            n_bands = (ny + self.band - 1) // self.band
            self.counts = [len(self._band_rows(r, n_bands)) for r in range(world_size)]
            self.begin, self.end, self.step = min(rank * self.band, ny), ny, world_size * self.band
            bands_per_rank = (n_bands + world_size - 1) // world_size
            self.max_rows = bands_per_rank * self.band          # every rank's block is whole band slots
        else:
            self.band = 1
            base, extra = divmod(ny, world_size)
            self.counts = [base + (1 if r < extra else 0) for r in range(world_size)]
            self.begins = [sum(self.counts[:r]) for r in range(world_size)]
            self.begin = self.begins[rank]
            self.end = self.begin + self.counts[rank]
            self.step = 1
            self.max_rows = max(self.counts)
        self.n_rows = self.counts[rank]
        self.uniform = len(set(self.counts)) == 1 and self.max_rows == self.counts[0]

    def _band_rows(self, r, n_bands):
        return [row for b in range(r, n_bands, self.world_size) for row in range(b * self.band, min((b + 1) * self.band, self.ny))]

    def rows(self, rank=None):
        """Global row indices of a rank, in the order they are packed."""
        r = self.rank if rank is None else rank
        if self.interleaved:
            return self._band_rows(r, (self.ny + self.band - 1) // self.band)
        b = sum(self.counts[:r])
        return list(range(b, b + self.counts[r]))

    def describe(self):
        kind = (f"interleaved bands of {self.band} row(s) (rank r: bands r, r+G, ...)" if self.interleaved else "contiguous blocks")
        return f"{max(self.counts)} rows x {self.world_size} rank(s), {kind}"


def assemble(parts, shard, xp=np):
    """[world_size, max_rows, nx, 3] packed blocks -> [ny, nx, 3] frame.  Works on numpy arrays and
    torch tensors (one strided copy)."""
    G, mx = shard.world_size, shard.max_rows
    if shard.interleaved:
        # band slot k of rank r is global band k * G + r: [G, slots, B, nx, 3] -> [slots, G, B, nx, 3]
        nx, B = parts.shape[2], shard.band
        if xp is np:
            return np.ascontiguousarray(parts.reshape(G, mx // B, B, nx, 3).transpose(1, 0, 2, 3, 4).reshape(mx * G, nx, 3)[:shard.ny])
        return parts.reshape(G, mx // B, B, nx, 3).permute(1, 0, 2, 3, 4).reshape(mx * G, nx, 3)[:shard.ny].contiguous()
    pieces = [parts[r, :c] for r, c in enumerate(shard.counts)]
    if xp is np:
        return np.concatenate(pieces)
    import torch
    return torch.cat(pieces)


class ShardBuffers:
    """Device buffers of one rank, allocated once: `mine` (this rank's packed rows, padded to max_rows),
    `parts` (everybody's), `frame` (the assembled image)."""

    def __init__(self, nx, ny, shard, device, world=None):
        import torch
        self.shard = shard
        self.frame = torch.empty((ny, nx, 3), dtype=torch.float32, device=device)
        self.exchange = "none (one rank)" if shard.world_size == 1 else "one NCCL all_gather_into_tensor of the packed rows" + (
            " + one strided de-interleave copy" if (shard.interleaved or not shard.uniform) else ", in place")
        if shard.world_size == 1:
            self.mine = self.frame
            self.parts = None
        elif shard.interleaved or not shard.uniform:
            self.parts = torch.zeros((shard.world_size, shard.max_rows, nx, 3), dtype=torch.float32, device=device)
            self.mine = self.parts[shard.rank]          # gather in place: send = recv + rank * count
        else:
            self.parts = None
            self.mine = self.frame[shard.begin:shard.end]


    def close(self):
        self.frame = self.mine = self.parts = None


def render_sharded_device(nx, ny, ns, camera, world, bufs, seed=api.DEFAULT_SEED):
    """Device-resident step: render this rank's rows, all-gather, assemble into bufs.frame.  Everything is
    enqueued on the current stream; nothing is synchronised."""
    import torch.distributed as dist
    sh = bufs.shard
    if sh.n_rows > 0:
        api.render_rows_device(nx, ny, ns, camera, world, bufs.mine, (sh.begin, sh.end), seed=seed, row_step=sh.step, row_band=sh.band)
    if sh.world_size > 1:
        if bufs.parts is not None:
            dist.all_gather_into_tensor(bufs.parts.view(sh.world_size * sh.max_rows, nx, 3), bufs.mine)
            bufs.frame.copy_(assemble(bufs.parts, sh, xp=None))
        else:
            dist.all_gather_into_tensor(bufs.frame, bufs.mine)   # contiguous, equal blocks: in place
    return bufs.frame


def par_cast_e2e(nx, ny, ns, camera, world, bufs, host_out, seed=api.DEFAULT_SEED):
    """End-to-end step with host buffers.  One rank: exactly the C ABI's host call — scene descriptor
    H2D (fresh rtiow_b200_scene_create), rtiow_b200_render into `host_out` (numpy, pageable).  Several ranks:
    fresh scene upload, sharded render, all-gather, frame D2H into `host_out` (pinned torch tensor) on rank 0.
    Returns (h2d_bytes, d2h_bytes) of this rank."""
    import torch
    sh = bufs.shard
    dev = bufs.frame.device.index or 0
    world.upload_fresh(dev)
    if sh.world_size == 1:
        api._check(world.lib.rtiow_b200_render(world.gpu(dev), C.byref(camera.rec), nx, ny, ns, seed, host_out.ctypes.data), world.lib)
        return world.scene_bytes(dev) + C.sizeof(N.CameraRec), host_out.nbytes
    render_sharded_device(nx, ny, ns, camera, world, bufs, seed=seed)
    h2d = world.scene_bytes(dev) + C.sizeof(N.CameraRec)
    d2h = 0
    if sh.rank == 0:
        host_out.copy_(bufs.frame, non_blocking=True)
        d2h = bufs.frame.numel() * 4
    torch.cuda.current_stream(dev).synchronize()
    return h2d, d2h


def par_cast_distributed(nx, ny, ns, camera, world, seed=api.DEFAULT_SEED, interleaved=True):
    """par_cast across the ranks of the default process group (NCCL on GPUs).  Every rank returns the
    whole Image."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    ws = dist.get_world_size() if dist.is_initialized() else 1
    dev = torch.cuda.current_device()
    bufs = ShardBuffers(nx, ny, RowShard(ny, rank, ws, interleaved), f"cuda:{dev}")
    render_sharded_device(nx, ny, ns, camera, world, bufs, seed=seed)
    return api.Image(bufs.frame.cpu().numpy())


def gather_rows_cpu(local_rows, shard, nx):
    """The same exchange + assembly over any backend (used by the gloo tests on CPU).  `local_rows` is this
    rank's packed [n_rows, nx, 3] float32 numpy array."""
    import torch
    import torch.distributed as dist
    mine = torch.zeros((shard.max_rows, nx, 3), dtype=torch.float32)
    mine[:local_rows.shape[0]] = torch.from_numpy(np.ascontiguousarray(local_rows, np.float32))
    parts = torch.empty((shard.world_size, shard.max_rows, nx, 3), dtype=torch.float32)
    dist.all_gather_into_tensor(parts.view(shard.world_size * shard.max_rows, nx, 3), mine)
    return assemble(parts.numpy(), shard)
