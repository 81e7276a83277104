"""Multi-GPU sharding of par_cast: one process per GPU (torchrun).  Pixels are independent
(src/lib.rs:324-332 parallelises over scanlines) and the RNG is keyed by global pixel and sample index,
so the assembled image is bit-identical for every world size.

Two ways to share the frame out and put it together again:

  peer stores (default)  the frame's 8x4-pixel tiles (the megakernel's work tiles) are dealt round-robin — rank r renders
                         tiles r, r + G, ... of the whole frame, so every rank's share is spread evenly over the image —
                         and the kernel that finishes a pixel (the in-order sample fold) stores it straight into EVERY
                         rank's frame through NVLink peer pointers, at its final position (rtiow_b200_render_peers):
                         the exchange is fused into the fold, no collective is called on the data path, only the
                         ranks' 128-byte frame handles are exchanged once at set-up.
  NCCL                   BASELINE.json's wording, kept as the cross-check and fallback: the rows are partitioned
                         (RowShard), every rank renders its packed row block, ONE all_gather_into_tensor exchanges the
                         blocks.  Two partitions: interleaved (default) — bands of 4 scanlines, rank r renders bands
                         r, r + G, ..., de-interleaved by a single strided device copy after the gather — and
                         contiguous — rank r renders rows [r*ny/G, (r+1)*ny/G) straight into its slice of the frame,
                         the all-gather is in place; simpler, but on the book-1 scene the top ranks finish early.

With one rank nothing is exchanged at all.
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

    def describe(self, tiles=False):
        """`tiles`: the peer-store exchange deals 8x4-pixel tiles instead (rank r: tiles r, r + G, ... of the whole frame;
        rtiow_b200_render_peers) and the library assembles the frame."""
        if tiles:
            return f"{max(self.counts)} rows' worth x {self.world_size} rank(s), 8x4-pixel tiles dealt round-robin (rank r: tiles r, r+G, ...)"
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


class _DeviceArray:
    """A raw device pointer as something torch.as_tensor understands."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class PeerFrame:
    """This rank's peer-visible frame (rtiow_b200_peer_frame_*): created on `device`, handles exchanged over the default
    process group (any backend: 128 bytes per rank, once), peers mapped with CUDA IPC."""

    def __init__(self, lib, nx, ny, device_index, rank, world_size, connect=True):
        self.lib, self.nx, self.ny, self.device_index, self.world_size = lib, nx, ny, device_index, world_size
        self.h = C.c_void_p()
        api._check(lib.rtiow_b200_peer_frame_create(device_index, nx, ny, rank, world_size, C.byref(self.h)), lib)
        mine = (C.c_uint8 * N.PEER_HANDLE_BYTES)()
        api._check(lib.rtiow_b200_peer_frame_export(self.h, mine), lib)
        self.handle = bytes(mine)
        self.frame = None
        if connect:
            got = [self.handle]
            if world_size > 1:
                import torch.distributed as dist
                got = [None] * world_size
                dist.all_gather_object(got, self.handle)
            self.connect(got)

    def connect(self, handles):
        """`handles`: every rank's `.handle`, in rank order."""
        import torch
        blob = (C.c_uint8 * (N.PEER_HANDLE_BYTES * self.world_size)).from_buffer_copy(b"".join(handles))
        api._check(self.lib.rtiow_b200_peer_frame_connect(self.h, blob), self.lib)
        self._views = {}
        self._latest()

    def _latest(self):
        """`frame` = the buffer the most recent render assembled (the library alternates between two, one per epoch)."""
        import torch
        ptr = C.c_void_p()
        api._check(self.lib.rtiow_b200_peer_frame_ptr(self.h, C.byref(ptr)), self.lib)
        if ptr.value not in self._views:
            self._views[ptr.value] = torch.as_tensor(_DeviceArray(ptr.value, (self.ny, self.nx, 3)), device=f"cuda:{self.device_index}")
        self.frame = self._views[ptr.value]

    def render(self, nx, ny, ns, camera, world, seed=api.DEFAULT_SEED, stream=None):
        """This rank's share (rtiow_b200_render_peers), enqueued on `stream` (default: current)."""
        import torch
        s = stream if stream is not None else torch.cuda.current_stream(self.device_index)
        api._check(self.lib.rtiow_b200_render_peers(world.gpu(self.device_index), C.byref(camera.rec), nx, ny, ns, seed,
                                                         self.h, C.c_void_p(s.cuda_stream)), self.lib)
        self._latest()

    def close(self):
        if self.h:
            self.frame = None
            self._views = {}
            self.lib.rtiow_b200_peer_frame_destroy(self.h)
            self.h = None


class ShardBuffers:
    """Device buffers of one rank, allocated once.  Peer exchange: `frame` is this rank's peer-visible frame of the latest
    render (two buffers alternate), every rank's fold writes into it.  NCCL exchange: `mine` (this rank's packed rows, padded to max_rows), `parts` (everybody's),
    `frame` (the assembled image)."""

    def __init__(self, nx, ny, shard, device, world=None, exchange="auto"):
        import torch
        self.shard = shard
        self.peer = None
        dev_index = torch.device(device).index or 0
        want_peer = exchange in ("auto", "peer") and shard.world_size > 1 and shard.interleaved and world is not None
        if want_peer:
            try:
                self.peer = PeerFrame(world.lib, nx, ny, dev_index, shard.rank, shard.world_size)
            except Exception as e:      # noqa: BLE001  no IPC / no peer access on this box: every rank must fall back together
                if exchange == "peer":
                    raise
                self.peer, self.peer_error = None, str(e)
            import torch.distributed as dist
            ok = torch.tensor([1 if self.peer is not None else 0], device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok) == 0 and self.peer is not None:
                self.peer.close()
                self.peer = None
        if self.peer is not None:
            self._frame = None
            self.mine = self.parts = None
            self.exchange = ("fused into the sample fold: every finished pixel of this rank's tiles is stored into every rank's frame through NVLink peer "
                             "pointers (rtiow_b200_render_peers); no collective on the data path")
            return
        self._frame = torch.empty((ny, nx, 3), dtype=torch.float32, device=device)
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

    @property
    def frame(self):
        return self.peer.frame if self.peer is not None else self._frame

    def close(self):
        self._frame = self.mine = self.parts = None
        if self.peer is not None:
            self.peer.close()
            self.peer = None


def render_sharded_device(nx, ny, ns, camera, world, bufs, seed=api.DEFAULT_SEED):
    """Device-resident step: render this rank's rows, all-gather, assemble into bufs.frame.  Everything is
    enqueued on the current stream; nothing is synchronised."""
    import torch.distributed as dist
    sh = bufs.shard
    if bufs.peer is not None:
        bufs.peer.render(nx, ny, ns, camera, world, seed=seed)
        return bufs.frame
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
    import os
    import time
    import torch
    sh = bufs.shard
    dev = bufs.frame.device.index or 0
    trace = os.environ.get("RTIOW_E2E_TRACE")
    t0 = time.perf_counter()
    world.upload_fresh(dev)
    t1 = time.perf_counter()
    if sh.world_size == 1:
        api._check(world.lib.rtiow_b200_render(world.gpu(dev), C.byref(camera.rec), nx, ny, ns, seed, host_out.ctypes.data), world.lib)
        return world.scene_bytes(dev) + C.sizeof(N.CameraRec), host_out.nbytes
    render_sharded_device(nx, ny, ns, camera, world, bufs, seed=seed)
    t2 = time.perf_counter()
    d2h = 0
    if sh.rank == 0:
        host_out.copy_(bufs.frame, non_blocking=True)
        d2h = bufs.frame.numel() * 4
    torch.cuda.current_stream(dev).synchronize()
    t3 = time.perf_counter()
    h2d = world.scene_bytes(dev) + C.sizeof(N.CameraRec)     # after the sync: get_stats synchronises the device
    if trace:
        print(f"e2e rank {sh.rank}: upload {1e3 * (t1 - t0):.3f} ms, enqueue {1e3 * (t2 - t1):.3f} ms, finish+d2h {1e3 * (t3 - t2):.3f} ms, "
              f"stats {1e3 * (time.perf_counter() - t3):.3f} ms", flush=True)
    return h2d, d2h


def par_cast_distributed(nx, ny, ns, camera, world, seed=api.DEFAULT_SEED, interleaved=True):
    """par_cast across the ranks of the default process group (NCCL on GPUs).  Every rank returns the
    whole Image."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    ws = dist.get_world_size() if dist.is_initialized() else 1
    dev = torch.cuda.current_device()
    bufs = ShardBuffers(nx, ny, RowShard(ny, rank, ws, interleaved), f"cuda:{dev}", world=world)
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
