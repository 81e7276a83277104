"""world_size-2 test of the row-sharding / gather logic on CPU (gloo).  Each rank produces its row
block with the host-compiled copy of the kernel's per-path code (test infrastructure), the blocks
are all-gathered exactly as rtiow_rust_b200.dist does on NCCL, and the result must equal the
unsharded image bit for bit — the RNG is keyed by global pixel/sample, never by rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, ny, interleaved, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    import harness_lib as H
    import rtiow_rust_b200 as R
    from rtiow_rust_b200 import dist as rdist
    nx, ns = 40, 3
    world, cam = R.build_scene("kitchen_sink", nx, ny, use_bvh=True)
    shard = rdist.RowShard(ny, rank, world_size, interleaved)
    local, _ = H.render(world, cam, nx, ny, ns, rows=(shard.begin, shard.end), row_step=shard.step, row_band=shard.band)
    assert local.shape[0] == shard.n_rows
    full = rdist.gather_rows_cpu(local, shard, nx)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("interleaved", [True, False])
@pytest.mark.parametrize("ny", [30, 31, 70])   # 31: uneven split (16 + 15 rows); 70: bands of 4 rows, the last one partial
def test_two_rank_row_sharding_is_bit_identical(tmp_path, ny, interleaved):
    sys.path.insert(0, HERE)
    import harness_lib as H
    import rtiow_rust_b200 as R
    H.lib()  # build once, before forking
    mp.spawn(_worker, args=(2, _free_port(), ny, interleaved, str(tmp_path)), nprocs=2, join=True)
    world, cam = R.build_scene("kitchen_sink", 40, ny, use_bvh=True)
    want, _ = H.render(world, cam, 40, ny, 3)
    for r in range(2):
        got = np.load(tmp_path / f"rank{r}.npy")
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_row_shard_partition():
    from rtiow_rust_b200.dist import RowShard, assemble
    for ny, ws in ((800, 8), (800, 3), (5, 8), (1, 2), (3200, 8), (70, 2), (803, 8)):
        for inter in (True, False):
            shards = [RowShard(ny, r, ws, inter) for r in range(ws)]
            rows = sorted(sum((s.rows() for s in shards), []))
            assert rows == list(range(ny)), (ny, ws, inter)                       # every row exactly once
            assert all(len(s.rows()) == s.n_rows for s in shards)
            assert max(s.n_rows for s in shards) - min(s.n_rows for s in shards) <= shards[0].band
            for s in shards:                                                      # what the kernel is asked for
                want = [row for b in range(s.begin, s.end, s.step) for row in range(b, min(b + s.band, s.end))]
                assert want == s.rows()
            # assemble() puts packed row lr of rank r back at its global row
            parts = np.full((ws, shards[0].max_rows, 2, 3), -1, np.float32)
            for s in shards:
                for lr, g in enumerate(s.rows()):
                    parts[s.rank, lr] = g
            frame = assemble(parts, shards[0])
            assert frame.shape == (ny, 2, 3) and np.array_equal(frame[:, 0, 0], np.arange(ny, dtype=np.float32))
    # interleaving balances the book-1 frame: each rank's rows span the whole image
    s = RowShard(800, 3, 8)
    assert s.band == 4 and s.rows()[:5] == [12, 13, 14, 15, 44] and s.rows()[-1] == 783 and s.step == 32 and s.n_rows == 100
