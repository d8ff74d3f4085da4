"""N > 1 host logic on CPU: world_size-2 `gloo` run of the ownership / ordering scheme.

The compute on each rank is the oracle restricted to the rank's pairs (a stand-in for the GPU
kernels, which need a device); what is under test is shapes_b200.dist: slot-range ownership by
the larger key, the count exchange, and that rank-descending concatenation reproduces the
reference's global descending order exactly.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from shapes_b200 import scenes
from shapes_b200 import dist as sdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from oracle import binding as orc
    w = scenes.random_polygons(3000, density=2.0, static_frac=0.1, config=61)
    w.delete([7, 1500, 2999])
    c, s = orc.cos_sin(w.rot)
    full = orc.frame(w, c, s, broadphase="sweep")
    lo, hi = sdist.own_range(w.n_slots, rank, world_size)
    mine = (full["pair_i"] >= lo) & (full["pair_i"] < hi)
    assert np.array_equal(sdist.owner_of(full["pair_i"], w.n_slots, world_size)[mine], np.full(mine.sum(), rank))
    # this rank's slice: its pairs, and the contact rows generated from them
    emin, emax = full["ext_min"], full["ext_max"]
    wx, wy, nx, ny = full["world_x"], full["world_y"], full["normal_wx"], full["normal_wy"]
    rows = orc.contacts(w, full["pair_i"][mine], full["pair_j"][mine], wx, wy, nx, ny, emin, emax, 0.01, 0.01, 0.02)
    rows["pair_i"], rows["pair_j"] = full["pair_i"][mine], full["pair_j"][mine]
    # exchange #2: counts
    counts = torch.zeros(world_size, 2, dtype=torch.int64)
    mine_counts = torch.tensor([int(mine.sum()), len(rows["key_i"])], dtype=torch.int64)
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world_size)]
    dist.all_gather(gathered, mine_counts)
    counts = torch.stack(gathered)
    offs_pairs = sdist.global_row_offsets(counts[:, 0].tolist())
    offs_rows = sdist.global_row_offsets(counts[:, 1].tolist())
    # every rank can place its slice in the global arrays without seeing the others' data
    assert np.array_equal(full["pair_i"][offs_pairs[rank]:offs_pairs[rank] + int(mine.sum())], rows["pair_i"])
    assert np.array_equal(full["depth"][offs_rows[rank]:offs_rows[rank] + len(rows["key_i"])], rows["depth"])
    out = [None] * world_size if rank == 0 else None
    dist.gather_object(rows, out, dst=0)
    if rank == 0:
        glob = sdist.assemble_descending(out)
        ok = all(np.array_equal(glob[k], full[k], equal_nan=True) for k in glob)
        q.put((ok, int(counts[:, 0].sum()), len(full["pair_i"])))
    dist.destroy_process_group()


@pytest.mark.parametrize("world_size", [2, 3])
def test_slot_range_ownership_reassembles_global_order(world_size):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world_size, port, q)) for r in range(world_size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    ok, total, want = q.get(timeout=10)
    assert ok and total == want


def test_partition_helpers():
    assert sdist.own_range(10, 0, 3) == (0, 4) and sdist.own_range(10, 2, 3) == (8, 10)
    assert sdist.own_range(2, 3, 4) == (2, 2)
    assert sdist.global_row_offsets([5, 7, 2]) == [9, 2, 0]
    assert sdist.owner_of(np.array([0, 3, 4, 9]), 10, 3).tolist() == [0, 0, 1, 2]
