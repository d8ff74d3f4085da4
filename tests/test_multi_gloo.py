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


def _worker_rows(rank, world_size, port, q, fold):
    """Rows-mode result layout: every rank holds the pairs whose larger key lies in its two home blocks, high block
    first; shapes_rank_segments-style segment records are exchanged and assemble_runs rebuilds the global order."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from oracle import binding as orc
    w = scenes.random_polygons(3001, density=2.0, static_frac=0.1, config=62)
    w.delete([11, 1501, 3000])
    c, s = orc.cos_sin(w.rot)
    full = orc.frame(w, c, s, broadphase="sweep")
    (l0, l1), (h0, h1) = sdist.home_blocks(w.n_slots, rank, world_size, fold)
    pi = full["pair_i"]
    in_hi, in_lo = (pi >= h0) & (pi < h1), (pi >= l0) & (pi < l1)
    assert np.array_equal(sdist.home_of(pi, w.n_slots, world_size, fold) == rank, in_hi | in_lo)
    order = np.concatenate([np.nonzero(in_hi)[0], np.nonzero(in_lo)[0]])       # run 0 = high block, run 1 = low block
    emin, emax = full["ext_min"], full["ext_max"]
    wx, wy, nx, ny = full["world_x"], full["world_y"], full["normal_wx"], full["normal_wy"]
    rows = orc.contacts(w, pi[order], full["pair_j"][order], wx, wy, nx, ny, emin, emax, 0.01, 0.01, 0.02)
    n_hi_rows = int(((rows["key_i"] >= h0) & (rows["key_i"] < h1)).sum())
    rows["pair_i"], rows["pair_j"] = pi[order], full["pair_j"][order]
    seg = ((h0, h1, int(in_hi.sum()), n_hi_rows), (l0, l1, int(in_lo.sum()), len(rows["key_i"]) - n_hi_rows))
    segs = [None] * world_size
    dist.all_gather_object(segs, seg)
    out = [None] * world_size if rank == 0 else None
    dist.gather_object(rows, out, dst=0)
    if rank == 0:
        glob = sdist.assemble_runs(out, segs)
        ok = all(np.array_equal(glob[k], full[k], equal_nan=True) for k in glob)
        per_rank = [sg[0][2] + sg[1][2] for sg in segs]
        q.put((ok, sum(per_rank), len(pi), max(per_rank) / (sum(per_rank) / world_size)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world_size,fold", [(2, True), (3, True), (2, False)])
def test_rows_mode_home_blocks_reassemble_global_order(world_size, fold):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_rows, args=(r, world_size, port, q, fold)) for r in range(world_size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    ok, total, want, imbalance = q.get(timeout=10)
    assert ok and total == want
    if fold:
        assert imbalance < 1.06        # folded blocks: equal pairs per home although the larger key of a pair skews high


def test_partition_helpers():
    (l0, l1), (h0, h1) = sdist.home_blocks(10, 0, 2)
    assert (l0, l1, h0, h1) == (0, 3, 9, 10) and sdist.home_blocks(10, 1, 2) == ((3, 6), (6, 9))
    assert sdist.home_of(np.arange(10), 10, 2).tolist() == [0, 0, 0, 1, 1, 1, 1, 1, 1, 0]
    assert sdist.home_blocks(10, 1, 2, fold=False) == ((5, 10), (10, 10))
    assert sdist.own_range(10, 0, 3) == (0, 4) and sdist.own_range(10, 2, 3) == (8, 10)
    assert sdist.own_range(2, 3, 4) == (2, 2)
    assert sdist.global_row_offsets([5, 7, 2]) == [9, 2, 0]
    assert sdist.owner_of(np.array([0, 3, 4, 9]), 10, 3).tolist() == [0, 0, 1, 2]
