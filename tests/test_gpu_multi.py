"""Multi-GPU parity, one PROCESS per GPU (2 ranks): the slices, concatenated from the highest rank down, must be
bit-identical to the oracle's global result, for every exchange the library has -- "rows" (default with mapped peers:
work split by grid rows, results delivered to the slot-range homes), "p2p" (SHAPES_B200_NO_ROWS=1: the r1 key push /
box pull with slot-range ownership of the whole path) and "nccl" (no peer mapping: all-gather of the AABB records).
Skipped on boxes with fewer than 2 GPUs."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, nccl_id, q, kind, mode, port):
    p2p = mode != "nccl"
    if mode == "p2p":
        os.environ["SHAPES_B200_NO_ROWS"] = "1"
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from oracle import binding as orc
    from shapes_b200 import scenes
    from shapes_b200.engine import Engine
    w = scenes.box_pile(150, 120) if kind == "pile" else scenes.random_polygons(40_000, density=1.5, config=71)
    c, s = orc.cos_sin(w.rot)
    with Engine(w, device=rank, rank=rank, world_size=world_size, nccl_id=nccl_id) as eng:
        if p2p:   # AABB records stored straight into the peers' arrays instead of the NCCL all-gather
            blobs = [None] * world_size
            dist.all_gather_object(blobs, eng.ipc_export())
            eng.ipc_import(blobs)
        fr = eng.frame(cos_sin=(c, s))
        lo, hi, pairs0, contacts0 = eng.rank_info()
        assert pairs0[rank] == fr.n_pairs and contacts0[rank] == fr.n_contacts
        for _ in range(4):   # frames alternate exchange buffers (rows mode may re-home the slots after the first ones):
            fr3 = eng.frame(cos_sin=(c, s))      # the job's totals must not change
            lo, hi, pairs, contacts = eng.rank_info()
            assert pairs[rank] == fr3.n_pairs and contacts[rank] == fr3.n_contacts
            assert sum(pairs) == sum(pairs0) and sum(contacts) == sum(contacts0)
        q.put((rank, {k: np.array(v) for k, v in fr3.cols.items()}, eng.rank_segments(), pairs, contacts))
        dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["nccl", "p2p", "rows"])
@pytest.mark.parametrize("kind", ["pile", "polygons"])
def test_two_rank_slices_reassemble_to_the_oracle(oracle, kind, mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from conftest import assert_frames_match
    from shapes_b200 import dist as sdist, scenes
    from shapes_b200.engine import nccl_unique_id
    world_size = 2
    nccl_id = nccl_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world_size, nccl_id, q, kind, mode, port)) for r in range(world_size)]
    for p in procs:
        p.start()
    got, segs = {}, {}
    for _ in range(world_size):
        rank, cols, seg, pairs, contacts = q.get(timeout=300)
        got[rank] = cols
        segs[rank] = seg
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = scenes.box_pile(150, 120) if kind == "pile" else scenes.random_polygons(40_000, density=1.5, config=71)
    c, s = oracle.cos_sin(w.rot)
    want = oracle.frame(w, c, s, broadphase="sweep")
    glob = sdist.assemble_runs([got[r] for r in range(world_size)], [segs[r] for r in range(world_size)])
    assert_frames_match(glob, want)
    assert sum(pairs) == len(want["pair_i"]) and sum(contacts) == len(want["key_i"])
