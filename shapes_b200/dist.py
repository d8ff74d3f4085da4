"""Host-side logic of the multi-GPU path.

Ownership is by the LARGER key of a pair: a pair's rows end up on the rank that is HOME to its i.
Round-1 exchanges: one slot range per rank, the global descending (i, j) order is the slices concatenated from the
highest rank down.  Rows mode: the slot space is cut into 2 G blocks and rank g is home to blocks g and 2 G - 1 - g
(folded: equal slots and equal pairs per rank for any slot numbering); a rank's arrays hold the rows of its HIGH block
first (run 0), then those of its low block (run 1), and the global order is run 0 of ranks 0 .. G-1 followed by run 1
of ranks G-1 .. 0 -- still a fixed concatenation, no merge.  (The sweep / SAT work is partitioned by grid rows on the
device; that partition does not show in the result layout.)
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


def chunk_size(n_slots: int, world_size: int) -> int:
    """Slots per rank = the all-gather granule (ceil division, as the library does)."""
    return (n_slots + world_size - 1) // world_size


def own_range(n_slots: int, rank: int, world_size: int) -> tuple[int, int]:
    c = chunk_size(n_slots, world_size)
    return min(rank * c, n_slots), min((rank + 1) * c, n_slots)


def owner_of(i: np.ndarray, n_slots: int, world_size: int) -> np.ndarray:
    """Rank owning each pair, from the pair's larger key."""
    return np.asarray(i) // max(chunk_size(n_slots, world_size), 1)


def global_row_offsets(counts: Sequence[int]) -> list[int]:
    """Global row offset of every rank's slice: rank G-1 comes first."""
    offs = [0] * len(counts)
    run = 0
    for r in range(len(counts) - 1, -1, -1):
        offs[r] = run
        run += int(counts[r])
    return offs


def assemble_descending(slices: Sequence[dict]) -> dict:
    """Concatenate per-rank column dicts (index = rank) into the global descending order."""
    keys = slices[0].keys()
    return {k: np.concatenate([np.asarray(slices[r][k]) for r in range(len(slices) - 1, -1, -1)]) for k in keys}


def assemble_runs(slices: Sequence[dict], segments: Sequence[tuple], pair_cols=("pair_i", "pair_j")) -> dict:
    """Global descending order from per-rank column dicts and their shapes_rank_segments: run 0 of ranks 0, 1, ...,
    G-1, then run 1 of ranks G-1, ..., 0 (a rank's arrays hold run 0 first).  `segments[r]` =
    ((lo0, hi0, pairs0, contacts0), (lo1, hi1, pairs1, contacts1)); exchanges with one slot range per rank have an empty
    run 0, which makes this the plain "highest rank first" concatenation."""
    G = len(slices)
    out = {}
    for k in slices[0].keys():
        which = 2 if k in pair_cols else 3
        parts = []
        for r in range(G):
            parts.append(np.asarray(slices[r][k])[: segments[r][0][which]])
        for r in range(G - 1, -1, -1):
            n0 = segments[r][0][which]
            parts.append(np.asarray(slices[r][k])[n0: n0 + segments[r][1][which]])
        out[k] = np.concatenate(parts)
    return out


def home_blocks(n_slots: int, rank: int, world_size: int, fold: bool = True) -> tuple[tuple[int, int], tuple[int, int]]:
    """Rows mode: ((low block lo, hi), (high block lo, hi)) of a rank, as rows_home_blocks in the library computes them.
    fold=False: one contiguous block per rank (the high block is empty)."""
    c = chunk_size(n_slots, world_size)
    if not fold:
        return (min(rank * c, n_slots), min((rank + 1) * c, n_slots)), (n_slots, n_slots)
    blk = max((c + 1) // 2, 1)
    lo = (min(rank * blk, n_slots), min((rank + 1) * blk, n_slots))
    hi = (min((2 * world_size - 1 - rank) * blk, n_slots), min((2 * world_size - rank) * blk, n_slots))
    return lo, hi


def home_of(i: np.ndarray, n_slots: int, world_size: int, fold: bool = True) -> np.ndarray:
    """Rows mode: the home rank of every slot (= owner of the pairs whose larger key it is)."""
    c = chunk_size(n_slots, world_size)
    if not fold:
        return np.asarray(i) // max(c, 1)
    blk = max((c + 1) // 2, 1)
    b = np.asarray(i) // blk
    return np.where(b < world_size, b, 2 * world_size - 1 - b)
