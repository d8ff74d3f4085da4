"""Host-side logic of the multi-GPU path.

Ownership is by the LARGER key of a pair: rank r owns the pairs / contacts whose i lies in its
slot range.  The reference order is descending (i, j), so the global result is simply the slices
concatenated from the highest rank down -- no merge.
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
