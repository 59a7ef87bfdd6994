"""Multi-GPU sharding of the hot path: the unit of independence is the FILE
(its blocks + its file-level filters).  Files are dealt to ranks greedily by block count;
the probe and the block-level build then need no exchange.  The only exchanges are
  * gathering each rank's candidate mask (bsg_allgather_masks / NCCL all-gather), and
  * OR-combining equal-(m,k) partial file-level bitsets built from disjoint shards of one
    file's entries (bsg_or_reduce).
This module is host logic only; the collective is injected so the same code runs over NCCL
(Context.allgather_masks) on GPUs and over gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np


class FileSharding:
    """Deterministic file -> rank assignment, identical on every rank."""

    def __init__(self, blocks_per_file: Sequence[int], world: int):
        self.world = world
        self.blocks_per_file = [int(b) for b in blocks_per_file]
        order = sorted(range(len(self.blocks_per_file)), key=lambda f: (-self.blocks_per_file[f], f))
        load = [0] * world
        self.owner = [0] * len(self.blocks_per_file)
        for f in order:  # longest-processing-time greedy: balances blocks (the probe's work unit)
            r = min(range(world), key=lambda i: (load[i], i))
            self.owner[f] = r
            load[r] += self.blocks_per_file[f]
        self.file_first_unit = np.concatenate([[0], np.cumsum(self.blocks_per_file)]).astype(np.int64)
        self.n_units = int(self.file_first_unit[-1])

    def files_of(self, rank: int) -> List[int]:
        return [f for f, r in enumerate(self.owner) if r == rank]

    def units_of(self, rank: int) -> np.ndarray:
        """Global unit (block) ids held by `rank`, in its local order."""
        parts = [np.arange(self.file_first_unit[f], self.file_first_unit[f + 1]) for f in self.files_of(rank)]
        return np.concatenate(parts).astype(np.int64) if parts else np.zeros(0, np.int64)

    def local_mask_words(self) -> int:
        """Every rank pads its mask to the same length so one all-gather moves them all."""
        most = max((len(self.units_of(r)) for r in range(self.world)), default=0)
        return (most + 63) // 64

    def pad_local_mask(self, mask_words: np.ndarray) -> np.ndarray:
        out = np.zeros(self.local_mask_words(), dtype=np.uint64)
        out[:len(mask_words)] = mask_words
        return out

    def assemble(self, gathered: np.ndarray) -> np.ndarray:
        """gathered[world, local_mask_words] -> bool[n_units] in global unit order."""
        out = np.zeros(self.n_units, dtype=bool)
        for r in range(self.world):
            units = self.units_of(r)
            bits = np.unpackbits(np.ascontiguousarray(gathered[r]).view(np.uint8), bitorder="little")[:len(units)]
            out[units] = bits.astype(bool)
        return out


def split_entries(n_entries: int, world: int, rank: int) -> slice:
    """Disjoint contiguous shard of one file's entry list for the partial file-level build."""
    lo = n_entries * rank // world
    hi = n_entries * (rank + 1) // world
    return slice(lo, hi)


def sharded_candidates(sharding: FileSharding, rank: int, local_mask_words: np.ndarray,
                       all_gather: Callable[[np.ndarray], np.ndarray]) -> np.ndarray:
    """Gather every rank's candidate mask and return the global candidate set (bool[n_units])."""
    gathered = all_gather(sharding.pad_local_mask(local_mask_words))
    return sharding.assemble(np.asarray(gathered, dtype=np.uint64).reshape(sharding.world, -1))
