"""Index arithmetic of the sequence-sharded retrieval-head cache (scope-table row f1).

Host-side only, no device code: where a token position lives when every retrieval head's cache is split over ``world``
ranks block-cyclically (block ``b`` of ``block`` consecutive positions lives on rank ``b % world``), so that the split
is balanced at ANY context length and a rank's slice is ordered by position (the decode kernel needs no mask inside a
slice: every cached key is older than the query).  The device pieces that consume it are ``duo_attention_partial`` /
``duo_merge_partials`` (include/duo_b200.h); the reference has no counterpart (it shards by head only,
duo_attn/utils.py:132-227)."""
from __future__ import annotations

from typing import List

import torch


class SeqShardPlan:
    def __init__(self, world: int, block: int = 1024):
        if world < 1 or block < 1:
            raise ValueError("world and block must be positive")
        self.world, self.block = int(world), int(block)

    def owner(self, pos: int) -> int:
        return (pos // self.block) % self.world

    def local_index(self, pos: int) -> int:
        """Row of ``pos`` inside its owner's slice."""
        return (pos // (self.block * self.world)) * self.block + pos % self.block

    def local_len(self, rank: int, n_tokens: int) -> int:
        """How many of the positions ``[0, n_tokens)`` live on ``rank``."""
        full_rounds, rem = divmod(n_tokens, self.block * self.world)
        extra = min(max(rem - rank * self.block, 0), self.block)
        return full_rounds * self.block + extra

    def capacity(self, max_tokens: int) -> int:
        """Rows a rank must allocate for a cache of ``max_tokens`` positions."""
        return max(self.local_len(r, max_tokens) for r in range(self.world))

    def positions(self, rank: int, n_tokens: int) -> torch.Tensor:
        """Global positions of ``rank``'s slice, in slice order (strictly increasing)."""
        n = self.local_len(rank, n_tokens)
        i = torch.arange(n, dtype=torch.long)
        return (i // self.block) * (self.block * self.world) + rank * self.block + i % self.block

    def split(self, n_tokens: int) -> List[torch.Tensor]:
        return [self.positions(r, n_tokens) for r in range(self.world)]
