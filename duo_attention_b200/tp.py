"""Head-parallel tensor parallelism for the DuoAttention hot path: one process per GPU, KV heads (with
their q-head groups) sharded across ranks, one all-reduce (sum) on the attention output per layer.

The reference shards the same way through the third-party ``tensor_parallel`` library inside ONE process
(duo_attn/utils.py:132-227: q/k/v split by KV-head chunks on dim 0, o_proj on dim 1, outputs "sum",
``full_attention_heads`` buffer split on dim 0).  Differences here, on purpose:

* one process per GPU + ``torch.distributed`` (NCCL over NVLink on the box, gloo in the CPU tests);
* heads are dealt to ranks BEFORE the retrieval-first reorder so every rank gets a balanced mix — the
  reference reorders first and then cuts contiguous chunks, which hands rank 0 most retrieval heads.

Host-side only: planning, weight slicing and the collective.  The attention itself is the CUDA path.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist


class HeadPlan:
    """Which original KV heads each rank owns in each layer (retrieval heads first inside a rank)."""

    def __init__(self, owners: List[List[List[int]]], mask: np.ndarray, world: int):
        self.owners = owners  # [layer][rank] -> list of original kv head ids, retrieval heads first
        self.mask = mask
        self.world = world

    def local_mask(self, rank: int) -> np.ndarray:
        """``[layers, kv_heads / world]`` binary mask of the rank's heads in its local order."""
        return np.array([[float(self.mask[l][h] > 0.5) for h in self.owners[l][rank]]
                         for l in range(len(self.owners))])

    def n_full(self, layer: int, rank: int) -> int:
        return int(sum(self.mask[layer][h] > 0.5 for h in self.owners[layer][rank]))


def plan_heads(mask: Sequence[Sequence[float]], world: int) -> HeadPlan:
    """Deal each layer's retrieval heads round-robin over the ranks (rotating the starting rank from layer
    to layer so the per-rank totals even out), then top every rank up to ``kv_heads / world`` with streaming
    heads."""
    mask = np.asarray(mask, dtype=np.float64)
    L, H = mask.shape
    if H % world != 0:
        raise ValueError(f"{H} KV heads cannot be split over {world} ranks")
    quota = H // world
    owners = []
    load = [0] * world  # retrieval heads handed out so far (across layers)
    for l in range(L):
        full = [h for h in range(H) if mask[l][h] > 0.5]
        stream = [h for h in range(H) if mask[l][h] <= 0.5]
        per_rank: List[List[int]] = [[] for _ in range(world)]
        for h in full:
            # least-loaded rank that still has room in this layer
            cand = [r for r in range(world) if len(per_rank[r]) < quota]
            r = min(cand, key=lambda r_: (load[r_], len(per_rank[r_]), r_))
            per_rank[r].append(h)
            load[r] += 1
        it = iter(stream)
        for r in range(world):
            while len(per_rank[r]) < quota:
                per_rank[r].append(next(it))
        owners.append(per_rank)
    return HeadPlan(owners, mask, world)


@torch.no_grad()
def shard_attention_weights(wq, wk, wv, wo, heads: Sequence[int], group: int, head_dim: int):
    """Rows of q/k/v and columns of o that belong to the given original KV heads, in that order."""
    qrows = torch.cat([torch.arange(h * group * head_dim, (h + 1) * group * head_dim) for h in heads])
    krows = torch.cat([torch.arange(h * head_dim, (h + 1) * head_dim) for h in heads])
    return wq[qrows], wk[krows], wv[krows], wo[:, qrows]


def all_reduce_sum(x: torch.Tensor, group=None) -> torch.Tensor:
    """The per-layer exchange step: sum of the row-parallel o_proj partials (duo_attn/utils.py:174-176)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
    return x


def install_allreduce(model, group=None):
    """Mark a patched (enable_duo_attention_eval) per-rank model shard as tensor-parallel: the driver then
    all-reduces the attention and MLP outputs of every layer."""
    model._duo_tp_group = group
    model._duo_tp = True
    return model
