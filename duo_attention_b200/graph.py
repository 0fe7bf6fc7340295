"""CUDA-graph replay of the decode step (scope-table row f3).

A decode step of the patched model is ~32 x (3 GEMMs + 6 of our kernels [+ 2 all-reduces]): at short context and
under tensor parallelism it is bound by launch latency and by the Python driver, not by the GPU.  The kernels
read the cache occupancy from device memory (``duo_cache_state.device_state``) and RoPE positions come from a
device tensor, so ONE captured step can be replayed for every generated token:

    g = DuoDecodeGraph(model, cache)          # cache: DuoKVCache / DuoAttentionStaticKVCache after prefill
    logits = g.step(next_token_tensor)        # [B, 1] int64 on the GPU; returns [B, 1, vocab]

``cache.evict_last`` / ``clear`` keep working between steps (they refresh the device copy).
"""
from __future__ import annotations

import torch


class DuoDecodeGraph:
    def __init__(self, model, cache, warmup: int = 2):
        if cache.growable:
            raise ValueError("DuoDecodeGraph needs a pre-allocated cache (DuoAttentionStaticKVCache): a growable "
                             "cache may be re-allocated, which would leave stale pointers in the captured graph")
        self.model, self.cache = model, cache
        cache.graph_attached = True  # the captured launches hold raw buffer addresses: no re-allocation from now on
        dev = cache.device
        B = cache.batch_size
        cache.enable_device_state()
        self.ids = torch.zeros(B, 1, dtype=torch.long, device=dev)
        self.pos = torch.zeros(1, 1, dtype=torch.long, device=dev)
        self.graph = torch.cuda.CUDAGraph()
        snap = (list(cache.kv_seq_len_list), list(cache.total_list), list(cache.lo_list))
        ring = cache.snapshot_ring()  # warm-up steps commit a throw-away token into the ring: undone below

        def restore():
            cache.kv_seq_len_list[:], cache.total_list[:], cache.lo_list[:] = (list(x) for x in snap)
            cache.sync_device_state()
            self.pos.fill_(cache.kv_seq_len)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):  # lazy weight fusion, cuBLAS workspaces, cudaFuncSetAttribute ...
                restore()
                self._forward()
            restore()
            torch.cuda.synchronize(dev)
            with torch.cuda.graph(self.graph, stream=side):
                self.logits = self._forward()
            cache.restore_ring(ring)
        torch.cuda.current_stream(dev).wait_stream(side)
        restore()  # capture itself did not execute anything

    def _forward(self):
        out = self.model(input_ids=self.ids, position_ids=self.pos, past_key_values=self.cache, use_cache=True)
        self.pos.add_(1)
        return out.logits

    def step(self, token: torch.Tensor) -> torch.Tensor:
        """Run one decode step for ``token`` ([B,1] int64, device or pinned host)."""
        c = self.cache
        for l in range(c.num_layers):  # the capture-time overflow check does not re-run on replay: same error as eager
            if c.num_full_kv_head_list[l] > 0 and c._rows_needed(l, 1) > c.full_cap_list[l]:
                raise ValueError(f"Trying to put 1 KVs into a cache with max size {c.max_size}, "
                                 f"current size: {c.kv_seq_len_list[l]}.")
        self.ids.copy_(token, non_blocking=True)
        self.graph.replay()
        self.cache.advance_host(1)
        return self.logits

    def resync(self):
        """Call after evict_last()/clear(): positions restart from the cache length."""
        self.pos.fill_(self.cache.kv_seq_len)
