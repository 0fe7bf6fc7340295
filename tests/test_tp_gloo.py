"""World-size-2 and -4 gloo (CPU) tests of the head-parallel host logic: planning, weight slicing and the one
all-reduce per layer.  The attention op itself is injected (oracle, CPU) — on the GPU box it is the CUDA path."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from duo_attention_b200 import tp
from oracle import duo_oracle as O

D = 128


def test_plan_is_a_balanced_partition():
    import json

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "patterns.json")))
    rows = gold["Llama-3-8B-Instruct-Gradient-1048k|0.5"]["mask_rows"]
    mask = np.array([[int(c) for c in r] for r in rows], dtype=float)
    for world in (1, 2, 4, 8):
        plan = tp.plan_heads(mask, world)
        tot = [0] * world
        for l in range(mask.shape[0]):
            seen = sorted(h for r in range(world) for h in plan.owners[l][r])
            assert seen == list(range(8))
            for r in range(world):
                hs = plan.owners[l][r]
                assert len(hs) == 8 // world
                flags = [mask[l][h] > 0.5 for h in hs]
                assert flags == sorted(flags, reverse=True), "retrieval heads must come first inside a rank"
                tot[r] += sum(flags)
            per = [plan.n_full(l, r) for r in range(world)]
            assert max(per) - min(per) <= 1, "retrieval heads of a layer are spread evenly"
        assert sum(tot) == 128 and max(tot) - min(tot) <= 1, tot
        lm = plan.local_mask(0)
        assert lm.shape == (32, 8 // world)
    with pytest.raises(ValueError):
        tp.plan_heads(mask, 3)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)  # same full weights / inputs on every rank
        Hq, Hkv, hid, S, sink, recent = 8, 4, 64, 12, 2, 3
        G = Hq // Hkv
        mask = np.array([[1.0, 0.0, 1.0, 0.0]])
        wq, wk = torch.randn(Hq * D, hid) * 0.2, torch.randn(Hkv * D, hid) * 0.2
        wv, wo = torch.randn(Hkv * D, hid) * 0.2, torch.randn(hid, Hq * D) * 0.05
        chunks = [torch.randn(1, S, hid), torch.randn(1, 1, hid), torch.randn(1, 5, hid)]
        plan = tp.plan_heads(mask, world)
        heads = plan.owners[0][rank]
        sq, sk, sv, so = tp.shard_attention_weights(wq, wk, wv, wo, heads, G, D)
        w_local = O.AttnWeights(sq, sk, sv, so, len(heads) * G, len(heads), plan.n_full(0, rank))
        # single-process reference: reference reorder (retrieval heads first) of the full weights
        gate = torch.tensor(mask[0], dtype=torch.float32)
        fq, _ = O.reorder_rows_or_cols(wq, None, gate, G * D, "out")
        fk, _ = O.reorder_rows_or_cols(wk, None, gate, D, "out")
        fv, _ = O.reorder_rows_or_cols(wv, None, gate, D, "out")
        fo, _ = O.reorder_rows_or_cols(wo, None, gate, G * D, "in")
        w_full = O.AttnWeights(fq, fk, fv, fo, Hq, Hkv, 2)
        past_l = past_f = None
        pos = 0
        ok = True
        for hs in chunks:
            n = hs.shape[1]
            cos, sin = O.hf_cos_sin(torch.arange(pos, pos + n)[None], D, 10000.0, torch.float32)
            part, past_l = O.tuple_forward(w_local, hs, cos, sin, past_l, sink, recent)
            tp.all_reduce_sum(part)  # the exchange step under test
            full, past_f = O.tuple_forward(w_full, hs, cos, sin, past_f, sink, recent)
            ok = ok and torch.allclose(part, full, rtol=1e-4, atol=1e-5)
            pos += n
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_layers_sum_to_the_single_process_result(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {r: True for r in range(world)}


def _reshard_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from duo_attention_b200.seqshard import SeqShardPlan

        torch.manual_seed(0)  # same "global" cache on every rank
        mask = np.array([[1.0, 0.0, 1.0, 1.0], [0.0, 0.0, 0.0, 0.0], [1.0, 1.0, 1.0, 1.0]])
        n_tok, cap, block, B, Dm = 77, 96, 8, 2, 4
        plan = tp.plan_heads(mask, world)
        sp = SeqShardPlan(world, block)
        ok = True
        for l in range(mask.shape[0]):
            glob = torch.randn(B, 4, cap, Dm)                      # [B, original head id, position, d]
            mine = [h for h in plan.owners[l][rank] if mask[l][h] > 0.5]
            src = glob[:, mine].contiguous() if mine else torch.zeros(B, 0, cap, Dm)
            full_ids = [h for h in range(4) if mask[l][h] > 0.5]
            dst = torch.zeros(B, len(full_ids), sp.capacity(cap) + 1, Dm)
            tp.reshard_heads_to_seq(src, plan.owners[l], mask[l], rank, world, n_tok, block, dst)
            pos = sp.positions(rank, n_tok)
            want = glob[:, full_ids][:, :, pos] if full_ids else dst[:, :, : len(pos)]
            ok = ok and torch.equal(dst[:, :, : len(pos)], want) and bool((dst[:, :, len(pos):] == 0).all())
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_reshard_from_head_parallel_to_sequence_sharded_layout(world):
    """tp.reshard_heads_to_seq (point-to-point, what DuoSeqShardKVCache.load_from_head_parallel runs per layer):
    every rank ends up with its block-cyclic position slice of EVERY retrieval head, in reordered head order."""
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000) + world
    mp.spawn(_reshard_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {r: True for r in range(world)}
