"""tp.shard_model against the sharding rule table of the reference (duo_attn/utils.py:132-227), extracted by RUNNING
get_mistral_config with recording stand-ins for the absent tensor_parallel package (tests/golden/make_golden.py ->
tests/golden/tp_rules.json)."""
import json
import os
import re

import numpy as np
import torch

from duo_attention_b200 import tp

HERE = os.path.dirname(os.path.abspath(__file__))
RULES = json.load(open(os.path.join(HERE, "golden", "tp_rules.json")))


def tiny():
    from transformers import MistralConfig, MistralForCausalLM

    torch.manual_seed(3)
    cfg = MistralConfig(hidden_size=1024, num_attention_heads=8, num_key_value_heads=4, head_dim=128,
                        num_hidden_layers=2, intermediate_size=96, vocab_size=40, max_position_embeddings=256,
                        sliding_window=None, attn_implementation="eager", tie_word_embeddings=False)
    return MistralForCausalLM(cfg).eval()


def rule_for(name):
    hits = [v for k, v in RULES["state_rules"].items() if re.match(k, name)]
    assert len(hits) <= 1
    return hits[0] if hits else None


def test_shards_follow_the_reference_rule_table():
    model = tiny()
    world = 2
    gates = (np.random.RandomState(0).rand(2, 4) > 0.5).astype(float)
    shards = [tp.shard_model(model, gates, r, world) for r in range(world)]
    full = dict(model.named_parameters())
    head_dim, group = 128, 2
    seen = set()
    for name, w in full.items():
        rule = rule_for(name)
        parts = [dict(s.named_parameters())[name] for s, _ in shards]
        if rule is None:  # norms: replicated
            assert all(torch.equal(p, w) for p in parts), name
            continue
        seen.add(name.split(".")[-2])
        dim = rule["dim"]
        if rule["kind"] == "SplitInChunks":
            # reference chunk = one KV head's rows (k, v) or its q-head group (q rows, o columns): chunk_size is
            # head_dim resp. q_per_kv * head_dim (here 128 / 256; 128 / 512 in the 32/8-head fixture)
            chunk = head_dim * (group if re.search(r"[qo]_proj", name) else 1)
            assert rule["chunk_size"] == 128 * (4 if re.search(r"[qo]_proj", name) else 1)
            chunks = list(torch.split(w, chunk, dim=dim))
            used = []
            for p in parts:
                assert p.shape[dim] * world == w.shape[dim]
                for piece in torch.split(p, chunk, dim=dim):
                    idx = [i for i, c in enumerate(chunks) if torch.equal(c, piece)]
                    assert len(idx) == 1, f"{name}: shard piece is not a whole reference chunk"
                    used.append(idx[0])
            assert sorted(used) == list(range(len(chunks))), f"{name}: the ranks' chunks do not partition the tensor"
        elif name.endswith(("embed_tokens.weight", "lm_head.weight")):
            # documented deviation: the reference splits the vocabulary matrices and gathers; they are replicated here
            assert all(torch.equal(p, w) for p in parts), name
        else:  # Split: rank r owns the r-th contiguous part
            for r, p in enumerate(parts):
                assert torch.equal(p, torch.chunk(w, world, dim=dim)[r]), name
    assert {"q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"} <= seen
    # q/k/v/o of a rank are the SAME kv heads (the rule table cuts all four with the same chunk index)
    for l in range(2):
        for (s, _), owned in zip(shards, tp.plan_heads(gates, world).owners[l]):
            a = s.model.layers[l].self_attn
            src = model.model.layers[l].self_attn
            for j, h in enumerate(owned):
                assert torch.equal(a.k_proj.weight[j * 128:(j + 1) * 128], src.k_proj.weight[h * 128:(h + 1) * 128])
                assert torch.equal(a.q_proj.weight[j * 256:(j + 1) * 256], src.q_proj.weight[h * 256:(h + 1) * 256])
                assert torch.equal(a.o_proj.weight[:, j * 256:(j + 1) * 256], src.o_proj.weight[:, h * 256:(h + 1) * 256])


def test_exchange_and_buffer_rules():
    # one "sum" after the attention block and one after the MLP: the two all-reduces of the driver
    assert RULES["output_rules"][".*self_attn$"]["0"] == "sum" and RULES["output_rules"][".*mlp$"]["0"] == "sum"
    # full_attention_heads is split on dim 0 (utils.py:219-221): every rank gets the mask entries of ITS heads
    assert RULES["buffer_rule"][".*full_attention_heads$"] == {"kind": "Split", "dim": 0}
    gates = (np.random.RandomState(1).rand(3, 8) > 0.5).astype(float)
    for world in (2, 4, 8):
        plan = tp.plan_heads(gates, world)
        for l in range(3):
            owned = sorted(h for r in range(world) for h in plan.owners[l][r])
            assert owned == list(range(8))
            for r in range(world):
                assert [gates[l][h] for h in plan.owners[l][r]] == list(plan.local_mask(r)[l])
