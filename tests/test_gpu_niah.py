"""Needle-in-a-Haystack generation protocol (eval/needle/niah_protocol.py, the reference's
eval/needle/needle_in_haystack.py:262-314) on a tiny random-init model through the drop-in API, tuple/dynamic cache as the
harness uses it: chunked pre-fill -> token-by-token "simulation" -> greedy decoding, against the CPU oracle model fed
the same tokens (teacher forced), with the generated tokens required to agree wherever the oracle's top-2 margin is clear."""
import copy
import importlib.util
import os

import numpy as np
import pytest
import torch

from duo_attn.patch import enable_duo_attention_eval
from oracle import duo_oracle as O
from test_gpu_model import tiny_model

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _protocol():
    spec = importlib.util.spec_from_file_location("niah_protocol", os.path.join(ROOT, "eval", "needle", "niah_protocol.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("kind,chunk", [("llama", 128), ("mistral", None)])
def test_niah_protocol_on_a_tiny_random_model(kind, chunk):
    model = tiny_model(kind, seed=21)
    gates = np.array([[1.0, 0.0], [0.0, 1.0]])
    sink, recent = 8, 32
    oracle = O.OracleModel(copy.deepcopy(model), gates, sink, recent)
    enable_duo_attention_eval(model, gates, sink, recent)
    model.cuda()
    g = torch.Generator().manual_seed(9)
    prompt = torch.randint(0, 512, (1, 700), generator=g)
    sim, new = 20, 12
    gen, past, logits = _protocol().niah_generate(model, prompt.cuda(), simulation_length=sim, prefilling_chunk_size=chunk,
                                                  max_new_tokens=new, collect_logits=True)
    assert len(gen) == new + 1 and past.kv_seq_len == 700 + new
    # replay the very same calls through the oracle (teacher forced with the product's tokens)
    calls = []
    ctx = prompt[:, : 700 - sim]
    step = chunk or ctx.size(1)
    calls += [ctx[:, i : i + step] for i in range(0, ctx.size(1), step)]
    calls += [prompt[:, i : i + 1] for i in range(700 - sim, 700)]
    calls += [torch.tensor([[t]]) for t in gen[:-1]]
    assert len(calls) == len(logits)
    past_o = None
    first_gen = len(calls) - new - 1
    for i, ids in enumerate(calls):
        lo, past_o = oracle(ids, past_o)
        torch.testing.assert_close(logits[i], lo[:, 0], rtol=5e-2, atol=5e-2)
        if i >= first_gen:  # this call's argmax is generated token i - first_gen
            top2 = lo[0, 0].topk(2).values
            if (top2[0] - top2[1]).item() > 0.15:
                assert gen[i - first_gen] == int(lo[0, 0].argmax()), f"generated token {i - first_gen}"
