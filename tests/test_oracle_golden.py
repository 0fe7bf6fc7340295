"""Pin the CPU oracle against fixtures produced by the REFERENCE's own forward code
(tests/golden/make_golden.py).  fp32, CPU."""
import os

import numpy as np
import pytest
import torch

import golden_cases as GC
from oracle import duo_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def oracle_weights(case, data):
    gate = data["gate"]
    G = case["Hq"] // case["Hkv"]
    wq, _ = O.reorder_rows_or_cols(data["wq"], None, gate, G * GC.D, "out")
    wk, _ = O.reorder_rows_or_cols(data["wk"], None, gate, GC.D, "out")
    wv, _ = O.reorder_rows_or_cols(data["wv"], None, gate, GC.D, "out")
    wo, _ = O.reorder_rows_or_cols(data["wo"], None, gate, G * GC.D, "in")
    return O.AttnWeights(wq, wk, wv, wo, case["Hq"], case["Hkv"], int((gate > 0.5).sum()))


def run_oracle_layer(case, data):
    w = oracle_weights(case, data)
    outs = []
    pos = 0
    if case["path"] == "tuple":
        rot = GC.hf_rotary(case)
        past = None
        for hs in data["chunks"]:
            S = hs.shape[1]
            cos, sin = rot(hs, torch.arange(pos, pos + S)[None])
            out, past = O.tuple_forward(w, hs, cos, sin, past, case["sink"], case["recent"])
            outs.append(out)
            pos += S
        lens = (past[0].shape[2], past[1].shape[2])
    else:
        cache = O.OracleStaticKVCache(1, case["Hq"], case["Hkv"], GC.D, [data["gate"].numpy()], data["B"],
                                      case["max_size"], case["sink"], case["recent"])
        for i, hs in enumerate(data["chunks"]):
            S = hs.shape[1]
            out = O.static_forward(w, hs, torch.arange(pos, pos + S)[None], cache, 0, case["theta"],
                                   case.get("rope_factor") or 1.0)
            outs.append(out)
            pos += S
            ev = case.get("evict_after", {}).get(i, 0)
            if ev:
                cache.evict_last(ev)
                pos -= ev
        lens = (cache.kv_seq_len, cache.streaming_kv_seq_len)
    return torch.cat(outs, dim=1), lens


@pytest.mark.parametrize("case", GC.LAYER_CASES, ids=[c["name"] for c in GC.LAYER_CASES])
def test_oracle_matches_reference_forward(case):
    gold = np.load(os.path.join(GOLD, f"layer_{case['name']}.npz"))
    data = GC.make_layer_inputs(case)
    assert abs(GC.checksum(data) - float(gold["checksum"])) < 1e-6 * abs(float(gold["checksum"])), "RNG drift"
    with torch.no_grad():
        out, lens = run_oracle_layer(case, data)
    assert lens == (int(gold["final_full_len"]), int(gold["final_stream_len"]))
    np.testing.assert_allclose(out.numpy(), gold["out"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("seed", range(6))
def test_closed_form_mask_equals_cache_semantics(seed):
    """The chunk-relative streaming mask == cat/compaction semantics of the forward."""
    rng = np.random.RandomState(seed)
    sink, recent = int(rng.randint(0, 5)), int(rng.randint(1, 7))
    Hq, Hkv, n_full = 4, 2, int(rng.randint(0, 3))
    N = int(rng.randint(8, 40))
    cuts = sorted(set(rng.randint(1, N, size=rng.randint(1, 8)).tolist()))
    starts = [0] + cuts
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(1, N, Hq, 16, generator=g, dtype=torch.float64)
    k = torch.randn(1, N, Hkv, 16, generator=g, dtype=torch.float64)
    v = torch.randn(1, N, Hkv, 16, generator=g, dtype=torch.float64)
    dense = O.dense_duo_attention(q, k, v, starts, n_full, Hq // Hkv, sink, recent)
    past, outs = None, []
    for a, b in zip(starts, starts[1:] + [N]):
        o, past = O.tuple_attention_core(q[:, a:b], k[:, a:b], v[:, a:b], past, n_full, Hq // Hkv, sink, recent)
        outs.append(o)
    np.testing.assert_allclose(torch.cat(outs, 1).numpy(), dense.numpy(), rtol=2e-4, atol=2e-5)  # contract computes in fp32


def test_flash_contract_bottom_right_and_gqa():
    g = torch.Generator().manual_seed(0)
    q = torch.randn(2, 3, 4, 8, generator=g)
    k = torch.randn(2, 7, 2, 8, generator=g)
    v = torch.randn(2, 7, 2, 8, generator=g)
    out = O.flash_attn_contract(q, k, v, causal=True)
    for b in range(2):
        for h in range(4):
            for i in range(3):
                vis = i + 7 - 3 + 1
                s = (q[b, i, h] @ k[b, :vis, h // 2].T) / 8 ** 0.5
                ref = torch.softmax(s, -1) @ v[b, :vis, h // 2]
                torch.testing.assert_close(out[b, i, h], ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", GC.MISTRAL_CASES)
def test_oracle_matches_reference_mistral_twin(name):
    """The same oracle serves Mistral: the reference's mistral.py forwards produce bit-identical outputs to its
    llama.py forwards on these cases (recorded when the fixtures were generated), and the oracle matches both."""
    case = next(c for c in GC.LAYER_CASES if c["name"] == name)
    gold = np.load(os.path.join(GOLD, f"layer_mistral_{name}.npz"))
    assert bool(gold["identical_to_llama"])
    data = GC.make_layer_inputs(case)
    assert abs(GC.checksum(data) - float(gold["checksum"])) < 1e-6 * abs(float(gold["checksum"]))
    with torch.no_grad():
        out, _ = run_oracle_layer(case, data)
    np.testing.assert_allclose(out.numpy(), gold["out"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", GC.INT4_CASES, ids=[c["name"] for c in GC.INT4_CASES])
def test_oracle_int4_matches_reference_int4_forward(case):
    """int4_attention_core vs the fixture produced by RUNNING the reference's INT4 cache class (demo/int4_kv.py) and
    attention forward (demo/w8a8kv4_llama.py:174-287) on the same q/k/v (tests/golden/make_golden.py)."""
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", f"layer_{case['name']}.npz"))
    chunks = GC.make_int4_inputs(case)
    assert abs(GC.int4_checksum(chunks) - float(fx["checksum"])) < 1e-6 * float(fx["checksum"]), "input RNG drift"
    past, outs = None, []
    G = case["Hq"] // case["Hkv"]
    for i, (q, k, v) in enumerate(chunks):
        out, past = O.int4_attention_core(q, k, v, past, case["n_full"], G, case["sink"], case["recent"])
        outs.append(out)
        assert past[0].shape[2] == int(fx["lens"][i][0])                       # retrieval cache length
        if case["n_full"] < case["Hkv"]:
            assert past[1].shape[2] == int(fx["lens"][i][1])                   # compacted streaming cache length
    got = torch.cat(outs, dim=1).float().numpy()
    np.testing.assert_allclose(got, fx["out"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", GC.MODEL_CASES, ids=[c["name"] for c in GC.MODEL_CASES])
def test_oracle_model_matches_reference_patched_model(case):
    """OracleModel vs logits produced by RUNNING the reference's patched models (tuple driver and static driver +
    DuoAttentionStaticKVCache + flashinfer-rmsnorm patch, Llama and Mistral; tests/golden/make_golden.py).  The static
    driver rotates with fp32 on-the-fly angles instead of HF's cos/sin tables: same logits up to fp32 rounding."""
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", f"model_{case['name']}.npz"))
    model, ids, wsum = GC.make_tiny_model(case)
    assert abs(wsum - float(fx["checksum"])) < 1e-9 * float(fx["checksum"]), "weight RNG drift"
    om = O.OracleModel(model, np.array(case["gates"]), case["sink"], case["recent"])
    past, got = None, []
    for i, x in enumerate(ids):
        lo, past = om(x, past)
        got.append(lo)
        ev = case.get("evict_after", {}).get(i, 0)
        if ev:  # DuoAttentionStaticKVCache.evict_last (static_kv_cache.py:300-315) on the tuple layout
            past = tuple((f[:, :, :-ev], s[:, :, :-ev]) for f, s in past)
        if case["path"] == "static":
            assert past[-1][0].shape[2] == int(fx["lens"][i][0])
            n_stream = past[-1][1].shape[1]
            if n_stream:
                assert past[-1][1].shape[2] == int(fx["lens"][i][1])
    got = torch.cat(got, dim=1).numpy()
    tol = dict(rtol=1e-5, atol=1e-5) if case["path"] == "tuple" else dict(rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(got, fx["logits"], **tol)
    # the schedule really exercised the streaming window (the cache was compacted at least once)
    if case["path"] == "static" and past[0][1].shape[1]:
        assert int(fx["lens"][:, 1].max()) <= case["sink"] + case["recent"] < int(fx["lens"][:, 0].max())


# ---------------------------------------------------------------------------------------------------------
# Training-time streaming (Lambda) mask: reference generate_streaming_mask / streaming_attn_sdpa run on the CPU
# (duo_attn/patch/streaming_attn.py:14-42) -> tests/golden/training_masks.npz.  SURVEY 8c item 3.
# ---------------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("case", GC.TRAIN_MASK_CASES, ids=lambda c: c["name"])
def test_training_mask_closed_form_matches_reference_mask(case):
    f = np.load(os.path.join(GOLD, "training_masks.npz"))
    q, k, v = GC.make_train_mask_inputs(case)
    assert abs(float(f[f"checksum_{case['name']}"]) - sum(float(t.double().abs().sum()) for t in (q, k, v))) < 1e-6
    ref_mask = torch.from_numpy(f[f"mask_{case['name']}"])
    assert torch.equal(O.training_streaming_mask(case["S"], case["sink"], case["recent"]), ref_mask)
    out = O.training_streaming_attention(q, k, v, case["sink"], case["recent"])
    torch.testing.assert_close(out, torch.from_numpy(f[f"out_{case['name']}"]), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", [c for c in GC.TRAIN_MASK_CASES if c["recent"] >= 2], ids=lambda c: c["name"])
def test_deploy_decode_equals_training_mask_with_recent_plus_one(case):
    """The chunk = 1 limit of the deploy-time path (prefill of one token, then token-by-token decode through the tuple
    cache) is the training-time mask with ``recent + 1``: the decode step sees the ring (``recent`` keys) plus the new
    token, the training mask counts the query itself inside its window.  Held against the REFERENCE's mask + SDPA."""
    f = np.load(os.path.join(GOLD, "training_masks.npz"))
    q, k, v = GC.make_train_mask_inputs(case)
    groups = case["Hq"] // case["Hkv"]
    outs, past = [], None
    for t in range(case["S"]):
        o, past = O.tuple_attention_core(q[:, t:t + 1], k[:, t:t + 1], v[:, t:t + 1], past, 0, groups, case["sink"],
                                         case["recent"] - 1)
        outs.append(o.float())
    torch.testing.assert_close(torch.cat(outs, dim=1), torch.from_numpy(f[f"out_{case['name']}"]), rtol=1e-4, atol=1e-5)
    # and cell by cell: closed-form deploy visibility (chunk start = the token itself) == the reference's mask
    ref_mask = f[f"mask_{case['name']}"]
    for t in range(case["S"]):
        for j in range(case["S"]):
            assert O.streaming_visible(t, j, t, case["sink"], case["recent"] - 1) == bool(ref_mask[t, j]), (t, j)
