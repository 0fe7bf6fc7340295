"""Hardware checks of the EXPERIMENTAL, opt-in code paths.  Skipped unless DUO_EXPERIMENTAL=1, and each path is only
taken when its own switch is set in the environment of the test process:

    DUO_EXPERIMENTAL=1 DUO_INT4_SWAPAB=1 DUO_WIDE_MERGE=1 python -m pytest tests/test_gpu_experimental.py -x -q

(plus the regular suites under the same switches: tests/test_gpu_int4_attention.py, tests/test_gpu_attention.py,
tests/test_gpu_model.py).  See profiles/validate_experimental.sh."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("DUO_EXPERIMENTAL") != "1", reason="set DUO_EXPERIMENTAL=1")]


def test_int4_many_splits():
    """> 16 splits per retrieval head: the split-KV merge loop takes more than one pass per warp."""
    from test_gpu_int4_attention import run

    run(4, 1, 1, 64, 256, [20000, 1, 2, 1], seed=16, stage_cap=20000)


def test_int4_batch2_decode():
    from test_gpu_int4_attention import run

    run(8, 2, 1, 8, 24, [3000, 1, 2, 1, 1], seed=17, B=2, stage_cap=3000)


def test_int4_mha_rows_up_to_8():
    """group 1: q_len 1..8 all fit the 8-row decode kernel."""
    from test_gpu_int4_attention import run

    run(4, 4, 2, 4, 12, [50, 1, 8, 7, 5, 1, 3], seed=18, stage_cap=50)


def test_int4_cuda_graph_decode_matches_eager():
    """Device-resident occupancy (dstate) path of the INT4 decode kernels: graph replay == eager, token by token."""
    from duo_attention_b200.graph import DuoDecodeGraph
    from duo_attn.patch import DuoAttentionStaticKVCache, enable_llama_duo_attention_static_kv_cache_eval
    from test_gpu_model import tiny_model

    model = tiny_model("llama", seed=7).to(torch.float16)
    gates = np.array([[1.0, 0.0], [0.0, 1.0]])
    sink, recent = 4, 6
    enable_llama_duo_attention_static_kv_cache_eval(model, gates)
    model.cuda()
    ca = DuoAttentionStaticKVCache(model, gates, 1, 256, sink, recent, kv_format="int4")
    cb = DuoAttentionStaticKVCache(model, gates, 1, 256, sink, recent, kv_format="int4")
    g = torch.Generator().manual_seed(4)
    ids = torch.randint(0, 512, (1, 37), generator=g).cuda()
    with torch.no_grad():
        model(input_ids=ids, past_key_values=ca, use_cache=True)
        model(input_ids=ids, past_key_values=cb, use_cache=True)
        graph = DuoDecodeGraph(model, cb)
        toks = torch.randint(0, 512, (20, 1, 1), generator=g).cuda()
        for i in range(20):
            want = model(input_ids=toks[i], past_key_values=ca, use_cache=True).logits
            got = graph.step(toks[i])
            torch.testing.assert_close(got.float(), want.float(), rtol=0, atol=0, msg=lambda m: f"step {i}: {m}")
            if i == 9:
                ca.evict_last(1)
                cb.evict_last(1)
                graph.resync()


def test_int4_large_chunks_batch2():
    """Chunks of >= 128 tokens after the first call: the INT4 mma kernel by default, the dequantise-to-scratch +
    tcgen05 path under DUO_INT4_PREFILL_SCRATCH=1 — both must match the oracle's dequantise-everything attention."""
    from test_gpu_int4_attention import run

    run(8, 2, 1, 16, 48, [300, 130, 1, 256, 1, 2, 128, 1], seed=19, B=2, stage_cap=300)
    run(8, 2, 0, 16, 48, [200, 129, 1, 140], seed=20, stage_cap=200)   # no retrieval head in the layer
    run(8, 2, 2, 16, 48, [200, 129, 1, 140], seed=21, stage_cap=200)   # no streaming head in the layer


@pytest.mark.parametrize("extra", [[], ["--cuda_graph"], ["--kv_format", "int4"]])
def test_benchmark_static_harness_smoke(tmp_path, extra):
    """eval/efficiency/benchmark_static.py (reference protocol) end to end on a 2-layer random-init model."""
    import importlib.util

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("benchmark_static",
                                                  os.path.join(root, "eval", "efficiency", "benchmark_static.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    pat = os.path.join(root, "attn_patterns", "Llama-3-8B-Instruct-Gradient-1048k",
                       "lr=0.02-reg=0.05-ctx=1000_32000-multi_passkey10")
    mod.main(["--random_init", "llama3-8b-1048k", "--num_layers", "2", "--attn_load_dir", pat, "--sparsity", "0.5",
              "--max_length", "3000", "--prefilling_chunk_size", "1024", "--ctx_steps", "1", "--gen_steps", "3",
              "--output_dir", str(tmp_path)] + extra)
    lines = open(tmp_path / "benchmark_result.txt").read().splitlines()
    assert len(lines) == 9 and lines[0].startswith("Average generation time: ") and lines[5] == "Context length: 3000"


@pytest.mark.parametrize("n_full,B,q_len", [(2, 1, 1), (1, 2, 1), (2, 1, 3)])
def test_partial_attention_and_merge_reproduce_full_attention(n_full, B, q_len):
    """duo_attention_partial over two position slices of a retrieval head's cache + duo_merge_partials == attention over
    the whole cache (the building blocks of the sequence-sharded decode)."""
    import ctypes as C

    from duo_attention_b200 import _C
    from duo_attention_b200.kv_cache import DuoKVCache

    dev = torch.device("cuda:0")
    Hq, Hkv, D, N, sink, recent = 8, 2, 128, 3000, 8, 24
    g = torch.Generator().manual_seed(31 + n_full)
    lib = _C.load()
    stream = torch.cuda.current_stream(dev).cuda_stream

    def mk():
        return DuoKVCache(1, Hq, Hkv, D, [n_full], B, N + 64, sink, recent, torch.bfloat16, dev, stage_cap=N)

    cache = mk()
    qkv0 = torch.randn(B, N, (Hq + 2 * Hkv) * D, generator=g).to(torch.bfloat16).to(dev)
    cache.attend(0, qkv0, None, None, _C.ROPE_NONE, torch.empty(B, N, Hq, D, dtype=torch.bfloat16, device=dev))
    qkv = torch.randn(B, q_len, (Hq + 2 * Hkv) * D, generator=g).to(torch.bfloat16).to(dev)
    ref = torch.empty(B, q_len, Hq, D, dtype=torch.bfloat16, device=dev)
    # non-causal reference: a q_len-token "chunk" whose every row sees all N cached keys = q_len separate decode
    # calls on the same cache state; emulate by attending each token alone and evicting it again
    for t in range(q_len):
        cache.attend(0, qkv[:, t:t + 1].contiguous(), None, None, _C.ROPE_NONE, ref[:, t:t + 1])
        if t + 1 < q_len:
            cache.evict_last(1)
    n_keys = N + 1  # the last decode token is still appended
    nq = n_full * (Hq // Hkv)

    def partial(c, n):
        o = torch.full((B, q_len, Hq, D), float("nan"), dtype=torch.float32, device=dev)
        lse = torch.full((B, q_len, Hq), float("nan"), dtype=torch.float32, device=dev)
        _C.check(lib.duo_attention_partial(c.handles[0], n, qkv.data_ptr(), qkv.stride(1), o.data_ptr(), lse.data_ptr(),
                                           q_len, D ** -0.5, c.workspace.data_ptr(), c.workspace.numel(), stream))
        return o, lse

    # rows of the earlier tokens of the "chunk" saw a different last key (their own); compare the LAST token exactly
    # and the whole set against a two-slice merge of itself
    o_all, lse_all = partial(cache, n_keys)
    torch.testing.assert_close(o_all[:, -1, :nq], ref[:, -1, :nq].float(), rtol=2e-2, atol=2e-2)
    assert torch.isnan(o_all[:, :, nq:]).all(), "streaming-head rows must be left untouched"
    M1 = 1700
    c2 = mk()
    for name in ("full_k", "full_v"):
        c2.tensors[0][name][:, :, : n_keys - M1].copy_(cache.tensors[0][name][:, :, M1:n_keys])
    o1, l1 = partial(cache, M1)
    o2, l2 = partial(c2, n_keys - M1)
    o3, l3 = partial(c2, 0)  # an empty slice contributes nothing
    assert torch.isneginf(l3[:, :, :nq]).all()
    parts_o = torch.stack([o1, o2, o3]).contiguous()
    parts_l = torch.stack([l1, l2, l3]).contiguous()
    out = torch.full((B, q_len, Hq, D), 7.0, dtype=torch.bfloat16, device=dev)
    _C.check(lib.duo_merge_partials(parts_o.data_ptr(), parts_l.data_ptr(), 3, B * q_len, Hq, nq, out.data_ptr(),
                                    _C.DT_BF16, stream))
    torch.testing.assert_close(out[:, :, :nq].float(), o_all[:, :, :nq], rtol=1e-2, atol=1e-2)
    assert (out[:, :, nq:] == 7.0).all()
    # log-sum-exp of the merged slices == log-sum-exp of the whole
    mx = torch.maximum(l1, l2)
    torch.testing.assert_close((mx + torch.log2(torch.exp2(l1 - mx) + torch.exp2(l2 - mx)))[:, :, :nq],
                               lse_all[:, :, :nq], rtol=1e-4, atol=1e-3)
