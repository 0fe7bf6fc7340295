"""Device pieces of the sequence-sharded retrieval-head decode (scope row f1): slice attention + cross-slice merge."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


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
