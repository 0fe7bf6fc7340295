"""INT4-KV attention with the dequantisation folded into the K/V load vs the oracle restating
demo/w8a8kv4_llama.py:215-278 + demo/int4_kv.py (quantise on put, dequantise-everything on get)."""
import numpy as np
import pytest
import torch

from duo_attention_b200 import _C
from duo_attention_b200.kv_cache import DuoKVCache
from oracle import duo_oracle as O
from oracle import int4_oracle as Q
from parity import assert_parity

pytestmark = pytest.mark.gpu
D = 128


def run(Hq, Hkv, n_full, sink, recent, chunks, seed=0, B=1, stage_cap=8, scale_kv=1.0):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    cache = DuoKVCache(1, Hq, Hkv, D, [n_full], B, sum(chunks) + 8, sink, recent, torch.float16, dev,
                       stage_cap=stage_cap, kv_format="int4")
    past = None
    for i, S in enumerate(chunks):
        qkv = torch.randn(B, S, (Hq + 2 * Hkv) * D, generator=g).to(torch.float16)
        qkv[..., Hq * D :] *= scale_kv
        out = torch.empty(B, S, Hq, D, dtype=torch.float16, device=dev)
        cache.attend(0, qkv.to(dev), None, None, _C.ROPE_NONE, out)
        q = qkv[..., : Hq * D].reshape(B, S, Hq, D)
        k = qkv[..., Hq * D : (Hq + Hkv) * D].reshape(B, S, Hkv, D)
        v = qkv[..., (Hq + Hkv) * D :].reshape(B, S, Hkv, D)
        ref, past = O.int4_attention_core(q, k, v, past, n_full, Hq // Hkv, sink, recent)
        assert_parity(out.float().cpu(), ref.float(), f"chunk {i} (len {S})")
        assert cache.kv_seq_len == past[0].shape[2] and cache.streaming_kv_seq_len == past[1].shape[2]
    return cache


@pytest.mark.parametrize("n_full", [0, 1, 2])
def test_int4_decode_and_small_chunks(n_full):
    run(8, 2, n_full, 8, 24, [40, 1, 1, 3, 1, 30, 1, 5, 1], seed=n_full)


def test_int4_deploy_config_long():
    run(16, 4, 2, 64, 256, [700, 1, 1, 200, 1, 2, 1], seed=5, stage_cap=700)


def test_int4_split_kv_long_context():
    run(4, 1, 1, 64, 256, [12000, 1, 1, 4, 1], seed=6, stage_cap=12000)


def test_int4_odd_window_and_mha():
    run(4, 4, 2, 3, 7, [5, 1, 1, 1, 20, 1, 1, 2, 1, 1], seed=7)


def test_int4_cache_content_is_k1_quantisation():
    """What lands in the cache is bit-for-bit the oracle's K1 output (scale/zero/codes)."""
    cache = run(8, 2, 1, 4, 4, [10], seed=9)
    g = torch.Generator().manual_seed(9)
    qkv = torch.randn(1, 10, 12 * D, generator=g).to(torch.float16)
    k = qkv[..., 8 * D : 10 * D].reshape(1, 10, 2, D)
    p, s, z = Q.quantize_int4(k[:, :, 0].numpy())
    t = cache.tensors[0]
    assert np.array_equal(t["full_k"][0, 0, :10].cpu().numpy(), p[0])
    assert np.array_equal(t["full_k_scale"][0, 0, :10].cpu().numpy(), s[0, :, 0])
    assert np.array_equal(t["full_k_zero"][0, 0, :10].cpu().numpy(), z[0, :, 0])
