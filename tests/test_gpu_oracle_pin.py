"""Pin the oracle's restatement of flash_attn_func's contract against the installed flash_attn library —
the very kernels the reference calls (llama.py:227,239,252,...) — on the GPU box, and cross-check the
CUDA product against that library too (second, independent GPU oracle)."""
import pytest
import torch

from duo_attention_b200 import _C
from duo_attention_b200.kv_cache import DuoKVCache
from oracle import duo_oracle as O

pytestmark = pytest.mark.gpu
D = 128


@pytest.mark.parametrize("Sq,Sk,Hq,Hkv", [(1, 700, 8, 2), (37, 37, 8, 8), (16, 300, 4, 1), (128, 500, 8, 2)])
def test_contract_restatement_vs_flash_attn(Sq, Sk, Hq, Hkv):
    fa = pytest.importorskip("flash_attn")
    g = torch.Generator().manual_seed(Sq * 1000 + Sk)
    q = torch.randn(2, Sq, Hq, D, generator=g).to(torch.bfloat16)
    k = torch.randn(2, Sk, Hkv, D, generator=g).to(torch.bfloat16)
    v = torch.randn(2, Sk, Hkv, D, generator=g).to(torch.bfloat16)
    ref = O.flash_attn_contract(q, k, v, causal=True)
    got = fa.flash_attn_func(q.cuda(), k.cuda(), v.cuda(), causal=True, dropout_p=0.0).cpu()
    torch.testing.assert_close(got.float(), ref.float(), rtol=1e-2, atol=2e-3)


def test_product_vs_reference_forward_restated_with_flash_attn():
    """llama.py:374-421 restated verbatim on the GPU with the installed flash_attn_func, token-major caches,
    torch.cat and compaction — vs the fused B200 path."""
    fa = pytest.importorskip("flash_attn")
    dev = torch.device("cuda:0")
    Hq, Hkv, n_full, sink, recent = 32, 8, 3, 64, 256
    G = Hq // Hkv
    g = torch.Generator().manual_seed(0)
    cache = DuoKVCache(1, Hq, Hkv, D, [n_full], 1, 4096, sink, recent, torch.bfloat16, dev, stage_cap=1024)
    fk = fv = sk = sv = None
    for S in [1000, 1, 1, 600, 1, 64, 1]:
        qkv = torch.randn(1, S, (Hq + 2 * Hkv) * D, generator=g).to(torch.bfloat16).to(dev)
        q = qkv[..., : Hq * D].reshape(1, S, Hq, D)
        k = qkv[..., Hq * D : (Hq + Hkv) * D].reshape(1, S, Hkv, D)
        v = qkv[..., (Hq + Hkv) * D :].reshape(1, S, Hkv, D)
        if fk is None:
            ref = fa.flash_attn_func(q, k, v, causal=True)
            fk, fv, sk, sv = k[:, :, :n_full], v[:, :, :n_full], k[:, :, n_full:], v[:, :, n_full:]
        else:
            fk = torch.cat([fk, k[:, :, :n_full]], 1)
            fv = torch.cat([fv, v[:, :, :n_full]], 1)
            sk = torch.cat([sk, k[:, :, n_full:]], 1)
            sv = torch.cat([sv, v[:, :, n_full:]], 1)
            a = fa.flash_attn_func(q[:, :, : n_full * G], fk, fv, causal=True)
            b = fa.flash_attn_func(q[:, :, n_full * G :], sk, sv, causal=True)
            ref = torch.cat([a, b], dim=2)
        if sk.shape[1] > sink + recent:
            sk = torch.cat([sk[:, :sink], sk[:, -recent:]], 1)
            sv = torch.cat([sv[:, :sink], sv[:, -recent:]], 1)
        out = torch.empty(1, S, Hq, D, dtype=torch.bfloat16, device=dev)
        cache.attend(0, qkv.clone(), None, None, _C.ROPE_NONE, out)
        torch.testing.assert_close(out.float(), ref.float(), rtol=1e-2, atol=2e-3)
