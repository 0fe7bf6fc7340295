"""The W8A8KV4 demo's enable function for fused-qkv attention modules (demo/w8a8kv4_llama.py:659-729) held to a fixture
produced by RUNNING the reference's own function (tests/golden/make_golden.py, w8a8kv4_enable.npz)."""
import os
import types

import numpy as np
import torch

from duo_attention_b200.patch import w8a8kv4

HERE = os.path.dirname(os.path.abspath(__file__))


def test_enable_reorders_fused_qkv_rows_with_their_dequant_scales():
    gold = np.load(os.path.join(HERE, "golden", "w8a8kv4_enable.npz"))
    gates = gold["gates"]
    Hq, Hkv, Dh = 8, 4, 8
    layers = []
    for i in range(len(gates)):
        qkv = types.SimpleNamespace(weight=types.SimpleNamespace(data=torch.from_numpy(gold[f"qkv_w_in_{i}"].copy())),
                                    dequant_scale=torch.from_numpy(gold[f"qkv_s_in_{i}"].copy()))
        o = types.SimpleNamespace(weight=types.SimpleNamespace(data=torch.from_numpy(gold[f"o_w_in_{i}"].copy())))
        attn = types.SimpleNamespace(qkv_proj=qkv, o_proj=o, q_size=Hq * Dh, kv_size=Hkv * Dh, num_heads=Hq,
                                     num_kv_heads=Hkv, head_dim=Dh)
        attn.register_buffer = lambda name, t, a=attn: setattr(a, name, t)
        layers.append(types.SimpleNamespace(self_attn=attn))
    holder = torch.nn.Linear(1, 1).to(torch.float16)
    model = types.SimpleNamespace(parameters=lambda: holder.parameters(), model=types.SimpleNamespace(layers=layers))
    w8a8kv4.enable_llama_duo_attention_eval(model, gates.tolist(), 64, 256)
    for i, layer in enumerate(layers):
        a = layer.self_attn
        assert np.array_equal(a.qkv_proj.weight.data.numpy(), gold[f"qkv_w_{i}"])
        assert np.array_equal(a.qkv_proj.dequant_scale.numpy(), gold[f"qkv_s_{i}"])
        assert np.array_equal(a.o_proj.weight.data.numpy(), gold[f"o_w_{i}"])
        assert np.array_equal(a.full_attention_heads.float().numpy(), gold[f"heads_{i}"])
        assert a.sink_size == 64 and a.recent_size == 256


def test_rope_tables_match_the_flashinfer_restatement():
    from oracle import duo_oracle as O

    g = torch.Generator().manual_seed(0)
    q = torch.randn(1, 9, 4, 128, generator=g)
    k = torch.randn(1, 9, 2, 128, generator=g)
    qr, kr = O.rope_flashinfer(q, k, 12345, 1.0, 10000.0)
    cos, sin = w8a8kv4.rope_tables_fp32(12345, 9, 128, 10000.0, 1.0, "cpu")
    rot = lambda x: torch.cat([-x[..., 64:], x[..., :64]], -1)  # noqa: E731
    torch.testing.assert_close(q * cos[None, :, None] + rot(q) * sin[None, :, None], qr, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(k * cos[None, :, None] + rot(k) * sin[None, :, None], kr, rtol=1e-5, atol=1e-5)
