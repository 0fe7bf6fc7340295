"""DuoAttention on a W8A8-linear / INT4-KV serving model (the reference's demo: demo/w8a8kv4_llama.py).

The demo's model is QServe's Llama: every attention module has ONE fused ``qkv_proj`` (int8 weight ``[q_size + 2 kv_size,
hidden]`` plus a per-output-channel ``dequant_scale``) and an int8 ``o_proj``.  QServe itself is outside this repo's
scope (and absent from the image); what belongs to the hot path is

* ``enable_llama_duo_attention_eval``  (demo/w8a8kv4_llama.py:659-729): the retrieval-first reorder of the q / k / v row
  blocks of the fused weight TOGETHER WITH their ``dequant_scale`` entries, the column reorder of ``o_proj``, and the
  ``full_attention_heads`` / ``sink_size`` / ``recent_size`` attributes — any module exposing those tensors works
  (no dependency on QServe's classes);
* ``duo_w8a8kv4_attention`` — what ``LlamaAttention.forward`` (demo/w8a8kv4_llama.py:174-287) does between the fused
  projection and the output quantisation: split, flashinfer-style fp32 RoPE, INT4-KV append, mixed-head attention,
  streaming compaction — here ONE call into ``DuoKVCache.attend`` on an INT4 cache (quantise fused into the append,
  dequantise fused into the attention kernel's K/V load; no fp16 image of the cache for decode).
"""
from __future__ import annotations

import torch

from .. import _C
from .reorder import reorder_full_attn_heads


@torch.no_grad()
def reorder_linear_weights(weight, full_attention_heads: torch.Tensor, repeat_num, reorder_channel, dequant_scale=None):
    """demo/w8a8kv4_llama.py:628-656: stable partition (retrieval heads first) of the rows ("out": with their
    ``dequant_scale``) or columns ("in") of ``weight``.  Returns ``(weight, dequant_scale)``."""
    assert reorder_channel in ["in", "out"]
    mask = torch.repeat_interleave(full_attention_heads, repeats=repeat_num).to(weight.device) > 0.5
    order = torch.cat([torch.nonzero(mask).flatten(), torch.nonzero(~mask).flatten()])
    if reorder_channel == "in":
        return weight.index_select(1, order), dequant_scale
    return weight.index_select(0, order), dequant_scale.index_select(0, order.to(dequant_scale.device))


def enable_llama_duo_attention_eval(model, full_attention_heads, sink_size, recent_size):
    """Same name / arguments as demo/w8a8kv4_llama.py:659-729, for models whose attention modules carry a fused
    ``qkv_proj`` (``.weight``, ``.dequant_scale``), ``o_proj.weight``, ``q_size``, ``kv_size``, ``num_heads``,
    ``num_kv_heads`` and ``head_dim``."""
    p = next(model.parameters())
    for idx, layer in enumerate(model.model.layers):
        m = layer.self_attn
        gate = torch.tensor(full_attention_heads[idx], device=p.device, dtype=p.dtype)
        group = m.num_heads // m.num_kv_heads
        w, ds = m.qkv_proj.weight.data, m.qkv_proj.dequant_scale
        for lo, hi, rep in ((0, m.q_size, group * m.head_dim), (m.q_size, m.q_size + m.kv_size, m.head_dim),
                            (m.q_size + m.kv_size, m.q_size + 2 * m.kv_size, m.head_dim)):
            w[lo:hi], ds[lo:hi] = reorder_linear_weights(w[lo:hi], gate, rep, "out", ds[lo:hi])
        m.o_proj.weight.data, _ = reorder_linear_weights(m.o_proj.weight.data, gate, group * m.head_dim, "in")
        m.sink_size = sink_size
        m.recent_size = recent_size
        m.register_buffer("full_attention_heads", reorder_full_attn_heads(gate))


def rope_tables_fp32(pos0: int, n: int, head_dim: int, rope_theta: float, rope_scale: float, device):
    """cos / sin ``[n, head_dim]`` fp32 for positions pos0 .. pos0+n-1, flashinfer conventions
    (duo_attn/patch/flashinfer_utils.py:29-59: non-interleaved halves, ``position / rope_scale``)."""
    idx = torch.arange(head_dim // 2, dtype=torch.float32, device=device)
    freq = torch.pow(torch.tensor(float(rope_theta), device=device), -2.0 * idx / head_dim)
    ang = (torch.arange(pos0, pos0 + n, dtype=torch.float32, device=device) / float(rope_scale))[:, None] * freq[None]
    return torch.cat([ang.cos(), ang.cos()], -1).contiguous(), torch.cat([ang.sin(), ang.sin()], -1).contiguous()


def duo_w8a8kv4_attention(module, qkv_act: torch.Tensor, kv_cache, q_len: int) -> torch.Tensor:
    """The attention core of demo/w8a8kv4_llama.py:174-287.  ``qkv_act``: fp16 ``[bsz * q_len, q_size + 2 kv_size]``
    (the activation buffer the fused int8 projection wrote); returns ``[bsz * q_len, hidden]`` for the output
    quantisation + ``o_proj`` that follow.  ``kv_cache``: a ``DuoAttentionStaticINT4KVCache``."""
    width = qkv_act.shape[-1]
    bsz = qkv_act.shape[0] // q_len
    qkv = qkv_act.view(bsz, q_len, width)
    cos, sin = rope_tables_fp32(kv_cache.kv_seq_len_list[module.layer_idx], q_len, module.head_dim,
                                module.rope_theta, 1.0, qkv.device)
    out = torch.empty(bsz, q_len, module.num_heads, module.head_dim, dtype=qkv.dtype, device=qkv.device)
    kv_cache.attend(module.layer_idx, qkv, cos, sin, _C.ROPE_FP32, out)
    return out.view(bsz * q_len, module.num_heads * module.head_dim)
