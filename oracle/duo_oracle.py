"""CPU oracle for DuoAttention's mixed-head attention hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``duo_attention_b200``)
imports this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only as the
checker / the timed CPU baseline.

It restates, with plain PyTorch ops on CPU, the reference's deploy-time forward:

* ``flash_attn_contract``      – the public contract of ``flash_attn.flash_attn_func``
  (third-party; pinned 2.6.3 in the reference README:44, 2.8.3 installed here) at the
  reference's call sites ``duo_attn/patch/llama.py:227,239,252,366,394,406``.
* ``tuple_forward``            – ``duo_attn/patch/llama.py:146-306``
  (``llama_duo_attention_forward_one_way_reordered``; mistral twin is identical).
* ``OracleStaticKVCache``      – ``duo_attn/patch/static_kv_cache.py:18-315``.
* ``static_forward``           – ``duo_attn/patch/llama.py:309-434``.
* ``rope_flashinfer``          – ``flashinfer.rope.apply_rope_inplace`` as called from
  ``duo_attn/patch/flashinfer_utils.py:29-59`` (non-interleaved, pos = offset + i).
* ``load_attn_pattern`` / ``sparsify_attention_heads`` – ``duo_attn/utils.py:326-336,353-373``.
* ``reorder_linear_weights`` / ``reorder_full_attn_heads`` – ``duo_attn/patch/utils.py:6-45``.
* ``streaming_visible``        – closed form of the deploy-time streaming mask implied by
  ``llama.py:202-223,273-290``.
* ``OracleModel``              – ``duo_attn/patch/tuple_kv_cache.py:241-490`` (layer loop,
  position ids from the layer-0 cache length, last-token logits).

Parity pin: the reference ships no tests or golden vectors.  The oracle is
pinned instead against the reference's *own code* executed in the build container with
only ``flash_attn_func`` / ``apply_rope_inplace`` (CUDA-only third-party calls) swapped for
the contract restatements in this file — see ``tests/golden/make_golden.py`` and the
fixtures it writes — and, on the GPU box, ``flash_attn_contract`` is checked against the
installed ``flash_attn_func`` itself (``tests/test_gpu_oracle_pin.py``).  The same script also runs
the reference's patched Llama/Mistral MODELS (tuple and static drivers) and its INT4 cache class +
W8A8KV4 attention forward; ``OracleModel`` and ``int4_attention_core`` are held to those outputs
(``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import json
import math
import os
from typing import List, Optional, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# flash_attn_func contract
# --------------------------------------------------------------------------------------


def flash_attn_contract(q, k, v, causal=True, dropout_p=0.0, softmax_scale=None, block=2048):
    """``flash_attn_func(q, k, v, causal=True)`` restated.

    q ``[B, Sq, Hq, D]``, k/v ``[B, Sk, Hkv, D]``; q-head ``i`` reads kv-head
    ``i // (Hq // Hkv)``; ``causal`` is BOTTOM-RIGHT aligned (row ``i`` sees keys
    ``j <= i + Sk - Sq``); scale ``1/sqrt(D)``; softmax in fp32; the un-normalised P = exp(s - max) is
    cast to the input dtype before the PV product and the division by the fp32 row sum happens last
    (FA2 behaviour); output in the input dtype.
    Processes query rows in blocks so that [Sq, Sk] never has to exist at once.
    """
    assert dropout_p == 0.0
    B, Sq, Hq, D = q.shape
    _, Sk, Hkv, _ = k.shape
    G = Hq // Hkv
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(D)
    out = torch.empty_like(q)
    kf = k.float().permute(0, 2, 3, 1)  # [B,Hkv,D,Sk]
    vdt = v.permute(0, 2, 1, 3)  # [B,Hkv,Sk,D] in input dtype
    jj = torch.arange(Sk)
    for r0 in range(0, Sq, block):
        r1 = min(Sq, r0 + block)
        qb = q[:, r0:r1].float().permute(0, 2, 1, 3)  # [B,Hq,R,D]
        qb = qb.reshape(B, Hkv, G * (r1 - r0), D)
        s = torch.matmul(qb, kf) * scale  # [B,Hkv,G*R,Sk]
        s = s.view(B, Hkv, G, r1 - r0, Sk)
        if causal:
            ii = torch.arange(r0, r1)
            masked = jj[None, :] > (ii[:, None] + (Sk - Sq))
            s = s.masked_fill(masked[None, None, None], float("-inf"))
        # FA2 arithmetic: P = exp(s - rowmax) is rounded to the input dtype UN-normalised, the row sum l is
        # accumulated in fp32 from the unrounded P, and O = (P_rounded @ V) / l at the very end.
        m = s.amax(dim=-1, keepdim=True)
        m = torch.where(torch.isinf(m), torch.zeros_like(m), m)  # rows with no visible key -> output 0
        p = torch.exp(s - m)
        l = p.sum(dim=-1, keepdim=True)
        p = p.to(q.dtype).float().view(B, Hkv, G * (r1 - r0), Sk)
        o = torch.matmul(p, vdt.float())  # [B,Hkv,G*R,D]
        l = l.view(B, Hkv, G * (r1 - r0), 1)
        o = torch.where(l > 0, o / l.clamp_min(1e-38), torch.zeros_like(o))
        o = o.view(B, Hq, r1 - r0, D).permute(0, 2, 1, 3)
        out[:, r0:r1] = o.to(q.dtype)
    return out


# --------------------------------------------------------------------------------------
# RoPE
# --------------------------------------------------------------------------------------


def rotate_half(x):
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2 :]
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary_pos_emb_hf(q, k, cos, sin, unsqueeze_dim=2):
    """HF ``apply_rotary_pos_emb`` as called at llama.py:177-184 (``unsqueeze_dim=2``).

    cos/sin ``[B, S, D]`` already in the activation dtype; every op rounds to that dtype."""
    cos = cos.unsqueeze(unsqueeze_dim)
    sin = sin.unsqueeze(unsqueeze_dim)
    q_embed = (q * cos) + (rotate_half(q) * sin)
    k_embed = (k * cos) + (rotate_half(k) * sin)
    return q_embed, k_embed


def hf_cos_sin(position_ids, head_dim, rope_theta, dtype, rope_factor=1.0, attention_scaling=1.0):
    """``LlamaRotaryEmbedding.forward`` (default / linear scaling): fp32 tables cast to dtype."""
    inv_freq = 1.0 / (rope_theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    inv_freq = inv_freq / rope_factor
    freqs = position_ids[:, :, None].float() * inv_freq[None, None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return (emb.cos() * attention_scaling).to(dtype), (emb.sin() * attention_scaling).to(dtype)


def rope_flashinfer(q, k, offset, rope_scale, rope_theta):
    """``flashinfer.rope.apply_rope_inplace(interleave=False)`` as used by the static path
    (flashinfer_utils.py:29-59): fp32 on-the-fly trig, ``pos = offset + i``,
    ``freq_i = theta^(-2 (i mod D/2)/D) / rope_scale``; pairs ``(i, i + D/2)``.
    Returns new tensors (the reference mutates in place)."""
    B, S, _, D = q.shape
    pos = (torch.arange(S, dtype=torch.float32) + float(offset))[None, :, None, None]
    idx = torch.arange(D // 2, dtype=torch.float32)
    freq = torch.pow(torch.tensor(float(rope_theta)), -2.0 * idx / D) / rope_scale
    ang = pos * freq[None, None, None, :]
    cos, sin = torch.cos(ang), torch.sin(ang)

    def rot(x):
        xf = x.float()
        x1, x2 = xf[..., : D // 2], xf[..., D // 2 :]
        return torch.cat((x1 * cos - x2 * sin, x2 * cos + x1 * sin), dim=-1).to(x.dtype)

    return rot(q), rot(k)


# --------------------------------------------------------------------------------------
# Pattern loading / sparsify / reorder
# --------------------------------------------------------------------------------------


def load_attn_pattern(attn_load_dir):
    """duo_attn/utils.py:326-336."""
    h = np.loadtxt(os.path.join(attn_load_dir, "full_attention_heads.tsv"), dtype=float, delimiter="\t")
    h = np.clip(h, 0, 1)
    with open(os.path.join(attn_load_dir, "config.json")) as f:
        cfg = json.load(f)
    return h, cfg["sink_size"], cfg["recent_size"]


def sparsify_attention_heads(full_attention_heads, threshold=None, sparsity=None):
    """duo_attn/utils.py:353-373 (tie-break noise is added IN PLACE with np.random)."""
    full_attention_heads += np.random.uniform(0, 1e-6, full_attention_heads.shape)
    if sparsity is not None:
        threshold = np.quantile(full_attention_heads, sparsity)
    else:
        assert threshold is not None, "Either threshold or sparsity must be provided"
    if sparsity >= 1:
        threshold = 2
    if sparsity <= 0:
        threshold = -1
    full_attention_heads = (full_attention_heads >= threshold).astype(float)
    sparsity = 1 - np.mean(full_attention_heads)
    return full_attention_heads, sparsity


@torch.no_grad()
def reorder_rows_or_cols(weight, bias, head_mask, repeat_num, channel):
    """duo_attn/patch/utils.py:6-34 on raw tensors: stable partition, full heads first."""
    m = torch.repeat_interleave(head_mask, repeats=repeat_num) > 0.5
    if channel == "in":
        w = torch.cat([weight[:, m], weight[:, ~m]], dim=1)
    else:
        w = torch.cat([weight[m, :], weight[~m, :]], dim=0)
    b = None
    if bias is not None:
        b = torch.cat([bias[m], bias[~m]], dim=0)
    return w, b


def reorder_full_attn_heads(head_mask):
    """duo_attn/patch/utils.py:37-45."""
    n = int((head_mask > 0.5).sum().item())
    out = torch.zeros_like(head_mask)
    out[:n] = 1
    return out


# --------------------------------------------------------------------------------------
# Closed-form deploy-time streaming mask
# --------------------------------------------------------------------------------------


def streaming_visible(t, j, cs, sink, recent):
    """Is absolute key ``j`` visible to absolute query ``t`` of a chunk starting at ``cs``
    on a streaming head?  ``cs == 0`` (first call) is plain causal."""
    if j > t:
        return False
    if cs <= sink + recent:
        return True
    return j < sink or j >= cs - recent


def training_streaming_mask(seq_len, sink, recent):
    """Training-time streaming (Lambda) mask, True = visible (duo_attn/patch/streaming_attn.py:14-24): query t sees key j
    iff ``j <= t and (j < sink or j > t - recent)`` — the window INCLUDES the query itself, so it holds ``recent`` keys
    where the deploy-time decode step sees ``recent + 1`` (ring + the new token).  Padded to a multiple of 8 like the
    reference."""
    n = (seq_len + 7) // 8 * 8
    t = torch.arange(n)[:, None]
    j = torch.arange(n)[None, :]
    return (j <= t) & ((j < sink) | (j > t - recent))


def training_streaming_attention(q, k, v, sink, recent):
    """Whole-sequence streaming attention under the training-time mask (streaming_attn.py:27-42: SDPA with the boolean
    mask, GQA by head repetition, scale 1/sqrt(D)).  q ``[B,S,Hq,D]``, k/v ``[B,S,Hkv,D]``; fp32 math."""
    B, S, Hq, D = q.shape
    g = Hq // k.shape[2]
    mask = training_streaming_mask(S, sink, recent)[:S, :S]
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float().repeat_interleave(g, dim=2)) / math.sqrt(D)
    s = s.masked_fill(~mask[None, None], float("-inf"))
    return torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), v.float().repeat_interleave(g, dim=2))


# --------------------------------------------------------------------------------------
# Attention-layer oracle, tuple cache  (llama.py:146-306)
# --------------------------------------------------------------------------------------


class AttnWeights:
    """The (already reordered) projections of one attention layer + its head split."""

    def __init__(self, wq, wk, wv, wo, num_heads, num_kv_heads, num_full_kv_heads, bq=None, bk=None, bv=None, bo=None):
        self.wq, self.wk, self.wv, self.wo = wq, wk, wv, wo
        self.bq, self.bk, self.bv, self.bo = bq, bk, bv, bo
        self.num_heads = num_heads
        self.num_kv_heads = num_kv_heads
        self.head_dim = wq.shape[0] // num_heads
        self.groups = num_heads // num_kv_heads
        self.n_full = int(num_full_kv_heads)
        self.n_stream = num_kv_heads - self.n_full


def _lin(x, w, b):
    return torch.nn.functional.linear(x, w, b)


def tuple_attention_core(q, k, v, past, n_full, groups, sink, recent, use_cache=True):
    """Everything of llama.py:168-301 after RoPE: split, cat with the tuple cache, the two
    attention calls, and the streaming compaction.  q ``[B,S,Hq,D]``, k/v ``[B,S,Hkv,D]``
    (post-RoPE).  ``past`` is ``None`` or ``(full_KV [2B,n_f,N,D], stream_KV [2B,n_s,<=W,D])``.
    Returns ``(attn_out [B,S,Hq,D], new_past)``."""
    B, S = q.shape[:2]
    kv_seq_len = S + (past[0].shape[2] if past is not None else 0)
    n_full_q = n_full * groups
    fk, fv = k[:, :, :n_full], v[:, :, :n_full]
    sk, sv = k[:, :, n_full:], v[:, :, n_full:]
    if past is not None:
        pf = past[0].transpose(1, 2)
        ps = past[1].transpose(1, 2)
        fk = torch.cat([pf[:B], fk], dim=1)
        fv = torch.cat([pf[B:], fv], dim=1)
        sk = torch.cat([ps[:B], sk], dim=1)
        sv = torch.cat([ps[B:], sv], dim=1)
    if S == kv_seq_len:
        out = flash_attn_contract(q, k, v, causal=True)
    else:
        outs = []
        if n_full > 0:
            outs.append(flash_attn_contract(q[:, :, :n_full_q], fk, fv, causal=True))
        if q.shape[2] - n_full_q > 0:
            outs.append(flash_attn_contract(q[:, :, n_full_q:], sk, sv, causal=True))
        out = outs[0] if len(outs) == 1 else torch.cat(outs, dim=2)
    if sk.shape[1] > recent + sink:
        sk = torch.cat([sk[:, :sink], sk[:, -recent:]], dim=1)
        sv = torch.cat([sv[:, :sink], sv[:, -recent:]], dim=1)
    new_past = None
    if use_cache:
        new_past = (
            torch.cat([fk, fv], dim=0).transpose(1, 2).contiguous(),
            torch.cat([sk, sv], dim=0).transpose(1, 2).contiguous(),
        )
    return out, new_past


def int4_roundtrip(x):
    """quantise -> dequantise of an fp16 ``[..., 128]`` tensor with the reference's K1/K2 arithmetic
    (what ``DuoAttentionStaticINT4KVCache.put`` + ``get`` hand to attention, demo/int4_kv.py:261-436)."""
    from oracle import int4_oracle as Q

    a = x.detach().cpu().numpy()
    p, s, z = Q.quantize_int4(a)
    return torch.from_numpy(Q.dequantize_int4(p, s, z)).to(x.dtype)


def int4_attention_core(q, k, v, past, n_full, groups, sink, recent):
    """demo/w8a8kv4_llama.py:215-278 after RoPE, fp16: the cache holds quantised K/V; the FIRST call attends
    the raw k/v (:229-238), every later call attends the quantise->dequantise round trip of everything,
    the new tokens included (:222-227,239-274).  Same tuple layout for ``past`` as tuple_attention_core,
    holding the round-tripped values."""
    kq, vq = int4_roundtrip(k), int4_roundtrip(v)
    if past is None:
        out = flash_attn_contract(q, k, v, causal=True)
        _, new_past = tuple_attention_core(q, kq, vq, None, n_full, groups, sink, recent)
        return out, new_past
    return tuple_attention_core(q, kq, vq, past, n_full, groups, sink, recent)


def tuple_forward(w: AttnWeights, hidden, cos, sin, past, sink, recent, core=None):
    """llama.py:146-306.  hidden ``[B,S,hidden]``; cos/sin ``[B,S,D]`` in hidden.dtype.  ``core``: the attention
    core after RoPE (default tuple_attention_core; int4_attention_core for the INT4-KV demo forward)."""
    B, S, _ = hidden.shape
    q = _lin(hidden, w.wq, w.bq).view(B, S, w.num_heads, w.head_dim)
    k = _lin(hidden, w.wk, w.bk).view(B, S, w.num_kv_heads, w.head_dim)
    v = _lin(hidden, w.wv, w.bv).view(B, S, w.num_kv_heads, w.head_dim)
    q, k = apply_rotary_pos_emb_hf(q, k, cos, sin, unsqueeze_dim=2)
    out, new_past = (core or tuple_attention_core)(q, k, v, past, w.n_full, w.groups, sink, recent)
    out = out.reshape(B, S, w.num_heads * w.head_dim)
    return _lin(out, w.wo, w.bo), new_past


# --------------------------------------------------------------------------------------
# Static KV cache  (static_kv_cache.py:18-315) and static forward (llama.py:309-434)
# --------------------------------------------------------------------------------------


class OracleStaticKVCache:
    """Token-major pre-allocated cache, semantics of DuoAttentionStaticKVCache."""

    def __init__(self, num_layers, num_heads, num_kv_heads, head_dim, full_attention_heads,
                 batch_size, max_size, sink_size, recent_size, dtype=torch.float32):
        self.batch_size, self.max_size = batch_size, max_size
        self.sink_size, self.recent_size = sink_size, recent_size
        self.num_layers, self.num_heads, self.num_kv_heads = num_layers, num_heads, num_kv_heads
        self.num_kv_groups = num_heads // num_kv_heads
        self.head_dim = head_dim
        self.num_full_kv_head_list = []
        self.num_streaming_kv_head_list = []
        self.kv_seq_len_list = [0] * num_layers
        self.streaming_kv_seq_len_list = [0] * num_layers
        self.full_k, self.full_v, self.str_k, self.str_v = [], [], [], []
        W = sink_size + recent_size
        for row in full_attention_heads:
            nf = int((torch.as_tensor(np.asarray(row)) > 0.5).sum().item())
            ns = num_kv_heads - nf
            self.num_full_kv_head_list.append(nf)
            self.num_streaming_kv_head_list.append(ns)
            self.full_k.append(torch.zeros(batch_size, max_size, nf, head_dim, dtype=dtype))
            self.full_v.append(torch.zeros(batch_size, max_size, nf, head_dim, dtype=dtype))
            self.str_k.append(torch.zeros(batch_size, W, ns, head_dim, dtype=dtype))
            self.str_v.append(torch.zeros(batch_size, W, ns, head_dim, dtype=dtype))

    @property
    def kv_seq_len(self):
        return self.kv_seq_len_list[-1]

    @property
    def streaming_kv_seq_len(self):
        return self.streaming_kv_seq_len_list[-1]

    def put_full_kv(self, l, fk, fv):
        inc, cur = fk.shape[1], self.kv_seq_len_list[l]
        if inc + cur > self.max_size:
            raise ValueError(
                f"Trying to put {inc} KVs into a cache with max size {self.max_size}, current size: {cur}."
            )
        self.full_k[l][:, cur : cur + inc].copy_(fk)
        self.full_v[l][:, cur : cur + inc].copy_(fv)
        self.kv_seq_len_list[l] += inc
        n = self.kv_seq_len_list[l]
        return self.full_k[l][:, :n], self.full_v[l][:, :n]

    def get_streaming_kv(self, l):
        n = self.streaming_kv_seq_len_list[l]
        return self.str_k[l][:, :n], self.str_v[l][:, :n]

    def compress_and_replace_streaming_kv(self, l, sk, sv):
        inc = sk.shape[1]
        W = self.sink_size + self.recent_size
        if inc <= W:
            self.str_k[l][:, :inc].copy_(sk)
            self.str_v[l][:, :inc].copy_(sv)
            self.streaming_kv_seq_len_list[l] = inc
        else:
            s, r = self.sink_size, self.recent_size
            self.str_k[l][:, :s].copy_(sk[:, :s])
            self.str_k[l][:, s : s + r].copy_(sk[:, inc - r : inc])
            self.str_v[l][:, :s].copy_(sv[:, :s])
            self.str_v[l][:, s : s + r].copy_(sv[:, inc - r : inc])
            self.streaming_kv_seq_len_list[l] = W

    def split_kv(self, l, k, v):
        nf = self.num_full_kv_head_list[l]
        return k[:, :, :nf], v[:, :, :nf], k[:, :, nf:], v[:, :, nf:]

    def clear(self):
        for l in range(self.num_layers):
            self.kv_seq_len_list[l] = 0
            self.streaming_kv_seq_len_list[l] = 0

    def evict_last(self, n):
        for l in range(self.num_layers):
            self.kv_seq_len_list[l] = max(0, self.kv_seq_len_list[l] - n)
            self.streaming_kv_seq_len_list[l] = max(0, self.streaming_kv_seq_len_list[l] - n)

    @property
    def memory_usage(self):
        tot = 0
        for lst in (self.full_k, self.full_v, self.str_k, self.str_v):
            for t in lst:
                tot += t.element_size() * t.numel()
        return tot


def static_attention_core(q, k, v, kv_cache: OracleStaticKVCache, layer_idx):
    """llama.py:354-425 (post-RoPE part of the static forward)."""
    B, S = q.shape[:2]
    kv_seq_len = S + kv_cache.kv_seq_len
    groups = q.shape[2] // k.shape[2]
    fk, fv, sk, sv = kv_cache.split_kv(layer_idx, k, v)
    fk, fv = kv_cache.put_full_kv(layer_idx, fk, fv)
    if S == kv_seq_len:
        out = flash_attn_contract(q, k, v, causal=True)
    else:
        nfq = kv_cache.num_full_kv_head_list[layer_idx] * groups
        ck, cv = kv_cache.get_streaming_kv(layer_idx)
        sk = torch.cat([ck, sk], dim=1)
        sv = torch.cat([cv, sv], dim=1)
        outs = []
        if nfq > 0:
            outs.append(flash_attn_contract(q[:, :, :nfq], fk, fv, causal=True))
        if q.shape[2] - nfq > 0:
            outs.append(flash_attn_contract(q[:, :, nfq:], sk, sv, causal=True))
        out = outs[0] if len(outs) == 1 else torch.cat(outs, dim=2)
    kv_cache.compress_and_replace_streaming_kv(layer_idx, sk, sv)
    return out


def static_forward(w: AttnWeights, hidden, position_ids, kv_cache, layer_idx, rope_theta, rope_scale=1.0):
    """llama.py:309-434 with flashinfer RoPE restated."""
    B, S, _ = hidden.shape
    q = _lin(hidden, w.wq, w.bq).view(B, S, w.num_heads, w.head_dim)
    k = _lin(hidden, w.wk, w.bk).view(B, S, w.num_kv_heads, w.head_dim)
    v = _lin(hidden, w.wv, w.bv).view(B, S, w.num_kv_heads, w.head_dim)
    q, k = rope_flashinfer(q, k, int(position_ids[0, 0]), rope_scale, rope_theta)
    out = static_attention_core(q, k, v, kv_cache, layer_idx)
    out = out.reshape(B, S, w.num_heads * w.head_dim)
    return _lin(out, w.wo, w.bo)


# --------------------------------------------------------------------------------------
# Dense-mask cross-check (independent formulation used to validate the two above)
# --------------------------------------------------------------------------------------


def dense_duo_attention(q_all, k_all, v_all, chunk_starts, n_full, groups, sink, recent):
    """Whole-sequence attention with the closed-form mask; q_all/k_all/v_all hold ALL N tokens
    (post-RoPE), ``chunk_starts`` the chunk schedule.  fp64, O(N^2) – small cases only."""
    B, N, Hq, D = q_all.shape
    qd, kd, vd = q_all.double(), k_all.double(), v_all.double()
    out = torch.zeros_like(qd)
    bounds = list(chunk_starts) + [N]
    cs_of = np.zeros(N, dtype=np.int64)
    for a, b in zip(bounds[:-1], bounds[1:]):
        cs_of[a:b] = a
    for h in range(Hq):
        kvh = h // groups
        full = kvh < n_full
        s = torch.einsum("bqd,bkd->bqk", qd[:, :, h], kd[:, :, kvh]) / math.sqrt(D)
        mask = torch.zeros(N, N, dtype=torch.bool)
        for t in range(N):
            for j in range(N):
                vis = j <= t if full else streaming_visible(t, j, int(cs_of[t]), sink, recent)
                mask[t, j] = not vis
        s = s.masked_fill(mask[None], float("-inf"))
        out[:, :, h] = torch.einsum("bqk,bkd->bqd", torch.softmax(s, -1), vd[:, :, kvh])
    return out


# --------------------------------------------------------------------------------------
# Model-level oracle  (tuple_kv_cache.py:241-490)
# --------------------------------------------------------------------------------------


def rms_norm(x, weight, eps):
    xf = x.float()
    var = xf.pow(2).mean(-1, keepdim=True)
    return weight * (xf * torch.rsqrt(var + eps)).to(x.dtype)


class OracleModel:
    """Llama/Mistral decoder run the way the reference's patched HF-4.34-style driver runs
    it.  Holds plain tensors copied from an (un-patched) HF model plus the head pattern."""

    def __init__(self, hf_model, full_attention_heads, sink, recent, kv_format="same"):
        cfg = hf_model.config
        self.cfg = cfg
        # "int4": every layer's attention core is the INT4-KV one (demo/w8a8kv4_llama.py:215-278)
        self.core = int4_attention_core if kv_format == "int4" else None
        self.sink, self.recent = sink, recent
        self.num_heads = cfg.num_attention_heads
        self.num_kv_heads = cfg.num_key_value_heads
        self.head_dim = getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads
        self.eps = cfg.rms_norm_eps
        m = hf_model.model
        self.embed = m.embed_tokens.weight.detach().clone()
        self.norm_w = m.norm.weight.detach().clone()
        self.lm_head = hf_model.lm_head.weight.detach().clone()
        self.rotary = m.rotary_emb
        self.layers = []
        G = self.num_heads // self.num_kv_heads
        for idx, layer in enumerate(m.layers):
            a = layer.self_attn
            mask = torch.tensor(np.asarray(full_attention_heads[idx]), dtype=torch.float32)
            wq, bq = reorder_rows_or_cols(a.q_proj.weight.detach(), _b(a.q_proj), mask, G * self.head_dim, "out")
            wk, bk = reorder_rows_or_cols(a.k_proj.weight.detach(), _b(a.k_proj), mask, self.head_dim, "out")
            wv, bv = reorder_rows_or_cols(a.v_proj.weight.detach(), _b(a.v_proj), mask, self.head_dim, "out")
            wo, _ = reorder_rows_or_cols(a.o_proj.weight.detach(), None, mask, G * self.head_dim, "in")
            w = AttnWeights(wq, wk, wv, wo, self.num_heads, self.num_kv_heads, int((mask > 0.5).sum()),
                            bq, bk, bv, _b(a.o_proj))
            self.layers.append(dict(
                attn=w,
                ln1=layer.input_layernorm.weight.detach().clone(),
                ln2=layer.post_attention_layernorm.weight.detach().clone(),
                gate=layer.mlp.gate_proj.weight.detach().clone(),
                up=layer.mlp.up_proj.weight.detach().clone(),
                down=layer.mlp.down_proj.weight.detach().clone(),
            ))

    @torch.no_grad()
    def __call__(self, input_ids, past_key_values=None):
        """Returns ``(logits [B,1,V] fp32, new_past)`` (last-token logits only, .float(),
        tuple_kv_cache.py:283-288)."""
        B, S = input_ids.shape
        past_len = 0 if past_key_values is None else past_key_values[0][0].shape[2]
        position_ids = torch.arange(past_len, past_len + S, dtype=torch.long)[None]
        h = torch.nn.functional.embedding(input_ids, self.embed)
        cos, sin = self.rotary(h, position_ids)
        new_past = []
        for idx, L in enumerate(self.layers):
            res = h
            x = rms_norm(h, L["ln1"], self.eps)
            past = None if past_key_values is None else past_key_values[idx]
            a, p = tuple_forward(L["attn"], x, cos, sin, past, self.sink, self.recent, core=self.core)
            new_past.append(p)
            h = res + a
            res = h
            x = rms_norm(h, L["ln2"], self.eps)
            x = _lin(torch.nn.functional.silu(_lin(x, L["gate"], None)) * _lin(x, L["up"], None), L["down"], None)
            h = res + x
        h = rms_norm(h, self.norm_w, self.eps)
        logits = _lin(h[:, -1:, :], self.lm_head, None).float()
        return logits, tuple(new_past)


def _b(lin):
    return None if lin.bias is None else lin.bias.detach().clone()
