"""Retrieval-heads-first weight reordering (the reference's duo_attn/patch/utils.py:6-45).

After this one-time permutation the only per-layer mask information the forward needs is the integer
``n_full``.  Works on CPU or GPU weights; it is plain index arithmetic on the
projection matrices, done before the model is moved to the GPU exactly like the reference.
"""
from __future__ import annotations

import torch


def _partition_index(head_gate: torch.Tensor, width: int, device) -> torch.Tensor:
    """Row/column permutation that moves the channels of retrieval heads (gate > 0.5) to the front,
    keeping the original order inside each class (a stable partition); ``width`` channels per head."""
    chan_is_full = torch.repeat_interleave(head_gate.to(device) > 0.5, repeats=width)
    pos = torch.arange(chan_is_full.numel(), device=device)
    return torch.cat([pos[chan_is_full], pos[~chan_is_full]])


@torch.no_grad()
def reorder_linear_weights(linear_module: torch.nn.Linear, full_attention_heads: torch.Tensor, repeat_num,
                           reorder_channel):
    """Same contract as the reference function of this name: permute output rows (``"out"``, incl.
    bias) or input columns (``"in"``) of ``linear_module`` in place and return it."""
    assert reorder_channel in ["in", "out"]
    w = linear_module.weight.data
    perm = _partition_index(full_attention_heads, repeat_num, w.device)
    if reorder_channel == "in":
        linear_module.weight.data = w.index_select(1, perm)
    else:
        linear_module.weight.data = w.index_select(0, perm)
        if linear_module.bias is not None:
            linear_module.bias.data = linear_module.bias.data.index_select(0, perm)
    return linear_module


@torch.no_grad()
def reorder_full_attn_heads(full_attention_heads: torch.Tensor):
    """``[1]*n_full + [0]*n_stream`` (in place, like the reference)."""
    n = int((full_attention_heads > 0.5).sum().item())
    full_attention_heads[:n] = 1
    full_attention_heads[n:] = 0
    return full_attention_heads
