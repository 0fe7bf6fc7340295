"""Pattern loading / head sparsification — the host side of the drop-in API.

Mirrors ``duo_attn/utils.py`` of the reference for the functions the inference path needs
(:326-336 ``load_attn_pattern``, :353-373 ``sparsify_attention_heads``, :339-350 ``seed_everything``)
without that module's training/eval dependencies (accelerate, matplotlib, tensor_parallel).
Pure NumPy host logic; nothing here touches the GPU.
"""
from __future__ import annotations

import json
import os
import random

import numpy as np

PATTERN_FILE = "full_attention_heads.tsv"
CONFIG_FILE = "config.json"


def load_attn_pattern(attn_load_dir):
    """Read ``<dir>/full_attention_heads.tsv`` (``[layers, kv_heads]`` gate values) and the training
    ``config.json``.  Returns ``(float64 array clipped to [0, 1], sink_size, recent_size)`` exactly as
    the reference (duo_attn/utils.py:326-336)."""
    gates = np.loadtxt(os.path.join(attn_load_dir, PATTERN_FILE), dtype=float, delimiter="\t")
    gates = np.clip(gates, 0, 1)
    with open(os.path.join(attn_load_dir, CONFIG_FILE)) as fh:
        cfg = json.load(fh)
    return gates, cfg["sink_size"], cfg["recent_size"]


def sparsify_attention_heads(full_attention_heads, threshold=None, sparsity=None):
    """Binarise the gate matrix (duo_attn/utils.py:353-373).

    Reference behaviour kept on purpose, including its quirks: tie-breaking noise ``U(0, 1e-6)`` is
    added IN PLACE from NumPy's global RNG; with ``sparsity`` given the threshold is that quantile;
    ``sparsity >= 1`` prunes everything, ``sparsity <= 0`` keeps everything; ``sparsity=None``
    raises (the reference crashes comparing ``None >= 1``) so ``threshold`` alone is not usable.
    Returns ``(binary float array, realised sparsity)``.
    """
    full_attention_heads += np.random.uniform(0, 1e-6, full_attention_heads.shape)
    if sparsity is not None:
        threshold = np.quantile(full_attention_heads, sparsity)
    else:
        assert threshold is not None, "Either threshold or sparsity must be provided"
    if sparsity >= 1:
        threshold = 2
    if sparsity <= 0:
        threshold = -1
    mask = (full_attention_heads >= threshold).astype(float)
    return mask, 1 - np.mean(mask)


def seed_everything(seed):
    """duo_attn/utils.py:339-350."""
    import torch

    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = True
