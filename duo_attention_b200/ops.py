"""Thin tensor-level wrappers over the C ABI for the caller-side glue kernels (include/duo_b200.h:
duo_add_rmsnorm, duo_silu_mul).  CUDA only — no fallback."""
from __future__ import annotations

import torch

from . import _C

LAUNCHES = 0  # kernels of libduo_b200 enqueued through this module (bench.py reports gpu_launches)


def _dt(t):
    if t.dtype == torch.bfloat16:
        return _C.DT_BF16
    if t.dtype == torch.float16:
        return _C.DT_FP16
    raise ValueError(f"dtype {t.dtype} not supported (bf16 / fp16)")


def add_rmsnorm(x: torch.Tensor, residual, weight: torch.Tensor, eps: float, inplace_residual: bool = True):
    """``h = residual + x`` (if residual is given; written back into ``residual`` when ``inplace_residual``),
    returns ``(rmsnorm(h) * weight, h)`` with HF LlamaRMSNorm arithmetic."""
    global LAUNCHES
    if not x.is_cuda:
        raise RuntimeError("duo_attention_b200 kernels need CUDA tensors (no CPU fallback)")
    hidden = x.shape[-1]
    x = x if x.is_contiguous() else x.contiguous()
    rows = x.numel() // hidden
    out = torch.empty_like(x)
    res_ptr, out_res, h = None, None, x
    if residual is not None:
        assert residual.is_contiguous() and residual.shape == x.shape and residual.dtype == x.dtype
        h = residual if inplace_residual else torch.empty_like(x)
        res_ptr, out_res = residual.data_ptr(), h.data_ptr()
    _C.check(_C.load().duo_add_rmsnorm(x.data_ptr(), res_ptr, weight.data_ptr(), out.data_ptr(), out_res, rows, hidden,
                                      float(eps), _dt(x), torch.cuda.current_stream(x.device).cuda_stream))
    LAUNCHES += 1
    return out, h


def silu_mul(gate_up: torch.Tensor):
    """``silu(gate) * up`` for ``gate_up = [..., gate | up]`` (HF LlamaMLP arithmetic)."""
    global LAUNCHES
    if not gate_up.is_cuda:
        raise RuntimeError("duo_attention_b200 kernels need CUDA tensors (no CPU fallback)")
    inter = gate_up.shape[-1] // 2
    gate_up = gate_up if gate_up.is_contiguous() else gate_up.contiguous()
    rows = gate_up.numel() // (2 * inter)
    out = torch.empty(*gate_up.shape[:-1], inter, dtype=gate_up.dtype, device=gate_up.device)
    _C.check(_C.load().duo_silu_mul(gate_up.data_ptr(), out.data_ptr(), rows, inter, _dt(gate_up),
                                   torch.cuda.current_stream(gate_up.device).cuda_stream))
    LAUNCHES += 1
    return out
