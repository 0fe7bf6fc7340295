"""NumPy oracle for the reference's INT4 KV quantisation (TEST INFRASTRUCTURE ONLY).

Restates ``demo/quantize_int4.cu`` of the reference:

* ``quantize_int4``   – kernel K1, quantize_int4.cu:73-144: per group of ``group_size`` (=128 =
  head_dim) fp16 values, fp32 min/max scan, ``scale = (max-min)/15 + 1e-8``, ``zero = min``,
  ``q = clamp(roundf((x-zero)/scale), 0, 15)`` (roundf = half away from zero), pack
  ``(q_even << 4) | q_odd``; scale / zero stored as fp16 (round-to-nearest-even).
  NOTE the quantisation divides by the *fp32* scale, the fp16-rounded one is only stored.
* ``dequantize_int4`` – kernel K2, quantize_int4.cu:9-42: ``out = __hadd(__hmul(half(q), s), z)``.
  AS BUILT (nvcc, default -fmad=true, also under the reference's --use_fast_math) the mul+add pair is
  contracted into a single ``HFMA2`` — checked in the SASS of oracle/_ref/quantize_int4_ref.so — so the
  result is ``fp16(q*s + z)`` with ONE rounding.  The oracle restates that as-built behaviour
  (``fused=True``); ``fused=False`` gives the literal two-rounding reading of the source.

The reference builds with ``--use_fast_math`` (demo/int4_kv.py:54), which turns the fp32
division into an approximate one: codes may differ by +-1 on (near-)exact .5 ties.  The oracle
uses IEEE division; tests allow that tie slack for codes and demand bit-exact scale/zero.
"""
from __future__ import annotations

import numpy as np


def quantize_int4(x: np.ndarray, group_size: int = 128):
    """x: float16 array ``[..., head_dim]`` -> (packed uint8 ``[..., head_dim//2]``,
    scale fp16 ``[..., head_dim//group_size]``, zero fp16 same shape)."""
    assert x.dtype == np.float16
    hd = x.shape[-1]
    ng = hd // group_size
    xf = x.astype(np.float32).reshape(*x.shape[:-1], ng, group_size)
    gmin = xf.min(axis=-1, keepdims=True)
    gmax = xf.max(axis=-1, keepdims=True)
    scale = ((gmax - gmin) / np.float32(15.0) + np.float32(1e-8)).astype(np.float32)
    zero = gmin
    qf = (xf - zero) / scale
    # roundf: half away from zero (qf >= 0 here up to rounding, keep the general form)
    qf = np.sign(qf) * np.floor(np.abs(qf) + np.float32(0.5))
    qf = np.clip(qf, 0.0, 15.0)
    q = qf.astype(np.uint8).reshape(*x.shape[:-1], hd)
    packed = ((q[..., 0::2] << 4) | q[..., 1::2]).astype(np.uint8)
    return packed, scale[..., 0].astype(np.float16), zero[..., 0].astype(np.float16)


def unpack_codes(packed: np.ndarray) -> np.ndarray:
    hi = (packed >> 4) & 0x0F
    lo = packed & 0x0F
    out = np.empty(packed.shape[:-1] + (packed.shape[-1] * 2,), dtype=np.uint8)
    out[..., 0::2] = hi
    out[..., 1::2] = lo
    return out


def dequantize_int4(packed: np.ndarray, scale: np.ndarray, zero: np.ndarray, group_size: int = 128, fused=True):
    """-> float16 ``[..., head_dim]``; fused multiply-add, one rounding (see the module docstring)."""
    codes = unpack_codes(packed)
    hd = codes.shape[-1]
    ng = hd // group_size
    c = codes.reshape(*codes.shape[:-1], ng, group_size).astype(np.float16)
    s = scale.astype(np.float16)[..., None]
    z = zero.astype(np.float16)[..., None]
    if fused:
        # exact in fp64 (<= 15 + 40 significant bits), then one correctly rounded conversion to fp16
        out = (c.astype(np.float64) * s.astype(np.float64) + z.astype(np.float64)).astype(np.float16)
    else:
        prod = (c * s).astype(np.float16)  # fp16 x fp16 is exact in fp32, then one RN to fp16
        out = (prod + z).astype(np.float16)
    return out.reshape(*codes.shape[:-1], hd)
