"""Head-major DuoAttention KV cache for B200.

Replaces the reference's token-major caches —

* ``DuoAttentionStaticKVCache``       (duo_attn/patch/static_kv_cache.py:18-315)
* the per-layer tuple cache           (duo_attn/patch/llama.py:168-171,292-301)
* ``DuoAttentionStaticINT4KVCache``   (demo/int4_kv.py:115-492)

— with one layout designed for coalesced 128 B HBM loads and TMA tiles:

    full_k / full_v : [B, n_full,   capacity,                 128]   retrieval heads
    ring_k / ring_v : [B, n_stream, sink + recent + stage_cap, 128]  streaming heads
                      slots [0,sink) sinks | [sink,sink+recent) true ring | staging for the chunk in flight

The streaming cache is a real ring (token ``p`` lives in slot ``sink + (p - sink) % recent``): the
reference's per-step compaction copies (static_kv_cache.py:127-167) become index arithmetic.  The Python
object only owns buffers and three integers per layer; all data movement and math is CUDA
(``csrc/``) reached through the C ABI.  Public methods keep the reference's names and semantics:
``kv_seq_len``, ``clear()``, ``evict_last(n)``, ``memory_usage``; overflow raises ``ValueError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _C


def _count_full(row) -> int:
    return int((np.asarray(row, dtype=np.float64) > 0.5).sum())


# ---- ring arithmetic (pure integers; the CUDA side implements the same formulas in duo_common.cuh) ----
def ring_slot(pos: int, sink: int, recent: int) -> int:
    """Slot of token ``pos`` in the streaming cache: sinks map to themselves, the rest into the ring."""
    return pos if pos < sink else sink + (pos - sink) % recent


def ring_advance(total: int, lo: int, n: int, sink: int, recent: int):
    """State after appending ``n`` tokens: the ring keeps the last ``recent`` positions."""
    total += n
    return total, max(lo, total - recent, sink)


def ring_evict(total: int, lo: int, n: int, sink: int):
    """State after ``evict_last(n)`` (static_kv_cache.py:290-297): the newest ``n`` tokens are dropped."""
    total = max(0, total - n)
    return total, max(sink, min(lo, total))


def ring_live_positions(total: int, lo: int, sink: int):
    """Token positions a streaming head can still see (what the reference's compacted cache holds)."""
    return list(range(0, min(total, sink))) + list(range(max(lo, sink), total))


class DuoKVCache:
    def __init__(
        self,
        num_layers: int,
        num_heads: int,
        num_kv_heads: int,
        head_dim: int,
        num_full_kv_head_list: Sequence[int],
        batch_size: int,
        max_size: int,
        sink_size: int,
        recent_size: int,
        dtype: torch.dtype,
        device,
        stage_cap: int = 64,
        kv_format: str = "same",
        growable: bool = False,
        local_full_cap: Optional[int] = None,
    ):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("DuoKVCache lives in GPU memory: the B200 kernels have no CPU fallback")
        if head_dim != 128:
            raise ValueError(f"head_dim {head_dim} not supported (the kernels are specialised for 128)")
        if dtype not in (torch.bfloat16, torch.float16):
            raise ValueError(f"dtype {dtype} not supported (bf16 / fp16)")
        if kv_format not in ("same", "int4"):
            raise ValueError(f"kv_format {kv_format!r} not supported")
        if kv_format == "int4" and dtype != torch.float16:
            raise ValueError("INT4 KV requires fp16 activations (demo/run_duo_w8a8kv4.py:41-45)")
        self.lib = _C.load()
        self.batch_size, self.max_size = int(batch_size), int(max_size)
        self.sink_size, self.recent_size = int(sink_size), int(recent_size)
        self.num_layers, self.num_heads, self.num_kv_heads = num_layers, num_heads, num_kv_heads
        self.num_kv_groups = num_heads // num_kv_heads
        self.head_dim = head_dim
        self.dtype, self.device = dtype, device
        self.kv_format = kv_format
        self.growable = growable
        self.num_full_kv_head_list = [int(n) for n in num_full_kv_head_list]
        self.num_streaming_kv_head_list = [num_kv_heads - n for n in self.num_full_kv_head_list]
        assert len(self.num_full_kv_head_list) == num_layers
        # occupancy, per layer (mirrors kv_seq_len_list / streaming_kv_seq_len_list of the reference)
        self.kv_seq_len_list = [0] * num_layers   # retrieval cache length
        self.total_list = [0] * num_layers        # tokens seen by the streaming heads
        self.lo_list = [self.sink_size] * num_layers
        # rows actually allocated per retrieval head (== max_size unless a subclass shards the positions over ranks)
        self.full_cap_list = [self.max_size if local_full_cap is None else int(local_full_cap)] * num_layers
        self.stage_cap_list = [max(1, int(stage_cap))] * num_layers
        self.tensors: List[dict] = []
        self.handles: List[Optional[int]] = [None] * num_layers
        for l in range(num_layers):
            self.tensors.append(self._alloc_layer(l, self.full_cap_list[l], self.stage_cap_list[l]))
            self._make_handle(l)
        self.graph_attached = False  # set by DuoDecodeGraph: buffers must then keep their addresses
        self.dev_state = None        # optional device copy of (full_len, total, lo): see enable_device_state()
        self.launch_count = 0        # kernels of this library enqueued through this cache
        self.profile_events = None   # set to [] to collect (start, end) CUDA events around every duo_attention
        ws = self.lib.duo_workspace_bytes(self.batch_size, num_kv_heads, self.num_kv_groups, _C.DECODE_MAX_Q)
        self.workspace = torch.zeros(ws, dtype=torch.uint8, device=device)

    # ------------------------------------------------------------------------------------------
    @property
    def W(self):
        return self.sink_size + self.recent_size

    @property
    def stage_off(self):
        """First staging slot (duo_b200.h): right after the ring, 64-aligned for INT4 caches."""
        return self.W if self.kv_format == "same" else (self.W + 63) // 64 * 64

    def _alloc_layer(self, l, full_cap, stage_cap, only=None):
        """Allocate layer ``l``'s buffers; ``only="ring"`` allocates just the streaming-head tensors."""
        nf, ns = self.num_full_kv_head_list[l], self.num_streaming_kv_head_list[l]
        B, D, dev = self.batch_size, self.head_dim, self.device
        slots = self.stage_off + stage_cap
        if self.kv_format == "int4":
            slots = (slots + 7) // 8 * 8
            full_cap = (full_cap + 63) // 64 * 64
        t = {}
        shapes = [(name, heads, rows) for name, heads, rows in (("full_k", nf, full_cap), ("full_v", nf, full_cap),
                                                                ("ring_k", ns, slots), ("ring_v", ns, slots))
                  if only is None or name.startswith(only)]
        if self.kv_format == "same":
            for name, heads, rows in shapes:
                t[name] = torch.zeros(B, heads, rows, D, dtype=self.dtype, device=dev)
        else:
            for name, heads, rows in shapes:
                t[name] = torch.zeros(B, heads, rows, D // 2, dtype=torch.uint8, device=dev)
                t[name + "_scale"] = torch.zeros(B, heads, rows, dtype=torch.float16, device=dev)
                t[name + "_zero"] = torch.zeros(B, heads, rows, dtype=torch.float16, device=dev)
        return t

    def _make_handle(self, l):
        if self.handles[l] is not None:
            self.lib.duo_layer_destroy(self.handles[l])
            self.handles[l] = None
        t = self.tensors[l]
        d = _C.LayerDesc()
        for name in ("full_k", "full_v", "ring_k", "ring_v"):
            setattr(d, name, t[name].data_ptr() if t[name].numel() else None)
            for suf in ("_scale", "_zero"):
                key = name + suf
                setattr(d, key, t[key].data_ptr() if key in t and t[key].numel() else None)
        d.full_cap = t["full_k"].shape[2]
        d.batch = self.batch_size
        d.n_full = self.num_full_kv_head_list[l]
        d.n_stream = self.num_streaming_kv_head_list[l]
        d.group = self.num_kv_groups
        d.head_dim = self.head_dim
        d.sink, d.recent = self.sink_size, self.recent_size
        d.stage_cap = t["ring_k"].shape[2] - self.stage_off
        d.dtype = _C.DT_BF16 if self.dtype == torch.bfloat16 else _C.DT_FP16
        d.kv_format = _C.KV_SAME if self.kv_format == "same" else _C.KV_INT4
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _C.check(self.lib.duo_layer_create(C.byref(d), C.byref(h)))
        self.handles[l] = h.value

    def __del__(self):
        try:
            for h in self.handles:
                if h is not None:
                    self.lib.duo_layer_destroy(h)
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def _ensure_room(self, l, q_len):
        """Grow the staging area (always allowed) and, for growable caches (tuple-path compatibility,
        where the reference simply torch.cat's), the retrieval cache."""
        need_full = self._rows_needed(l, q_len)
        grow_full = need_full > self.full_cap_list[l]
        if grow_full and not self.growable:
            # same check and message as static_kv_cache.py:112-115
            raise ValueError(
                f"Trying to put {q_len} KVs into a cache with max size {self.max_size}, "
                f"current size: {self.kv_seq_len_list[l]}."
            )
        grow_stage = q_len > self.stage_cap_list[l]
        if not (grow_full or grow_stage):
            return
        if self.graph_attached:
            raise ValueError("this cache is captured in a DuoDecodeGraph: its buffers cannot be re-allocated "
                             f"(chunk of {q_len} tokens > staging capacity {self.stage_cap_list[l]})")
        new_full = self.full_cap_list[l]
        if need_full > new_full:
            new_full = max(need_full, 2 * new_full, 256)
        new_stage = max(self.stage_cap_list[l], q_len)
        old = self.tensors[l]
        # only what must grow is re-allocated: a longer staging area leaves the (possibly multi-GB) retrieval cache alone
        new = self._alloc_layer(l, new_full, new_stage, only=None if grow_full else "ring")
        n = self.kv_seq_len_list[l]
        for name, t in old.items():
            if name not in new:
                new[name] = t
            elif name.startswith("full"):
                new[name][:, :, :n].copy_(t[:, :, :n])
            else:
                new[name][:, :, : self.W].copy_(t[:, :, : self.W])
        self.tensors[l] = new
        self.full_cap_list[l] = new_full
        self.stage_cap_list[l] = new_stage
        self._make_handle(l)

    def _first_chunk_scratch(self, S):
        sc = getattr(self, "_scratch", None)
        if sc is None or sc.max_size < S:
            sc = DuoKVCache(1, self.num_heads, self.num_kv_heads, self.head_dim, [self.num_kv_heads],
                            self.batch_size, max(S, 64), self.sink_size, self.recent_size, self.dtype, self.device,
                            stage_cap=max(S, 64), kv_format="same")  # no streaming heads: staging costs nothing
            self._scratch = sc
        return sc

    # ---- large chunks (>= 128 tokens) over an INT4 cache: tcgen05 prefill kernel on an fp16 image ----------------
    def _dequant_scratch(self, l, S):
        """fp16 image of layer ``l``'s INT4 cache for ONE attention call of a chunk of >= 128 tokens — what the
        reference does on EVERY call (``get()`` dequantises the whole cache, demo/int4_kv.py:373-436, then
        flash_attn_func runs on it, demo/w8a8kv4_llama.py:239-274).  For such a chunk the O(ctx) dequantisation pass
        is < 1 % of the chunk x ctx attention, and the attention runs on the tensor-core prefill kernel (measured
        on the box: INT4 128K prefill at the bf16 speed, profiles/r2_validation.md).  Decode and small chunks never
        come here: their kernels dequantise in the K/V load stage.  One flat zero-initialised fp16 buffer is shared
        by all layers (they are processed one after the other; rows beyond the dequantised range hold zeros or
        finite leftovers and are masked); a layer handle is created per distinct number of retrieval heads."""
        sc = self.__dict__.get("_dq")
        B, D, Hkv = self.batch_size, self.head_dim, self.num_kv_heads
        cap = max(self.full_cap_list)
        slots = self.W + max(max(self.stage_cap_list), S)
        if sc is None or sc["cap"] < cap or sc["slots"] < slots:
            nf_max = max(self.num_full_kv_head_list)
            ns_max = max(self.num_streaming_kv_head_list)
            sc = {"cap": cap, "slots": slots, "handles": {},
                  "full": [torch.zeros(B * nf_max * cap * D, dtype=self.dtype, device=self.device) for _ in range(2)],
                  "ring": [torch.zeros(B * ns_max * slots * D, dtype=self.dtype, device=self.device) for _ in range(2)]}
            old = self.__dict__.get("_dq")
            if old is not None:
                for hd in old["handles"].values():
                    self.lib.duo_layer_destroy(hd)
            self._dq = sc
        nf, ns = self.num_full_kv_head_list[l], self.num_streaming_kv_head_list[l]
        fk, fv = (t[: B * nf * cap * D].view(B, nf, cap, D) for t in sc["full"])
        rk, rv = (t[: B * ns * slots * D].view(B, ns, slots, D) for t in sc["ring"])
        if nf not in sc["handles"]:
            d = _C.LayerDesc()
            d.full_k, d.full_v = (fk.data_ptr(), fv.data_ptr()) if nf else (None, None)
            d.ring_k, d.ring_v = (rk.data_ptr(), rv.data_ptr()) if ns else (None, None)
            d.full_cap, d.batch, d.n_full, d.n_stream, d.group, d.head_dim = cap, B, nf, ns, self.num_kv_groups, D
            d.sink, d.recent, d.stage_cap = self.sink_size, self.recent_size, slots - self.W
            d.dtype = _C.DT_BF16 if self.dtype == torch.bfloat16 else _C.DT_FP16
            d.kv_format = _C.KV_SAME
            hd = C.c_void_p()
            with torch.cuda.device(self.device):
                _C.check(self.lib.duo_layer_create(C.byref(d), C.byref(hd)))
            sc["handles"][nf] = hd.value
        # dequantise what this call can see: retrieval rows [0, full_len + S), sink + ring slots, the staged chunk
        t = self.tensors[l]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        n_rows = self.kv_seq_len_list[l] + S
        W, so = self.W, self.stage_off
        n = 0
        for b in range(B):
            for hh in range(nf):
                for name, dst in (("full_k", fk), ("full_v", fv)):
                    _C.check(self.lib.duo_dequant_int4(t[name][b, hh].data_ptr(), t[name + "_scale"][b, hh].data_ptr(),
                                                       t[name + "_zero"][b, hh].data_ptr(), n_rows,
                                                       dst[b, hh].data_ptr(), stream))
                    n += 1
            for hh in range(ns):
                for name, dst in (("ring_k", rk), ("ring_v", rv)):
                    for src0, dst0, rows in ((0, 0, W), (so, W, S)):
                        _C.check(self.lib.duo_dequant_int4(
                            t[name][b, hh, src0:].data_ptr(), t[name + "_scale"][b, hh, src0:].data_ptr(),
                            t[name + "_zero"][b, hh, src0:].data_ptr(), rows, dst[b, hh, dst0:].data_ptr(), stream))
                        n += 1
        self.launch_count += n
        return sc["handles"][nf]

    def state(self, l) -> _C.CacheState:
        ds = self.dev_state.data_ptr() if self.dev_state is not None else None
        return _C.CacheState(self.kv_seq_len_list[l], self.total_list[l], self.lo_list[l], ds)

    def _rows_needed(self, l, q_len) -> int:
        """Rows of the (local) retrieval cache in use after appending ``q_len`` tokens."""
        return self.kv_seq_len_list[l] + q_len

    # ---- device-resident occupancy (CUDA-graph replay of decode steps) ---------------------------------
    def enable_device_state(self):
        """Keep a device copy of (full_len, total, lo) that the decode kernels read at launch, so that a captured
        decode step can be replayed while the context grows.  All layers hold the same occupancy at step
        boundaries, so one copy serves the whole cache."""
        if self.dev_state is None:
            self.dev_state = torch.zeros(4, dtype=torch.int64, device=self.device)
        self.sync_device_state()
        return self

    def sync_device_state(self):
        """Stream-ordered refresh of the device copy: the integers travel as kernel arguments (duo_state_set), so
        back-to-back evict_last()/clear() calls cannot race through a shared staging buffer."""
        if self.dev_state is None:
            return
        l = self.num_layers - 1
        _C.check(self.lib.duo_state_set(self.dev_state.data_ptr(), self.kv_seq_len_list[l], self.total_list[l],
                                        self.lo_list[l], torch.cuda.current_stream(self.device).cuda_stream))
        self.launch_count += 1

    def advance_device(self, n):
        """Enqueue full_len += n, total += n, lo = max(lo, total - recent, sink) on the device copy."""
        _C.check(self.lib.duo_state_advance(self.dev_state.data_ptr(), int(n), self.sink_size, self.recent_size,
                                            torch.cuda.current_stream(self.device).cuda_stream))
        self.launch_count += 1

    def snapshot_ring(self):
        """Copy of the sink+ring slots of every layer (a few hundred KB each): lets a caller run throw-away steps
        (CUDA-graph warm-up / capture) and put the streaming cache back exactly as it was."""
        W = self.W
        return [{k: v[:, :, :W].clone() for k, v in t.items() if k.startswith("ring")} for t in self.tensors]

    def restore_ring(self, snap):
        W = self.W
        for t, sn in zip(self.tensors, snap):
            for k, v in sn.items():
                t[k][:, :, :W].copy_(v)

    def advance_host(self, n):
        """Mirror of advance_device for the host integers (call once per graph replay)."""
        for l in range(self.num_layers):
            self.advance(l, n)

    def advance(self, l, q_len):
        self.kv_seq_len_list[l] += q_len
        self.total_list[l], self.lo_list[l] = ring_advance(self.total_list[l], self.lo_list[l], q_len,
                                                           self.sink_size, self.recent_size)

    # ------------------------------------------------------------------------------------------
    # reference-compatible surface (static_kv_cache.py:100-107, 285-315)
    @property
    def kv_seq_len(self):
        return self.kv_seq_len_list[-1]

    @property
    def streaming_kv_seq_len(self):
        """Rows the reference's compacted streaming cache would hold (sinks + live ring entries)."""
        tot, lo = self.total_list[-1], self.lo_list[-1]
        return min(tot, self.sink_size) + max(0, tot - max(lo, self.sink_size))

    def clear(self):
        for l in range(self.num_layers):
            self.kv_seq_len_list[l] = 0
            self.total_list[l] = 0
            self.lo_list[l] = self.sink_size
        self.sync_device_state()

    def evict_last(self, num_tokens):
        for l in range(self.num_layers):
            self.kv_seq_len_list[l] = max(0, self.kv_seq_len_list[l] - num_tokens)
            self.total_list[l], self.lo_list[l] = ring_evict(self.total_list[l], self.lo_list[l], num_tokens,
                                                             self.sink_size)
        self.sync_device_state()

    @property
    def memory_usage(self):
        tot = 0
        for t in self.tensors:
            for v in t.values():
                tot += v.element_size() * v.numel()
        return tot

    # ------------------------------------------------------------------------------------------
    def attend(self, l, qkv, cos, sin, rope_mode, out, scale=None, force_mma=False, fused=True):
        """The fused per-layer hot path: RoPE + append, mixed-head attention, ring commit.

        qkv  ``[B, S, (Hq + 2 Hkv) * D]`` (last dim contiguous) — q is rotated in place on the three-launch path,
             left untouched on the one-launch path.
        out  ``[B, S, Hq, D]`` contiguous, written.
        ``fused=False`` forces the three-launch path for decode-sized chunks (tests: both paths produce the same bits).
        """
        if not qkv.is_cuda or not out.is_cuda:
            raise RuntimeError("duo_attention_b200 kernels need CUDA tensors (no CPU fallback)")
        B, S, width = qkv.shape
        assert B == self.batch_size and width == (self.num_heads + 2 * self.num_kv_heads) * self.head_dim
        assert qkv.stride(2) == 1 and (B == 1 or qkv.stride(0) == S * qkv.stride(1)), "qkv rows must be uniformly strided"
        assert out.is_contiguous() and qkv.dtype == self.dtype and out.dtype == self.dtype
        self._ensure_room(l, S)
        st = self.state(l)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        h = self.handles[l]
        lib = self.lib
        if scale is None:
            scale = self.head_dim ** -0.5
        cp = cos.data_ptr() if cos is not None else None
        sp = sin.data_ptr() if sin is not None else None
        # decode-sized chunks take ONE launch: 16-bit caches up to 16 packed rows, INT4 caches up to 8 (the keys-as-M
        # kernel) — except the very first INT4 call, which attends the raw fp16 K/V (see below)
        one_launch = (S * self.num_kv_groups <= _C.DECODE_MAX_Q if self.kv_format == "same" else
                      S * self.num_kv_groups <= _C.DECODE_MAX_Q_INT4 and not (st.full_len == 0 and st.total == 0))
        if (one_launch and fused and not force_mma and qkv.stride(1) % 8 == 0 and qkv.data_ptr() % 16 == 0):
            # decode-sized chunk: RoPE + append + attention + ring commit in ONE launch (q is not written back)
            if self.profile_events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            _C.check(lib.duo_decode_fused(h, C.byref(st), qkv.data_ptr(), qkv.stride(1), cp, sp, rope_mode & 0xFF,
                                          out.data_ptr(), S, float(scale), self.workspace.data_ptr(),
                                          self.workspace.numel(), stream))
            if self.profile_events is not None:
                e1.record()
                self.profile_events.append((e0, e1))
            self.launch_count += 1
            self.advance(l, S)
            return out
        _C.check(lib.duo_rope_append(h, C.byref(st), qkv.data_ptr(), qkv.stride(1), cp, sp, rope_mode, S, stream))
        ah, ast = h, st
        if self.kv_format == "int4" and st.full_len == 0 and st.total == 0:
            # The reference attends the RAW fp16 K/V on the very first call and only later calls see the
            # quantise->dequantise round trip (demo/w8a8kv4_llama.py:229-238 vs :239-274).  The chunk has just
            # been quantised into the INT4 cache above; attention for this one call runs on an fp16 scratch layer
            # (every head is plain causal on the first call, so all heads are "retrieval" there).
            sc = self._first_chunk_scratch(S)
            ah, ast = sc.handles[0], _C.CacheState(0, 0, sc.sink_size)
            _C.check(lib.duo_rope_append(ah, C.byref(ast), qkv.data_ptr(), qkv.stride(1), cp, sp,
                                         rope_mode | _C.ROPE_SKIP_Q, S, stream))
            self.launch_count += 1
        elif self.kv_format == "int4" and S >= 128 and self.W <= 2048 and not force_mma:
            ah, ast = self._dequant_scratch(l, S), _C.CacheState(st.full_len, st.total, st.lo, None)
        fn = lib.duo_attention_mma if force_mma else lib.duo_attention
        if self.profile_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _C.check(fn(ah, C.byref(ast), qkv.data_ptr(), qkv.stride(1), out.data_ptr(), S, float(scale),
                    self.workspace.data_ptr(), self.workspace.numel(), stream))
        if self.profile_events is not None:
            e1.record()
            self.profile_events.append((e0, e1))
        _C.check(lib.duo_stream_commit(h, C.byref(st), S, stream))
        self.launch_count += 2 + (1 if self.num_streaming_kv_head_list[l] > 0 else 0)
        self.advance(l, S)
        return out


class DuoAttentionStaticKVCache(DuoKVCache):
    """Drop-in for the reference class of the same name (static_kv_cache.py:18-98): same constructor
    arguments (plus optional keyword extras), same ``clear`` / ``evict_last`` / ``memory_usage`` /
    ``kv_seq_len`` surface, head-major storage underneath."""

    def __init__(self, model, full_attention_heads, batch_size, max_size, sink_size, recent_size,
                 prefilling_chunk_size: int = 64, kv_format: str = "same"):
        p = next(model.parameters())
        cfg = model.config
        head_dim = getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads
        super().__init__(
            num_layers=cfg.num_hidden_layers,
            num_heads=cfg.num_attention_heads,
            num_kv_heads=cfg.num_key_value_heads,
            head_dim=head_dim,
            num_full_kv_head_list=[_count_full(r) for r in full_attention_heads],
            batch_size=batch_size,
            max_size=max_size,
            sink_size=sink_size,
            recent_size=recent_size,
            dtype=p.dtype,
            device=p.device,
            stage_cap=prefilling_chunk_size,
            kv_format=kv_format,
            growable=False,
        )


class DuoSeqShardKVCache(DuoKVCache):
    """Decode-phase cache of the SEQUENCE-SHARDED tensor-parallel layout (scope row f1; ``tp.install_seq_shard``):
    every rank holds all heads of a layer, but of each retrieval head only its block-cyclic slice of the token
    positions (``seqshard.SeqShardPlan``); streaming heads are replicated.  ``kv_seq_len`` keeps counting GLOBAL tokens.

    ``attend`` = duo_rope_append (appends only positions this rank owns) -> duo_attention_seq (slice partials for the
    retrieval heads, streaming heads final) -> duo_seq_merge (peer-memory exchange + merge) -> duo_stream_commit.
    Only decode-sized chunks (group x q_len <= 16): prefill runs head-parallel and ``tp.reshard_heads_to_seq`` moves
    the caches over (``load_from_head_parallel``).  Same ``clear`` / ``evict_last`` / ``memory_usage`` surface as the
    static cache; ``DuoDecodeGraph`` can capture it."""

    def __init__(self, model, full_attention_heads, batch_size, max_size, sink_size, recent_size, seq=None):
        from .seqshard import SeqShardPlan

        seq = seq if seq is not None else model._duo_seq
        self.seq = seq
        self.plan = SeqShardPlan(seq.world, seq.block)
        p = next(model.parameters())
        cfg = model.config
        head_dim = getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads
        super().__init__(
            num_layers=cfg.num_hidden_layers, num_heads=cfg.num_attention_heads, num_kv_heads=cfg.num_key_value_heads,
            head_dim=head_dim, num_full_kv_head_list=[_count_full(r) for r in full_attention_heads],
            batch_size=batch_size, max_size=max_size, sink_size=sink_size, recent_size=recent_size, dtype=p.dtype,
            device=p.device, stage_cap=_C.DECODE_MAX_Q, kv_format="same", growable=False,
            local_full_cap=self.plan.capacity(int(max_size)) + 1)
        q_max = _C.DECODE_MAX_Q // self.num_kv_groups
        if q_max < 1 or batch_size * q_max * self.num_heads > seq.comm.max_rows * 8:
            pass  # the per-call row check in duo_seq_merge is authoritative
        self.max_q = max(1, q_max)
        self.part_o = torch.zeros(batch_size, self.max_q, self.num_heads, head_dim, dtype=torch.float32, device=p.device)
        self.part_lse = torch.zeros(batch_size, self.max_q, self.num_heads, dtype=torch.float32, device=p.device)

    def state(self, l) -> _C.CacheState:
        st = super().state(l)
        st.seq_rank, st.seq_world, st.seq_block = self.seq.rank, self.seq.world, self.seq.block
        return st

    def _rows_needed(self, l, q_len) -> int:
        return self.plan.local_len(self.seq.rank, self.kv_seq_len_list[l] + q_len)

    def attend(self, l, qkv, cos, sin, rope_mode, out, scale=None, force_mma=False):
        if not qkv.is_cuda or not out.is_cuda:
            raise RuntimeError("duo_attention_b200 kernels need CUDA tensors (no CPU fallback)")
        B, S, width = qkv.shape
        if S > self.max_q:
            raise ValueError(f"sequence-sharded caches serve decode-sized chunks (<= {self.max_q} tokens, got {S}): "
                             "prefill head-parallel and move the caches over with load_from_head_parallel()")
        assert B == self.batch_size and width == (self.num_heads + 2 * self.num_kv_heads) * self.head_dim
        assert qkv.stride(2) == 1 and (B == 1 or qkv.stride(0) == S * qkv.stride(1))
        assert out.is_contiguous() and qkv.dtype == self.dtype and out.dtype == self.dtype
        self._ensure_room(l, S)
        st = self.state(l)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        h, lib = self.handles[l], self.lib
        if scale is None:
            scale = self.head_dim ** -0.5
        cp = cos.data_ptr() if cos is not None else None
        sp = sin.data_ptr() if sin is not None else None
        nfq = self.num_full_kv_head_list[l] * self.num_kv_groups
        po, pl = self.part_o[:, :S], self.part_lse[:, :S]
        fused = S == 1 and qkv.stride(1) % 8 == 0 and qkv.data_ptr() % 16 == 0
        if not fused:
            _C.check(lib.duo_rope_append(h, C.byref(st), qkv.data_ptr(), qkv.stride(1), cp, sp, rope_mode, S, stream))
        if S != self.max_q:  # the kernels index [batch][q_len][heads]: contiguous views of the right q_len
            po = self.part_o.view(-1)[: B * S * self.num_heads * self.head_dim].view(B, S, self.num_heads, self.head_dim)
            pl = self.part_lse.view(-1)[: B * S * self.num_heads].view(B, S, self.num_heads)
        if self.profile_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if fused:  # one token: RoPE + owner-only append + slice attention + streaming heads + ring commit, one launch
            _C.check(lib.duo_decode_fused_seq(h, C.byref(st), qkv.data_ptr(), qkv.stride(1), cp, sp, rope_mode & 0xFF,
                                              out.data_ptr(), po.data_ptr(), pl.data_ptr(), float(scale),
                                              self.workspace.data_ptr(), self.workspace.numel(), stream))
        else:
            _C.check(lib.duo_attention_seq(h, C.byref(st), qkv.data_ptr(), qkv.stride(1), out.data_ptr(), po.data_ptr(),
                                           pl.data_ptr(), S, float(scale), self.workspace.data_ptr(),
                                           self.workspace.numel(), stream))
        if self.profile_events is not None:
            e1.record()
            self.profile_events.append((e0, e1))
        if nfq:
            self.seq.comm.merge(po, pl, out, B * S, self.num_heads, nfq)
        if fused:
            self.launch_count += 1 + (1 if nfq else 0)
        else:
            _C.check(lib.duo_stream_commit(h, C.byref(st), S, stream))
            self.launch_count += 2 + (1 if nfq else 0) + (1 if self.num_streaming_kv_head_list[l] > 0 else 0)
        self.advance(l, S)
        return out

    def load_from_head_parallel(self, hp_cache: "DuoKVCache", head_plan, group=None):
        """Take over the contents of a head-parallel cache (this rank's heads, all positions — what the prefill
        phase filled) by point-to-point resharding; streaming rings are all-gathered (they are tiny)."""
        import torch.distributed as dist

        from . import tp

        rank, world = self.seq.rank, self.seq.world
        W = self.W
        for l in range(self.num_layers):
            n = hp_cache.kv_seq_len_list[l]
            mask_row = head_plan.mask[l]
            owners = head_plan.owners[l]
            for name in ("full_k", "full_v"):
                if self.tensors[l][name].numel():
                    tp.reshard_heads_to_seq(hp_cache.tensors[l][name], owners, mask_row, rank, world, n, self.seq.block,
                                            self.tensors[l][name], group)
            stream_ids = [h for h in range(len(mask_row)) if mask_row[h] <= 0.5]
            sidx = {h: i for i, h in enumerate(stream_ids)}
            for name in ("ring_k", "ring_v"):
                if not self.tensors[l][name].numel():
                    continue
                n_loc = max(len([h for h in owners[r] if mask_row[h] <= 0.5]) for r in range(world))
                mine = [h for h in owners[rank] if mask_row[h] <= 0.5]
                send = torch.zeros((self.batch_size, n_loc, W, self.head_dim), dtype=self.dtype, device=self.device)
                if mine:
                    send[:, : len(mine)] = hp_cache.tensors[l][name][:, :, :W]
                got = [torch.empty_like(send) for _ in range(world)]
                dist.all_gather(got, send, group=group)
                for r in range(world):
                    theirs = [h for h in owners[r] if mask_row[h] <= 0.5]
                    for i, h in enumerate(theirs):
                        self.tensors[l][name][:, sidx[h], :W] = got[r][:, i]
            self.kv_seq_len_list[l] = n
            self.total_list[l] = hp_cache.total_list[l]
            self.lo_list[l] = hp_cache.lo_list[l]
        self.sync_device_state()
        return self


class DuoAttentionStaticINT4KVCache(DuoAttentionStaticKVCache):
    """Drop-in for the demo's INT4 cache (demo/int4_kv.py:115-260): same constructor arguments.  Storage is the
    packed-nibble + fp16 scale/zero format of demo/quantize_int4.cu in head-major order; there is no fp16 scratch
    copy of the cache and no per-step ``get()`` dequantisation pass — the attention kernel dequantises in its
    K/V load stage.  The model must run in fp16, as in the demo (run_duo_w8a8kv4.py:41-45)."""

    def __init__(self, model, full_attention_heads, batch_size, max_size, sink_size, recent_size,
                 prefilling_chunk_size):
        super().__init__(model, full_attention_heads, batch_size, max_size, sink_size, recent_size,
                         prefilling_chunk_size=prefilling_chunk_size, kv_format="int4")
