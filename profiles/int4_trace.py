"""Launch timeline of the INT4 decode kernel from per-CTA %globaltimer stamps (debug build: `make -C
duo_attention_b200/csrc trace`, loaded through DUO_B200_LIB).  Run under gpurun:
    DUO_B200_LIB=duo_attention_b200/csrc/_trace/libduo_b200_trace.so python profiles/int4_trace.py
Prints, per (n_full, ctx): kernel duration (CUDA events), and relative to the first CTA start: when the last CTA
started, when the main loops ended (median / max), when the last partial was published and when the merging CTA exited."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from duo_attention_b200 import _C  # noqa: E402
from duo_attention_b200.kv_cache import DuoKVCache  # noqa: E402

dev = torch.device("cuda:0")
lib = _C.load()
lib.duo_debug_set_trace.argtypes = [C.c_void_p]
lib.duo_debug_set_trace_mma.argtypes = [C.c_void_p]
Hq, Hkv, D, sink, recent = 32, 8, 128, 64, 256
for kvf, dtype in (("int4", torch.float16), ("same", torch.bfloat16)):
    for n_full, N in ((1, 1 << 20), (2, 1 << 20), (4, 1 << 20), (4, 1 << 17), (4, 1 << 16), (8, 1 << 20)):
        cache = DuoKVCache(1, Hq, Hkv, D, [n_full], 1, N + 8, sink, recent, dtype, dev, kv_format=kvf)
        t = cache.tensors[0]
        g = torch.Generator(device=dev).manual_seed(0)
        for name in ("full_k", "full_v", "ring_k", "ring_v"):
            if t[name].dtype == torch.uint8:
                t[name].random_(0, 256, generator=g)
                t[name + "_scale"].fill_(0.1)
                t[name + "_zero"].fill_(-0.8)
            elif t[name].numel():
                t[name].normal_(generator=g)
        qkv = torch.randn(1, 1, (Hq + 2 * Hkv) * D, device=dev, generator=g).to(dtype)
        out = torch.empty(1, 1, Hq, D, dtype=dtype, device=dev)
        st = _C.CacheState(N, N, N - recent)
        stream = torch.cuda.current_stream().cuda_stream
        h = cache.handles[0]
        _C.check(lib.duo_rope_append(h, C.byref(st), qkv.data_ptr(), qkv.stride(1), None, None, 0, 1, stream))
        trace = torch.zeros(4096 * 4, dtype=torch.int64, device=dev)

        def run():
            if kvf == "int4":
                _C.check(lib.duo_attention(h, C.byref(st), qkv.data_ptr(), qkv.stride(1), out.data_ptr(), 1, D ** -0.5,
                                           cache.workspace.data_ptr(), cache.workspace.numel(), stream))
            else:  # the product's decode path: one fused launch (RoPE off: q/k already rotated in this micro-benchmark)
                _C.check(lib.duo_decode_fused(h, C.byref(st), qkv.data_ptr(), qkv.stride(1), None, None, 0, out.data_ptr(),
                                              1, D ** -0.5, cache.workspace.data_ptr(), cache.workspace.numel(), stream))

        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        line = f"{kvf:5s} n_full={n_full} ctx={N:8d}: {ms*1e3:7.1f} us/launch"
        if True:
            setter = lib.duo_debug_set_trace if kvf == "int4" else lib.duo_debug_set_trace_mma
            setter(trace.data_ptr())
            trace.zero_()
            run()
            torch.cuda.synchronize()
            setter(None)
            tr = trace.view(-1, 4).cpu()
            tr = tr[tr[:, 0] > 0]
            t0 = tr[:, 0].min()
            rel = (tr - t0).float() / 1e3  # us
            merged = tr[:, 3] > 0
            line += (f" | CTAs {len(tr)}: last start {rel[:,0].max():.1f}, loop end med {rel[:,1].median():.1f} max "
                     f"{rel[:,1].max():.1f}, last publish {rel[:,2].max():.1f}, merge exit "
                     f"{rel[merged,3].max().item() if merged.any() else float('nan'):.1f} us")
        print(line, flush=True)
        del cache
        torch.cuda.empty_cache()
