import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, numpy as np
from duo_attention_b200 import _C
from duo_attention_b200.kv_cache import DuoKVCache
from oracle import duo_oracle as O
dev = torch.device("cuda:0"); D = 128
torch.manual_seed(0)

N=131072
Hq, Hkv, n_full, sink, recent = 32, 8, 4, 64, 256
W=sink+recent
g = torch.Generator(device=dev).manual_seed(1)
# ---------- (B) noise floor: FA2 vs oracle, mine vs oracle, mine vs FA2
import flash_attn
def stats(a, b):
    e = (a - b).abs(); tol = 1e-3 + 1e-2 * b.abs()
    return "max %.4f viol %.2e" % (e.max().item(), (e > tol).float().mean().item())
for (Sq, Sk, qs) in [(1, 24, 1.0), (1, 700, 1.0), (64, 500, 1.0), (300, 300, 1.0), (700, 700, 6.0)]:
    gg = torch.Generator().manual_seed(Sq + Sk)
    Hq2, Hkv2 = 8, 2
    qkv2 = torch.randn(1, Sk, (Hq2 + 2 * Hkv2) * D, generator=gg).to(torch.bfloat16)
    qkv2[..., : Hq2 * D] *= qs
    q = qkv2[:, Sk - Sq:, : Hq2 * D].reshape(1, Sq, Hq2, D)
    k = qkv2[..., Hq2 * D: (Hq2 + Hkv2) * D].reshape(1, Sk, Hkv2, D)
    v = qkv2[..., (Hq2 + Hkv2) * D:].reshape(1, Sk, Hkv2, D)
    ref = O.flash_attn_contract(q, k, v).float()
    fa = flash_attn.flash_attn_func(q.cuda(), k.cuda(), v.cuda(), causal=True).float().cpu()
    c2 = DuoKVCache(1, Hq2, Hkv2, D, [Hkv2], 1, Sk + 8, 4, 4, torch.bfloat16, dev, stage_cap=max(1, Sk))
    if Sk > Sq:
        o0 = torch.empty(1, Sk - Sq, Hq2, D, dtype=torch.bfloat16, device=dev)
        c2.attend(0, qkv2[:, : Sk - Sq].to(dev).contiguous(), None, None, _C.ROPE_NONE, o0)
    o1 = torch.empty(1, Sq, Hq2, D, dtype=torch.bfloat16, device=dev)
    c2.attend(0, qkv2[:, Sk - Sq:].to(dev).contiguous(), None, None, _C.ROPE_NONE, o1)
    mine = o1.float().cpu()
    print(f"Sq={Sq} Sk={Sk} qs={qs}: FA2-vs-oracle {stats(fa, ref)} | mine-vs-oracle {stats(mine, ref)} | mine-vs-FA2 {stats(mine, fa)}")

# ---------- (C) decode kernel timing at 1M, n_f = 4
import ctypes as C
for N in (131072, 1048576):
    for nf in (1, 4, 8):
        cache = DuoKVCache(1, Hq, Hkv, D, [nf], 1, N + 8, sink, recent, torch.bfloat16, dev)
        for n in ("full_k", "ring_k", "full_v", "ring_v"):
            if cache.tensors[0][n].numel(): cache.tensors[0][n].normal_(generator=g)
        qkv = torch.randn(1, 1, (Hq + 2 * Hkv) * D, generator=g, device=dev).to(torch.bfloat16)
        out = torch.empty(1, 1, Hq, D, dtype=torch.bfloat16, device=dev)
        st = _C.CacheState(N, N, N - recent)
        lib = cache.lib; h = cache.handles[0]; stream = torch.cuda.current_stream().cuda_stream
        def run():
            _C.check(lib.duo_attention(h, C.byref(st), qkv.data_ptr(), qkv.stride(1), out.data_ptr(), 1, D ** -0.5, cache.workspace.data_ptr(), cache.workspace.numel(), stream))
        # need the new row appended first
        _C.check(lib.duo_rope_append(h, C.byref(st), qkv.data_ptr(), qkv.stride(1), None, None, 0, 1, stream))
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        by = (nf * (N + 1) + (Hkv - nf) * (W + 1)) * 2 * D * 2
        print(f"decode N={N} n_full={nf}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s  ({by/ms/1e6/6575.1*100:.1f}% of measured HBM peak)")
        del cache
