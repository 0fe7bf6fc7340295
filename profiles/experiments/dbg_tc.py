import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch
from duo_attention_b200 import _C
from duo_attention_b200.kv_cache import DuoKVCache
from oracle import duo_oracle as O
dev = torch.device("cuda:0"); D = 128
def stats(a, b):
    e = (a - b).abs(); tol = 1e-3 + 1e-2 * b.abs()
    return "max %.4f viol %.2e" % (e.max().item(), (e > tol).float().mean().item())
def case(Hq, Hkv, n_full, sink, recent, chunks, seed=0):
    g = torch.Generator().manual_seed(seed)
    tot = sum(chunks)
    a = DuoKVCache(1, Hq, Hkv, D, [n_full], 1, tot + 8, sink, recent, torch.bfloat16, dev, stage_cap=max(chunks))
    b = DuoKVCache(1, Hq, Hkv, D, [n_full], 1, tot + 8, sink, recent, torch.bfloat16, dev, stage_cap=max(chunks))
    past = None
    for S in chunks:
        qkv = torch.randn(1, S, (Hq + 2 * Hkv) * D, generator=g).to(torch.bfloat16)
        oa = torch.empty(1, S, Hq, D, dtype=torch.bfloat16, device=dev); ob = torch.empty_like(oa)
        a.attend(0, qkv.to(dev), None, None, _C.ROPE_NONE, oa)                    # dispatch (tc for S>=128)
        b.attend(0, qkv.to(dev), None, None, _C.ROPE_NONE, ob, force_mma=True)    # mma.sync family
        torch.cuda.synchronize()
        q = qkv[..., : Hq * D].reshape(1, S, Hq, D); k = qkv[..., Hq * D:(Hq + Hkv) * D].reshape(1, S, Hkv, D); v = qkv[..., (Hq + Hkv) * D:].reshape(1, S, Hkv, D)
        ref, past = O.tuple_attention_core(q, k, v, past, n_full, Hq // Hkv, sink, recent)
        print(f"  S={S}: tc-vs-oracle {stats(oa.float().cpu(), ref.float())} | mma-vs-oracle {stats(ob.float().cpu(), ref.float())} | nan={torch.isnan(oa.float()).any().item()}", flush=True)
print("case A: G=4, single 128 chunk, all full", flush=True)
case(4, 1, 1, 4, 4, [128])
print("case B: G=4 mix, 256 then 128 then 300", flush=True)
case(8, 2, 1, 16, 48, [256, 128, 1, 300, 1, 200])
print("case C: MHA", flush=True)
case(2, 2, 1, 8, 24, [384, 130, 1])
print("case D: deploy", flush=True)
case(32, 8, 3, 64, 256, [1000, 600, 1, 129])
