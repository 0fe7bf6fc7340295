"""Isolated tcgen05 prefill launch: one Llama-3-8B layer (32 q / 8 kv heads, n_full = 4), a 32,768-token chunk over 98,304
cached tokens.  DUO_B200_LIB selects a tuning build (make -C duo_attention_b200/csrc variants); DUO_TC_REF=<file> saves the
output of the first run and compares later runs against it (profiles/tc_variants.sh)."""
import sys, os, ctypes as C
sys.path.insert(0, os.getcwd())
import torch
from duo_attention_b200 import _C
from duo_attention_b200.kv_cache import DuoKVCache
dev = torch.device("cuda:0"); D = 128
Hq, Hkv, nf, G = 32, 8, 4, 4
SINK, RECENT = 64, 256; W = SINK + RECENT
chunk, past = 32768, 98304
g = torch.Generator(device=dev).manual_seed(3)
cache = DuoKVCache(1, Hq, Hkv, D, [nf], 1, past + chunk + 8, SINK, RECENT, torch.bfloat16, dev, stage_cap=chunk)
for n in ("full_k", "full_v", "ring_k", "ring_v"): cache.tensors[0][n].normal_(generator=g)
qkv = torch.randn(1, chunk, (Hq + 2 * Hkv) * D, device=dev, dtype=torch.bfloat16, generator=g)
out = torch.empty(1, chunk, Hq, D, device=dev, dtype=torch.bfloat16)
st = _C.CacheState(past, past, past - RECENT)
lib, h = cache.lib, cache.handles[0]; stream = torch.cuda.current_stream().cuda_stream
_C.check(lib.duo_rope_append(h, C.byref(st), qkv.data_ptr(), qkv.stride(1), None, None, 0, chunk, stream))
def run():
    _C.check(lib.duo_attention(h, C.byref(st), qkv.data_ptr(), qkv.stride(1), out.data_ptr(), chunk, D ** -0.5, cache.workspace.data_ptr(), cache.workspace.numel(), stream))
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = int(os.environ.get("DUO_TC_REPS", "4"))
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
pf = chunk * past + chunk * (chunk + 1) // 2; ps = chunk * W + chunk * (chunk + 1) // 2
fl = 4.0 * D * G * (nf * pf + (Hkv - nf) * ps)
msg = f"{os.environ.get('DUO_B200_LIB', 'default')} {ms:.2f} ms  {fl/ms/1e9:.0f} TFLOP/s  checksum {out.float().abs().mean().item():.6f}"
ref = os.environ.get("DUO_TC_REF")
if ref:
    if not os.path.exists(ref):
        torch.save(out.cpu(), ref)
    else:
        base = torch.load(ref).to(dev).float()
        d = (out.float() - base).abs()
        bad = (d > 1e-3 + 1e-2 * base.abs()).sum().item()
        msg += f"  vs base: max|d| {d.max().item():.3e} mean|d| {d.mean().item():.3e} outside(1e-2,1e-3) {bad} of {d.numel()}"
print(msg)
