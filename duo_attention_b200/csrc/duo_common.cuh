// Shared device helpers (sm_100a): mbarrier, TMA, ldmatrix, mma.sync wrappers + host-side
// error plumbing.  Everything here is hand-written PTX; no CUTLASS/CuTe.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/duo_b200.h"

namespace duo {

constexpr int kHeadDim = 128;

// ---------------------------------------------------------------------------------------------
// host-side error handling
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define DUO_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return ::duo::cuda_fail(_e, #expr); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// the opaque per-layer handle
// ---------------------------------------------------------------------------------------------
struct LayerMaps {
  // 3-D maps {head_dim, slots, batch*heads}; box {64 elems, tile rows, 1}; SWIZZLE_128B
  CUtensorMap full_k64, full_v64;    // 64-row boxes (mma.sync bandwidth kernel)
  CUtensorMap ring_k64, ring_v64;
  CUtensorMap full_k128, full_v128;  // 128-row boxes (tcgen05 prefill kernel)
  CUtensorMap ring_k128, ring_v128;
};

}  // namespace duo

struct duo_layer {
  duo_layer_desc d;
  duo::LayerMaps maps;
  bool has_full_maps;
  bool has_ring_maps;
};

namespace duo {

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a lost TMA transaction / missing arrival must not hang the GPU.  After 2^26 failed polls (a failed
// try_wait suspends the thread for an implementation-defined interval first, so this is seconds to a minute) the
// kernel traps, which surfaces as a CUDA launch failure (DUO_ECUDA at the next API call) instead of a wedged device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++polls == (1u << 26)) asm volatile("trap;");
  }
}
// Unbounded spin for register-starved consumers whose producer side is already bounded (a trap anywhere ends the grid).
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// TMA tiled 3-D load global -> shared, completion on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                 int c2, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, "
      "%4, %5}], [%2], %6;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

template <typename T>
struct MmaOp;
template <>
struct MmaOp<__nv_bfloat16> {
  __device__ __forceinline__ static void run(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  __device__ __forceinline__ static uint32_t pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};
template <>
struct MmaOp<__half> {
  __device__ __forceinline__ static void run(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  __device__ __forceinline__ static uint32_t pack(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};

// ---------------------------------------------------------------------------------------------
// RoPE element arithmetic, shared by rope_append_kernel (kv_ops.cu) and the fused decode kernel (attn_mma.cu) so that
// both produce the same bits.  `rot` is the rotate_half partner with its sign (exact in T).
//   HF  : T(T(x*cos) + T(rot*sin))   — transformers' apply_rotary_pos_emb on T tensors (llama.py:177-184)
//   fp32: T(x*cos + rot*sin), fp32 tables, one rounding (flashinfer semantics, flashinfer_utils.py:29-59)
// ---------------------------------------------------------------------------------------------
template <typename T>
struct RopeCvt;
template <>
struct RopeCvt<__nv_bfloat16> {
  __device__ static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <>
struct RopeCvt<__half> {
  __device__ static float to_f(__half v) { return __half2float(v); }
  __device__ static __half from_f(float v) { return __float2half_rn(v); }
};
template <typename T>
__device__ __forceinline__ float rope_hf(float x, float rot, float c, float s) {
  const float a = RopeCvt<T>::to_f(RopeCvt<T>::from_f(__fmul_rn(x, c)));
  const float r = RopeCvt<T>::to_f(RopeCvt<T>::from_f(__fmul_rn(rot, s)));
  return __fadd_rn(a, r);  // the caller's conversion to T is the third rounding
}
__device__ __forceinline__ float rope_f32(float x, float rot, float c, float s) {
  return __fmaf_rn(rot, s, __fmul_rn(x, c));
}

// RoPE of 8 consecutive head_dim elements d .. d+7 (d < 64) and their partners d+64 .. d+71, arithmetic of
// rope_append_kernel (kv_ops.cu): HF mode rounds every product and the sum to T, fp32 mode rounds once.
template <typename T>
__device__ __forceinline__ void rope8(uint4& lo, uint4& hi, const void* cos, const void* sin, int mode, int tok, int d) {
  T* xl = reinterpret_cast<T*>(&lo);
  T* xh = reinterpret_cast<T*>(&hi);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float a = RopeCvt<T>::to_f(xl[e]), bb = RopeCvt<T>::to_f(xh[e]);
    float ol, oh;
    if (mode == DUO_ROPE_HF) {
      const T* ct = reinterpret_cast<const T*>(cos) + (long long)tok * kHeadDim;
      const T* st = reinterpret_cast<const T*>(sin) + (long long)tok * kHeadDim;
      ol = rope_hf<T>(a, -bb, RopeCvt<T>::to_f(ct[d + e]), RopeCvt<T>::to_f(st[d + e]));
      oh = rope_hf<T>(bb, a, RopeCvt<T>::to_f(ct[d + 64 + e]), RopeCvt<T>::to_f(st[d + 64 + e]));
    } else {
      const float* ct = reinterpret_cast<const float*>(cos) + (long long)tok * kHeadDim;
      const float* st = reinterpret_cast<const float*>(sin) + (long long)tok * kHeadDim;
      ol = rope_f32(a, -bb, ct[d + e], st[d + e]);
      oh = rope_f32(bb, a, ct[d + 64 + e], st[d + 64 + e]);
    }
    xl[e] = RopeCvt<T>::from_f(ol);
    xh[e] = RopeCvt<T>::from_f(oh);
  }
}

// ---------------------------------------------------------------------------------------------
// One 128-element row held 4-per-lane by a warp (lane l: elements 4l .. 4l+3): RoPE and INT4 K1 quantisation.
// Shared by rope_append_kernel (kv_ops.cu) and the fused INT4 decode kernel (attn_int4.cu): same bits from both.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct alignas(8) Vec4 {
  T v[4];
};

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// RoPE of token `t` (row t of the cos / sin tables of this chunk) on a warp-held row; every lane must call.
template <typename T>
__device__ __forceinline__ void rope_row4(Vec4<T>& xv, int lane, int t, const void* cos, const void* sin, int mode) {
  // rotate_half partner: element i pairs with i +- 64  <=> lane +- 16
  Vec4<T> pv;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float mine = RopeCvt<T>::to_f(xv.v[i]);
    const float other = __shfl_xor_sync(0xffffffffu, mine, 16);
    pv.v[i] = RopeCvt<T>::from_f(lane < 16 ? -other : other);
  }
  if (mode == DUO_ROPE_HF) {
    const Vec4<T> cv = *reinterpret_cast<const Vec4<T>*>(reinterpret_cast<const T*>(cos) + (long long)t * kHeadDim + lane * 4);
    const Vec4<T> sv = *reinterpret_cast<const Vec4<T>*>(reinterpret_cast<const T*>(sin) + (long long)t * kHeadDim + lane * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      xv.v[i] = RopeCvt<T>::from_f(rope_hf<T>(RopeCvt<T>::to_f(xv.v[i]), RopeCvt<T>::to_f(pv.v[i]), RopeCvt<T>::to_f(cv.v[i]),
                                              RopeCvt<T>::to_f(sv.v[i])));
  } else {
    const float4 cv = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(cos) + (long long)t * kHeadDim + lane * 4);
    const float4 sv = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sin) + (long long)t * kHeadDim + lane * 4);
    const float c[4] = {cv.x, cv.y, cv.z, cv.w};
    const float s[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      xv.v[i] = RopeCvt<T>::from_f(rope_f32(RopeCvt<T>::to_f(xv.v[i]), RopeCvt<T>::to_f(pv.v[i]), c[i], s[i]));
  }
}

// K1 arithmetic (demo/quantize_int4.cu:73-144) for one 128-element group held 4-per-lane (fp16-representable values
// in x[]).  Writes 2 packed bytes per lane; lane 0 writes scale / zero.
__device__ __forceinline__ void quant_row_int4(const float (&x)[4], int lane, uint8_t* packed_row, __half* scale_p,
                                               __half* zero_p) {
  float mn = fminf(fminf(x[0], x[1]), fminf(x[2], x[3]));
  float mx = fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3]));
  mn = warp_min(mn);
  mx = warp_max(mx);
  const float scale = __fadd_rn(__fdiv_rn(__fsub_rn(mx, mn), 15.0f), 1e-8f);
  const float zero = mn;
  uint32_t q[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float qf = __fdiv_rn(__fsub_rn(x[i], zero), scale);
    qf = roundf(qf);
    qf = fminf(fmaxf(qf, 0.0f), 15.0f);
    q[i] = (uint32_t)qf;
  }
  const uint16_t two = (uint16_t)(((q[0] << 4) | q[1]) | (((q[2] << 4) | q[3]) << 8));
  *reinterpret_cast<uint16_t*>(packed_row + 2 * lane) = two;
  if (lane == 0) {
    *scale_p = __float2half_rn(scale);
    *zero_p = __float2half_rn(zero);
  }
}

// packed fp32 pairs (sm_100 FFMA2 / FADD2): one issue slot for two elements, same rounding as the scalar ops
__device__ __forceinline__ void fma2(float& o0, float& o1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n\t.reg .b64 a, b, c, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmov.b64 c, {%6, %7};\n\t"
      "fma.rn.f32x2 d, a, b, c;\n\tmov.b64 {%0, %1}, d;\n\t}"
      : "=f"(o0), "=f"(o1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void add2(float& o0, float& o1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 a, b, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\t"
      "add.rn.f32x2 d, a, b;\n\tmov.b64 {%0, %1}, d;\n\t}"
      : "=f"(o0), "=f"(o1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void mul2(float& o0, float& o1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 a, b, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\t"
      "mul.rn.f32x2 d, a, b;\n\tmov.b64 {%0, %1}, d;\n\t}"
      : "=f"(o0), "=f"(o1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Is ring/sink slot `j` (< sink+recent) of a streaming head holding a live token?
// See duo_cache_state in duo_b200.h.
__device__ __forceinline__ bool stream_slot_valid(int j, int sink, int recent, long long total, long long lo) {
  if (j < sink) return j < total;
  if (total <= sink) return false;
  long long last = total - 1;
  long long r = j - sink;
  long long diff = (last - sink - r) % recent;
  if (diff < 0) diff += recent;
  long long p = last - diff;
  return p >= lo;
}

// Split-KV merge: one warp merges its share of the `splits` partials (splits warp, warp+4, ...) for up to FOUR rows
// at once, i.e. 16 independent 512-byte loads in flight per iteration.  (A row-at-a-time loop costs splits/16
// dependent L2 round trips PER ROW, which dominated decode launches with a single retrieval head: measured on the
// B200 box, 1M-token decode 10.80 -> 10.54 ms of attention per step, profiles/r2_validation.md.)
//   po  : [splits][ROWS][128] un-normalised partial outputs, pml : [splits][ROWS][2] (max in log2 domain, sum)
//   rows r0 .. r0+nr-1 (nr <= 4); results: acc[q] (this lane's 4 output dims), mm[q], ll[q]
template <int ROWS>
__device__ __forceinline__ void split_merge_rows4(const float* po, const float* pml, int splits, int warp, int lane,
                                                  int r0, int nr, float4 (&acc)[4], float (&mm)[4], float (&ll)[4]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    mm[q] = -INFINITY;
    ll[q] = 0.f;
  }
  for (int s0 = warp; s0 < splits; s0 += 16) {
    float ms[4][4], ls[4][4];
    float4 vs[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int s2 = s0 + 4 * u;
      const bool sok = s2 < splits;
      const int sc2 = sok ? s2 : s0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool ok = sok && q < nr;
        const int r = r0 + (q < nr ? q : 0);
        ms[u][q] = ok ? __ldcg(&pml[(sc2 * ROWS + r) * 2]) : -INFINITY;
        ls[u][q] = __ldcg(&pml[(sc2 * ROWS + r) * 2 + 1]);
        vs[u][q] = __ldcg(reinterpret_cast<const float4*>(&po[((long long)sc2 * ROWS + r) * 128 + lane * 4]));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (ms[u][q] == -INFINITY) continue;
        const float mn = fmaxf(mm[q], ms[u][q]);
        const float fo = (mm[q] == -INFINITY) ? 0.f : fast_exp2(mm[q] - mn);
        const float fn = fast_exp2(ms[u][q] - mn);
        acc[q].x = acc[q].x * fo + vs[u][q].x * fn;
        acc[q].y = acc[q].y * fo + vs[u][q].y * fn;
        acc[q].z = acc[q].z * fo + vs[u][q].z * fn;
        acc[q].w = acc[q].w * fo + vs[u][q].w * fn;
        ll[q] = ll[q] * fo + ls[u][q] * fn;
        mm[q] = mn;
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Split-KV publish / merge protocol shared by the three bandwidth kernels (attn_mma.cu, attn_int4.cu x2).
//
// Every CTA of a (batch, kv head, row block) "item" writes its partial (un-normalised O, running max in the log2
// domain, row sum) to the workspace.  Merging is HIERARCHICAL: the splits of an item form groups of kMergeGroup; the
// last CTA of a group to arrive merges that group into a level-2 partial, the last GROUP to finish merges the level-2
// partials into the result.  Group merges happen while other CTAs are still streaming keys, so only the final merge
// of <= splits/16 partials is on the critical path.  (A single last-CTA merge of ~290-490 partials cost 34-82 us of
// a 120-150 us INT4 launch and ~30 us of a 114 us single-retrieval-head bf16 launch: profiles/r2_int4.md.)
// ---------------------------------------------------------------------------------------------
constexpr int kMergeGroup = 16;

struct SplitWs {        // device pointers into the caller's workspace
  int* counters;        // [items][1 + n_groups]  (top-level arrival count, then one per group); zero between launches
  float* ws_ml;         // [items][splits][ROWS][2]
  float* ws_o;          // [items][splits][ROWS][128]
  float* g_ml;          // [items][n_groups][ROWS][2]     level-2 partials
  float* g_o;           // [items][n_groups][ROWS][128]
  int n_groups;
};

inline int split_groups(int splits) { return splits <= kMergeGroup ? 1 : (splits + kMergeGroup - 1) / kMergeGroup; }

// The arrival counters live in a FIXED region at the start of the workspace (they must stay zero between launches, and
// launches of different geometry — layers with different numbers of retrieval heads, decode steps and chunks — share
// one workspace: a geometry-dependent counter region would overlap partial data written by an earlier launch).
constexpr size_t kSplitCounterBytes = 64 * 1024;

// bytes of workspace a launch with this geometry needs (0 if no split); SIZE_MAX if it has too many counters
inline size_t split_ws_bytes(long long items, int splits, int rows) {
  if (splits <= 1) return 0;
  const int ng = split_groups(splits);
  if ((size_t)items * (1 + ng) * 4 > kSplitCounterBytes) return (size_t)-1;
  const size_t a = 256;
  auto up = [&](size_t x) { return (x + a - 1) / a * a; };
  size_t tot = kSplitCounterBytes;
  tot += up((size_t)items * splits * rows * 2 * 4) + up((size_t)items * splits * rows * 128 * 4);
  if (ng > 1) tot += up((size_t)items * ng * rows * 2 * 4) + up((size_t)items * ng * rows * 128 * 4);
  return tot + 256;
}

inline SplitWs split_ws_carve(void* workspace, long long items, int splits, int rows) {
  SplitWs w{};
  const int ng = split_groups(splits);
  w.n_groups = ng;
  const size_t a = 256;
  auto up = [&](size_t x) { return (x + a - 1) / a * a; };
  uint8_t* p = reinterpret_cast<uint8_t*>(workspace);
  w.counters = reinterpret_cast<int*>(p);
  p += kSplitCounterBytes;
  w.ws_ml = reinterpret_cast<float*>(p);
  p += up((size_t)items * splits * rows * 2 * 4);
  w.ws_o = reinterpret_cast<float*>(p);
  p += up((size_t)items * splits * rows * 128 * 4);
  if (ng > 1) {
    w.g_ml = reinterpret_cast<float*>(p);
    p += up((size_t)items * ng * rows * 2 * 4);
    w.g_o = reinterpret_cast<float*>(p);
  }
  return w;
}

// CTA-wide (128 threads) merge of `n` partials po [n][ROWS][128] / pml [n][ROWS][2] for rows [0, rows):
// emit(r, d, a0, a1, mm, ll) receives the UN-NORMALISED sums of dims d, d+1 and the merged (max, row sum).
// cm_o: [4][16][128] floats, cm_ml: [4][16][2] floats of shared memory.
template <int ROWS, typename Emit>
__device__ __forceinline__ void merge_partials_cta(const float* po, const float* pml, int n, int rows, float* cm_o,
                                                   float* cm_ml, Emit emit) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int rg = 0; rg < rows; rg += 16) {
    const int rg_n = min(16, rows - rg);
    for (int rr0 = 0; rr0 < rg_n; rr0 += 4) {
      const int nr = min(4, rg_n - rr0);
      float4 acc4[4];
      float mm4[4], ll4[4];
      split_merge_rows4<ROWS>(po, pml, n, warp, lane, rg + rr0, nr, acc4, mm4, ll4);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q < nr) {
          *reinterpret_cast<float4*>(&cm_o[(warp * 16 + rr0 + q) * 128 + lane * 4]) = acc4[q];
          if (lane == 0) {
            cm_ml[(warp * 16 + rr0 + q) * 2] = mm4[q];
            cm_ml[(warp * 16 + rr0 + q) * 2 + 1] = ll4[q];
          }
        }
      }
    }
    __syncthreads();
    for (int idx = tid; idx < rg_n * 64; idx += 128) {
      const int rr = idx >> 6, d = (idx & 63) * 2;
      float mm = -INFINITY;
#pragma unroll
      for (int w = 0; w < 4; ++w) mm = fmaxf(mm, cm_ml[(w * 16 + rr) * 2]);
      float a0f = 0.f, a1f = 0.f, ll = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float mw = cm_ml[(w * 16 + rr) * 2];
        if (mw == -INFINITY) continue;
        const float f = fast_exp2(mw - mm);
        a0f += f * cm_o[(w * 16 + rr) * 128 + d];
        a1f += f * cm_o[(w * 16 + rr) * 128 + d + 1];
        ll += f * cm_ml[(w * 16 + rr) * 2 + 1];
      }
      emit(rg + rr, d, a0f, a1f, mm, ll);
    }
    __syncthreads();
  }
}

// The protocol.  Call with ALL 128 threads after this CTA's partial has been stored to w.ws_o / w.ws_ml.
// emit_final(r, d, v0, v1, mm, ll): NORMALISED outputs of dims d, d+1 of row r (+ the merged max / row sum), called by
// the one CTA of the item that performs the final merge.  s_flag: one int of shared memory.
template <int ROWS, typename EmitFinal>
__device__ __forceinline__ void split_kv_finish(const SplitWs& w, long long item, int split, int splits, int rows,
                                                float* cm_o, float* cm_ml, int* s_flag, EmitFinal emit_final) {
  const int tid = threadIdx.x;
  const int ng = w.n_groups;
  int* cnt = w.counters + item * (1 + ng);
  const float* po = w.ws_o + item * splits * (long long)(ROWS * 128);
  const float* pml = w.ws_ml + item * splits * (long long)(ROWS * 2);
  auto final_emit = [&](int r, int d, float a0, float a1, float mm, float ll) {
    const float inv = ll > 0.f ? 1.f / ll : 0.f;
    emit_final(r, d, a0 * inv, a1 * inv, mm, ll);
  };
  __threadfence();
  __syncthreads();
  if (ng == 1) {
    if (tid == 0) *s_flag = (atomicAdd(&cnt[0], 1) == splits - 1);
    __syncthreads();
    if (!*s_flag) return;
    __threadfence();
    merge_partials_cta<ROWS>(po, pml, splits, rows, cm_o, cm_ml, final_emit);
    if (tid == 0) cnt[0] = 0;  // leave the workspace ready for the next launch
    return;
  }
  const int grp = split / kMergeGroup;
  const int gsz = min(kMergeGroup, splits - grp * kMergeGroup);
  if (tid == 0) *s_flag = (atomicAdd(&cnt[1 + grp], 1) == gsz - 1);
  __syncthreads();
  if (!*s_flag) return;
  __threadfence();
  float* go = w.g_o + (item * ng + grp) * (long long)(ROWS * 128);
  float* gml = w.g_ml + (item * ng + grp) * (long long)(ROWS * 2);
  merge_partials_cta<ROWS>(po + (long long)grp * kMergeGroup * (ROWS * 128), pml + (long long)grp * kMergeGroup * (ROWS * 2),
                           gsz, rows, cm_o, cm_ml, [&](int r, int d, float a0, float a1, float mm, float ll) {
                             *reinterpret_cast<float2*>(&go[r * 128 + d]) = make_float2(a0, a1);
                             if (d == 0) {
                               gml[r * 2] = mm;
                               gml[r * 2 + 1] = ll;
                             }
                           });
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    cnt[1 + grp] = 0;
    *s_flag = (atomicAdd(&cnt[0], 1) == ng - 1);
  }
  __syncthreads();
  if (!*s_flag) return;
  __threadfence();
  merge_partials_cta<ROWS>(w.g_o + item * ng * (long long)(ROWS * 128), w.g_ml + item * ng * (long long)(ROWS * 2), ng, rows,
                           cm_o, cm_ml, final_emit);
  if (tid == 0) cnt[0] = 0;
}

// ---------------------------------------------------------------------------------------------
// per-DEVICE launch attributes (one process may drive several GPUs, as the reference's tensor_parallel mode does,
// duo_attn/utils.py:206-227: cudaFuncSetAttribute is per device, so the "already set" memo is a device bit mask)
// ---------------------------------------------------------------------------------------------
int sm_count_current_device();  // api.cu
template <typename K>
inline int ensure_dyn_smem(K kern, int bytes, unsigned long long* done_mask, bool max_carveout = false) {
  int dev = 0;
  DUO_CUDA_TRY(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (__atomic_load_n(done_mask, __ATOMIC_ACQUIRE) & bit) return DUO_OK;
  DUO_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (max_carveout)
    DUO_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  __atomic_fetch_or(done_mask, bit, __ATOMIC_RELEASE);
  return DUO_OK;
}

}  // namespace duo
