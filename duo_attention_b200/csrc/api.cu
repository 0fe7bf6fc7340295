// C ABI of libduo_b200.so (see include/duo_b200.h).  Thin: argument checks, TMA descriptor
// encoding, dispatch to the launchers.  Never throws, never allocates device memory.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "duo_common.cuh"

namespace duo {

int sm_count_current_device() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int& v = sms[dev & 63];
  if (v == 0) {
    int n = 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
    v = n;
  }
  return v;
}

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return DUO_ECUDA;
}

// launchers implemented in the other translation units
size_t mma_workspace_bytes(int batch, int n_kv, int group, int max_q_len);
int launch_attn_mma(const duo_layer* L, const duo_cache_state* st, const void* q, long long q_row_stride, void* out,
                    int q_len, float scale, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int launch_attn_int4(const duo_layer* L, const duo_cache_state* st, const void* q, long long q_row_stride, void* out,
                     int q_len, float scale, void* workspace, size_t workspace_bytes, cudaStream_t stream);
bool tc_prefill_supported(const duo_layer* L, const duo_cache_state* st, int q_len);
int launch_attn_tc(const duo_layer* L, const duo_cache_state* st, const void* q, long long q_row_stride, void* out,
                   int q_len, float scale, cudaStream_t stream);
int launch_rope_append(const duo_layer* L, const duo_cache_state* st, void* qkv, long long row_stride, const void* cos,
                       const void* sin, int rope_mode, int q_len, cudaStream_t stream);
int launch_stream_commit(const duo_layer* L, const duo_cache_state* st, int q_len, cudaStream_t stream);
int launch_quant_int4(const void* in, long long in_row_stride, long long rows, void* packed, void* scale, void* zero,
                      cudaStream_t stream);
int launch_state_advance(long long* st, int n, int sink, int recent, cudaStream_t stream);
int launch_state_set(long long* st, long long full_len, long long total, long long lo, cudaStream_t stream);
int launch_dequant_int4(const void* packed, const void* scale, const void* zero, long long rows, void* out,
                        cudaStream_t stream);

int launch_attn_mma_partial(const duo_layer* L, long long n_keys, const void* q, long long q_row_stride, float* out_o,
                            float* out_lse, int q_len, float scale, void* workspace, size_t workspace_bytes,
                            cudaStream_t stream);
int launch_attn_mma_seq(const duo_layer* L, const duo_cache_state* st, const void* q, long long q_row_stride, void* out,
                        float* part_o, float* part_lse, int q_len, float scale, void* workspace, size_t workspace_bytes,
                        cudaStream_t stream);
int launch_decode_fused(const duo_layer* L, const duo_cache_state* st, const void* qkv, long long row_stride,
                        const void* cos, const void* sin, int rope_mode, void* out, int q_len, float scale,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream);
int launch_decode_fused_int4(const duo_layer* L, const duo_cache_state* st, const void* qkv, long long row_stride,
                             const void* cos, const void* sin, int rope_mode, void* out, int q_len, float scale,
                             void* workspace, size_t workspace_bytes, cudaStream_t stream);
int launch_decode_fused_seq(const duo_layer* L, const duo_cache_state* st, const void* qkv, long long row_stride,
                            const void* cos, const void* sin, int rope_mode, void* out, float* part_o, float* part_lse,
                            float scale, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int launch_merge_partials(const float* o_parts, const float* lse_parts, int n_parts, long long tokens, int heads_total,
                          int heads_used, void* out, int dtype, cudaStream_t stream);
int launch_add_rmsnorm(const void* x, const void* residual, const void* weight, void* out_norm, void* out_res,
                       long long rows, int hidden, float eps, int dtype, cudaStream_t stream);
int launch_silu_mul(const void* gate_up, void* out, long long rows, int inter, int dtype, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// TMA descriptors (cuTensorMapEncodeTiled fetched through the runtime: no -lcuda link dependency)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
    set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// {head_dim, slots, batch*heads} 16-bit tensor; box {64, box_rows, 1}; 128B swizzle; OOB rows read as zero.
static int encode_kv_map(CUtensorMap* m, void* base, int dtype, long long slots, long long heads, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return DUO_ECUDA;
  cuuint64_t dims[3] = {(cuuint64_t)kHeadDim, (cuuint64_t)slots, (cuuint64_t)heads};
  cuuint64_t strides[2] = {(cuuint64_t)kHeadDim * 2, (cuuint64_t)slots * kHeadDim * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, dtype == DUO_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base,
                  dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (base=%p slots=%lld heads=%lld)", (int)r, base, slots,
              heads);
    return DUO_ECUDA;
  }
  return DUO_OK;
}

// first staging slot of the streaming cache: right after the ring for 16-bit caches (TMA takes any
// coordinate); rounded up to a multiple of 64 for INT4 so that every 64-key tile (and its scale/zero rows)
// starts 16-byte aligned for cp.async.
int stage_offset(const duo_layer_desc& d) {
  const int W = d.sink + d.recent;
  return d.kv_format == DUO_KV_INT4 ? (W + 63) / 64 * 64 : W;
}

}  // namespace duo

using namespace duo;

extern "C" {

const char* duo_last_error_string(void) { return g_err; }
int duo_version(void) { return 100; }

int duo_layer_create(const duo_layer_desc* desc, duo_layer** out) {
  if (!desc || !out) {
    set_error("duo_layer_create: null argument");
    return DUO_EINVAL;
  }
  *out = nullptr;
  if (desc->head_dim != kHeadDim) {
    set_error("duo_layer_create: head_dim %d unsupported (only 128)", desc->head_dim);
    return DUO_EINVAL;
  }
  if (desc->batch < 1 || desc->n_full < 0 || desc->n_stream < 0 || desc->group < 1 || desc->sink < 0 ||
      desc->recent < 1 || desc->stage_cap < 1 || desc->full_cap < 0) {
    set_error("duo_layer_create: bad geometry");
    return DUO_EINVAL;
  }
  if (desc->dtype != DUO_DT_BF16 && desc->dtype != DUO_DT_FP16) {
    set_error("duo_layer_create: bad dtype %d", desc->dtype);
    return DUO_EINVAL;
  }
  if (desc->kv_format == DUO_KV_INT4 && desc->dtype != DUO_DT_FP16) {
    set_error("duo_layer_create: INT4 KV needs fp16 activations (demo/run_duo_w8a8kv4.py:41-45)");
    return DUO_EINVAL;
  }
  if (desc->kv_format == DUO_KV_INT4) {
    const long long ring_slots = (long long)stage_offset(*desc) + desc->stage_cap;
    if (desc->full_cap % 8 != 0 || ring_slots % 8 != 0) {
      set_error("duo_layer_create: INT4 caches need full_cap and ring slots (%lld) to be multiples of 8", ring_slots);
      return DUO_EINVAL;
    }
  }
  duo_layer* L = new (std::nothrow) duo_layer();
  if (!L) {
    set_error("duo_layer_create: out of host memory");
    return DUO_EINVAL;
  }
  L->d = *desc;
  L->has_full_maps = false;
  L->has_ring_maps = false;
  memset(&L->maps, 0, sizeof(L->maps));
  if (desc->kv_format == DUO_KV_SAME) {
    const long long ring_slots = (long long)desc->sink + desc->recent + desc->stage_cap;
    int rc = DUO_OK;
    if (desc->n_full > 0 && desc->full_cap > 0) {
      const long long heads = (long long)desc->batch * desc->n_full;
      rc = encode_kv_map(&L->maps.full_k64, desc->full_k, desc->dtype, desc->full_cap, heads, 64);
      if (!rc) rc = encode_kv_map(&L->maps.full_v64, desc->full_v, desc->dtype, desc->full_cap, heads, 64);
      if (!rc) rc = encode_kv_map(&L->maps.full_k128, desc->full_k, desc->dtype, desc->full_cap, heads, 128);
      if (!rc) rc = encode_kv_map(&L->maps.full_v128, desc->full_v, desc->dtype, desc->full_cap, heads, 128);
      L->has_full_maps = (rc == DUO_OK);
    }
    if (!rc && desc->n_stream > 0) {
      const long long heads = (long long)desc->batch * desc->n_stream;
      rc = encode_kv_map(&L->maps.ring_k64, desc->ring_k, desc->dtype, ring_slots, heads, 64);
      if (!rc) rc = encode_kv_map(&L->maps.ring_v64, desc->ring_v, desc->dtype, ring_slots, heads, 64);
      if (!rc) rc = encode_kv_map(&L->maps.ring_k128, desc->ring_k, desc->dtype, ring_slots, heads, 128);
      if (!rc) rc = encode_kv_map(&L->maps.ring_v128, desc->ring_v, desc->dtype, ring_slots, heads, 128);
      L->has_ring_maps = (rc == DUO_OK);
    }
    if (rc) {
      delete L;
      return rc;
    }
  }
  *out = L;
  return DUO_OK;
}

void duo_layer_destroy(duo_layer* layer) { delete layer; }

size_t duo_workspace_bytes(int32_t batch, int32_t n_kv_heads, int32_t group, int32_t max_q_len) {
  return mma_workspace_bytes(batch, n_kv_heads, group, max_q_len);
}

static int check_chunk(const duo_layer* L, const duo_cache_state* st, int q_len, const char* who) {
  if (!L || !st) {
    set_error("%s: null layer/state", who);
    return DUO_EINVAL;
  }
  if (q_len < 1) {
    set_error("%s: q_len %d < 1", who, q_len);
    return DUO_EINVAL;
  }
  if (st->full_len < 0 || st->total < 0 || st->lo < 0) {
    set_error("%s: negative cache state", who);
    return DUO_EINVAL;
  }
  long long need = st->full_len + q_len;
  if (st->seq_world != 0) {
    if (st->seq_world < 2 || st->seq_world > 8 || st->seq_rank < 0 || st->seq_rank >= st->seq_world || st->seq_block < 1) {
      set_error("%s: bad sequence-shard descriptor (rank %d, world %d, block %d)", who, st->seq_rank, st->seq_world,
                st->seq_block);
      return DUO_EINVAL;
    }
    const long long round = (long long)st->seq_block * st->seq_world, rem = need % round;
    long long extra = rem - (long long)st->seq_rank * st->seq_block;
    extra = extra < 0 ? 0 : (extra > st->seq_block ? st->seq_block : extra);
    need = need / round * st->seq_block + extra;  // rows of this rank's slice after the append
  }
  if (L->d.n_full > 0 && need > L->d.full_cap) {
    set_error("Trying to put %d KVs into a cache with max size %lld, current size: %lld.", q_len,
              (long long)L->d.full_cap, (long long)st->full_len);
    return DUO_EOVERFLOW;
  }
  if (q_len > L->d.stage_cap) {
    set_error("%s: chunk of %d tokens exceeds the staging capacity %d", who, q_len, L->d.stage_cap);
    return DUO_EOVERFLOW;
  }
  return DUO_OK;
}

int duo_rope_append(const duo_layer* layer, const duo_cache_state* st, void* qkv, int64_t qkv_row_stride,
                    const void* cos, const void* sin, int32_t rope_mode, int32_t q_len, void* stream) {
  int rc = check_chunk(layer, st, q_len, "duo_rope_append");
  if (rc) return rc;
  if (!qkv || ((rope_mode & 0xff) != DUO_ROPE_NONE && (!cos || !sin))) {
    set_error("duo_rope_append: null buffer");
    return DUO_EINVAL;
  }
  if ((rope_mode & 0xff) < DUO_ROPE_NONE || (rope_mode & 0xff) > DUO_ROPE_FP32 || (rope_mode & ~(0xff | DUO_ROPE_SKIP_Q))) {
    set_error("duo_rope_append: bad rope_mode %d", rope_mode);
    return DUO_EINVAL;
  }
  if (qkv_row_stride % 4 != 0) {
    set_error("duo_rope_append: row stride must be a multiple of 4 elements");
    return DUO_EINVAL;
  }
  return launch_rope_append(layer, st, qkv, qkv_row_stride, cos, sin, rope_mode, q_len, (cudaStream_t)stream);
}

int duo_attention(const duo_layer* layer, const duo_cache_state* st, const void* q, int64_t q_row_stride, void* out,
                  int32_t q_len, float scale, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_chunk(layer, st, q_len, "duo_attention");
  if (rc) return rc;
  if (!q || !out) {
    set_error("duo_attention: null buffer");
    return DUO_EINVAL;
  }
  if (st->seq_world != 0) {
    set_error("duo_attention: sequence-sharded caches are attended with duo_attention_seq");
    return DUO_EINVAL;
  }
  if (layer->d.kv_format == DUO_KV_INT4)
    return launch_attn_int4(layer, st, q, q_row_stride, out, q_len, scale, workspace, workspace_bytes,
                            (cudaStream_t)stream);
  if (st->device_state && q_len > DUO_DECODE_MAX_Q) {
    set_error("duo_attention: device_state is only supported for chunks of <= %d tokens", DUO_DECODE_MAX_Q);
    return DUO_EINVAL;
  }
  if (tc_prefill_supported(layer, st, q_len))
    return launch_attn_tc(layer, st, q, q_row_stride, out, q_len, scale, (cudaStream_t)stream);
  return launch_attn_mma(layer, st, q, q_row_stride, out, q_len, scale, workspace, workspace_bytes,
                         (cudaStream_t)stream);
}

int duo_decode_fused(const duo_layer* layer, const duo_cache_state* st, const void* qkv, int64_t qkv_row_stride,
                     const void* cos, const void* sin, int32_t rope_mode, void* out, int32_t q_len, float scale,
                     void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_chunk(layer, st, q_len, "duo_decode_fused");
  if (rc) return rc;
  if (!qkv || !out || (rope_mode != DUO_ROPE_NONE && (!cos || !sin))) {
    set_error("duo_decode_fused: null buffer");
    return DUO_EINVAL;
  }
  if (rope_mode < DUO_ROPE_NONE || rope_mode > DUO_ROPE_FP32) {
    set_error("duo_decode_fused: bad rope_mode %d", rope_mode);
    return DUO_EINVAL;
  }
  const int max_rows = layer->d.kv_format == DUO_KV_INT4 ? DUO_DECODE_MAX_Q_INT4 : DUO_DECODE_MAX_Q;
  if (layer->d.group * q_len > max_rows || st->seq_world != 0) {
    set_error("duo_decode_fused: unsharded caches and group * q_len <= %d only (got group %d, q_len %d)", max_rows,
              layer->d.group, q_len);
    return DUO_EINVAL;
  }
  if (qkv_row_stride % 8 != 0 || (reinterpret_cast<uintptr_t>(qkv) & 15)) {
    set_error("duo_decode_fused: qkv rows must be 16-byte aligned (row stride a multiple of 8 elements)");
    return DUO_EINVAL;
  }
  if (layer->d.kv_format == DUO_KV_INT4)
    return launch_decode_fused_int4(layer, st, qkv, qkv_row_stride, cos, sin, rope_mode, out, q_len, scale, workspace,
                                    workspace_bytes, (cudaStream_t)stream);
  return launch_decode_fused(layer, st, qkv, qkv_row_stride, cos, sin, rope_mode, out, q_len, scale, workspace,
                             workspace_bytes, (cudaStream_t)stream);
}

int duo_decode_fused_seq(const duo_layer* layer, const duo_cache_state* st, const void* qkv, int64_t qkv_row_stride,
                         const void* cos, const void* sin, int32_t rope_mode, void* out, float* out_o, float* out_lse,
                         float scale, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_chunk(layer, st, 1, "duo_decode_fused_seq");
  if (rc) return rc;
  if (!qkv || !out || !out_o || !out_lse || (rope_mode != DUO_ROPE_NONE && (!cos || !sin))) {
    set_error("duo_decode_fused_seq: null buffer");
    return DUO_EINVAL;
  }
  if (rope_mode < DUO_ROPE_NONE || rope_mode > DUO_ROPE_FP32 || st->seq_world < 2) {
    set_error("duo_decode_fused_seq: bad rope_mode %d or no sequence-shard descriptor", rope_mode);
    return DUO_EINVAL;
  }
  if (layer->d.kv_format != DUO_KV_SAME || layer->d.group > 16) {
    set_error("duo_decode_fused_seq: 16-bit caches and group <= 16 only");
    return DUO_EINVAL;
  }
  if (qkv_row_stride % 8 != 0 || (reinterpret_cast<uintptr_t>(qkv) & 15)) {
    set_error("duo_decode_fused_seq: qkv rows must be 16-byte aligned (row stride a multiple of 8 elements)");
    return DUO_EINVAL;
  }
  return launch_decode_fused_seq(layer, st, qkv, qkv_row_stride, cos, sin, rope_mode, out, out_o, out_lse, scale, workspace,
                                 workspace_bytes, (cudaStream_t)stream);
}

// test / tuning hook: force the mma.sync kernel family even for shapes the tcgen05 kernel takes
int duo_attention_mma(const duo_layer* layer, const duo_cache_state* st, const void* q, int64_t q_row_stride,
                      void* out, int32_t q_len, float scale, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_chunk(layer, st, q_len, "duo_attention_mma");
  if (rc) return rc;
  if (layer->d.kv_format != DUO_KV_SAME) {
    set_error("duo_attention_mma: 16-bit KV only");
    return DUO_EINVAL;
  }
  return launch_attn_mma(layer, st, q, q_row_stride, out, q_len, scale, workspace, workspace_bytes,
                         (cudaStream_t)stream);
}

int duo_state_advance(int64_t* device_state, int32_t n, int32_t sink, int32_t recent, void* stream) {
  if (!device_state || n < 0 || recent < 1 || sink < 0) {
    set_error("duo_state_advance: bad argument");
    return DUO_EINVAL;
  }
  return launch_state_advance(reinterpret_cast<long long*>(device_state), n, sink, recent, (cudaStream_t)stream);
}

int duo_state_set(int64_t* device_state, int64_t full_len, int64_t total, int64_t lo, void* stream) {
  if (!device_state || full_len < 0 || total < 0 || lo < 0) {
    set_error("duo_state_set: bad argument");
    return DUO_EINVAL;
  }
  return launch_state_set(reinterpret_cast<long long*>(device_state), full_len, total, lo, (cudaStream_t)stream);
}

int duo_stream_commit(const duo_layer* layer, const duo_cache_state* st, int32_t q_len, void* stream) {
  int rc = check_chunk(layer, st, q_len, "duo_stream_commit");
  if (rc) return rc;
  return launch_stream_commit(layer, st, q_len, (cudaStream_t)stream);
}

int duo_quant_int4(const void* in, int64_t in_row_stride, int64_t rows, void* packed, void* scale, void* zero,
                   void* stream) {
  if (rows < 0 || (rows > 0 && (!in || !packed || !scale || !zero)) || in_row_stride % 4 != 0) {
    set_error("duo_quant_int4: bad argument");
    return DUO_EINVAL;
  }
  return launch_quant_int4(in, in_row_stride, rows, packed, scale, zero, (cudaStream_t)stream);
}

int duo_attention_partial(const duo_layer* layer, int64_t n_keys, const void* q, int64_t q_row_stride, float* out_o,
                          float* out_lse, int32_t q_len, float scale, void* workspace, size_t workspace_bytes,
                          void* stream) {
  if (!layer || !q || !out_o || !out_lse || n_keys < 0 || q_len < 1) {
    set_error("duo_attention_partial: bad argument");
    return DUO_EINVAL;
  }
  if (layer->d.kv_format != DUO_KV_SAME || layer->d.group * q_len > 16) {
    set_error("duo_attention_partial: 16-bit caches and group * q_len <= 16 only (got group %d, q_len %d)",
              layer->d.group, q_len);
    return DUO_EINVAL;
  }
  if (layer->d.n_full > 0 && n_keys > layer->d.full_cap) {
    set_error("duo_attention_partial: n_keys %lld exceeds the cache capacity %lld", (long long)n_keys,
              (long long)layer->d.full_cap);
    return DUO_EOVERFLOW;
  }
  return launch_attn_mma_partial(layer, n_keys, q, q_row_stride, out_o, out_lse, q_len, scale, workspace,
                                 workspace_bytes, (cudaStream_t)stream);
}

int duo_attention_seq(const duo_layer* layer, const duo_cache_state* st, const void* q, int64_t q_row_stride, void* out,
                      float* out_o, float* out_lse, int32_t q_len, float scale, void* workspace, size_t workspace_bytes,
                      void* stream) {
  int rc = check_chunk(layer, st, q_len, "duo_attention_seq");
  if (rc) return rc;
  if (!q || !out || !out_o || !out_lse) {
    set_error("duo_attention_seq: null buffer");
    return DUO_EINVAL;
  }
  if (st->seq_world < 2) {
    set_error("duo_attention_seq: the cache state carries no sequence-shard descriptor");
    return DUO_EINVAL;
  }
  if (layer->d.kv_format != DUO_KV_SAME || layer->d.group * q_len > 16) {
    set_error("duo_attention_seq: 16-bit caches and group * q_len <= 16 only (got group %d, q_len %d)", layer->d.group,
              q_len);
    return DUO_EINVAL;
  }
  return launch_attn_mma_seq(layer, st, q, q_row_stride, out, out_o, out_lse, q_len, scale, workspace, workspace_bytes,
                             (cudaStream_t)stream);
}

int duo_merge_partials(const float* o_parts, const float* lse_parts, int32_t n_parts, int64_t tokens,
                       int32_t heads_total, int32_t heads_used, void* out, int32_t dtype, void* stream) {
  if (n_parts < 1 || tokens < 0 || heads_total < 1 || heads_used < 0 || heads_used > heads_total ||
      (tokens > 0 && heads_used > 0 && (!o_parts || !lse_parts || !out)) ||
      (dtype != DUO_DT_BF16 && dtype != DUO_DT_FP16)) {
    set_error("duo_merge_partials: bad argument");
    return DUO_EINVAL;
  }
  return launch_merge_partials(o_parts, lse_parts, n_parts, tokens, heads_total, heads_used, out, dtype,
                               (cudaStream_t)stream);
}

int duo_add_rmsnorm(const void* x, const void* residual, const void* weight, void* out_norm, void* out_res, int64_t rows,
                    int32_t hidden, float eps, int32_t dtype, void* stream) {
  if (rows < 0 || hidden < 8 || hidden % 8 != 0 || hidden > 16384 || (rows > 0 && (!x || !weight || !out_norm)) ||
      (dtype != DUO_DT_BF16 && dtype != DUO_DT_FP16)) {
    set_error("duo_add_rmsnorm: bad argument");
    return DUO_EINVAL;
  }
  return launch_add_rmsnorm(x, residual, weight, out_norm, out_res, rows, hidden, eps, dtype, (cudaStream_t)stream);
}

int duo_silu_mul(const void* gate_up, void* out, int64_t rows, int32_t inter, int32_t dtype, void* stream) {
  if (rows < 0 || inter < 8 || inter % 8 != 0 || (rows > 0 && (!gate_up || !out)) ||
      (dtype != DUO_DT_BF16 && dtype != DUO_DT_FP16)) {
    set_error("duo_silu_mul: bad argument");
    return DUO_EINVAL;
  }
  return launch_silu_mul(gate_up, out, rows, inter, dtype, (cudaStream_t)stream);
}

int duo_dequant_int4(const void* packed, const void* scale, const void* zero, int64_t rows, void* out, void* stream) {
  if (rows < 0 || (rows > 0 && (!packed || !scale || !zero || !out))) {
    set_error("duo_dequant_int4: bad argument");
    return DUO_EINVAL;
  }
  return launch_dequant_int4(packed, scale, zero, rows, out, (cudaStream_t)stream);
}

}  // extern "C"
