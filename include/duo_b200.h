/*
 * duo_b200.h — C ABI of libduo_b200.so: DuoAttention's mixed-head (retrieval + streaming)
 * attention hot path, hand-written for NVIDIA B200 (sm_100a).
 *
 * Plain C: raw device pointers, integers and a cudaStream_t passed as void*.  No torch types.
 * Every function returns 0 on success or a negative DUO_E* code; the message of the last error
 * on the calling thread is available from duo_last_error_string().  No function throws,
 * none allocates device memory (the caller — PyTorch in the reference — owns all buffers),
 * and all work is enqueued on the given stream (re-entrant per stream).
 *
 * What each entry point replaces in the reference (paths relative to mit-han-lab/duo-attention):
 *
 *   duo_layer_create/destroy  – the per-layer views of DuoAttentionStaticKVCache
 *                               (duo_attn/patch/static_kv_cache.py:60-94) re-laid out head-major;
 *                               also pre-encodes the TMA descriptors of the four cache tensors.
 *   duo_rope_append           – apply_rotary_pos_emb (duo_attn/patch/llama.py:177-184) or
 *                               apply_rope_inplace (duo_attn/patch/flashinfer_utils.py:29-59),
 *                               kv_cache.split_kv + put_full_kv (static_kv_cache.py:252-263,109-125)
 *                               and the torch.cat of cached+new streaming KV (llama.py:385-390);
 *                               for INT4 caches also quantize_int4_with_zero_point_per_group
 *                               (demo/quantize_int4.cu:73-178 via demo/int4_kv.py:261-371).
 *   duo_attention             – the flash_attn_func call pair + torch.cat
 *                               (llama.py:225-267 / :364-421; demo/w8a8kv4_llama.py:229-274),
 *                               both head classes in ONE launch; for INT4 caches the
 *                               dequantize pass (demo/int4_kv.py:373-436) is folded into the K/V load.
 *   duo_stream_commit         – compress_and_replace_streaming_kv (static_kv_cache.py:127-167,
 *                               llama.py:273-290; demo/int4_kv.py:438-492) as a ring advance.
 *   duo_quant_int4 / duo_dequant_int4 – the two kernels of demo/quantize_int4.cu (K1 :73-144,
 *                               K2 :9-42) as stand-alone ops (K2 is test/diagnostic only: the
 *                               product never materialises a dequantised cache).
 */
#ifndef DUO_B200_H
#define DUO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DUO_API __attribute__((visibility("default")))
#else
#define DUO_API
#endif

#define DUO_OK 0
#define DUO_EINVAL -1     /* bad argument / unsupported shape          */
#define DUO_EOVERFLOW -2  /* KV cache capacity exceeded (ValueError in the reference,
                             static_kv_cache.py:112-115)               */
#define DUO_ECUDA -3      /* CUDA runtime / driver error                */
#define DUO_EWORKSPACE -4 /* workspace too small                        */

/* element type of activations and (for DUO_KV_SAME) of the KV cache */
#define DUO_DT_BF16 0
#define DUO_DT_FP16 1
/* KV cache storage */
#define DUO_KV_SAME 0 /* same dtype as activations                                     */
#define DUO_KV_INT4 1 /* packed u4 + fp16 scale/zero per (token, head); activations fp16 */

/* RoPE flavour of duo_rope_append */
#define DUO_ROPE_NONE 0  /* q/k already rotated                                              */
#define DUO_ROPE_HF 1    /* cos/sin tables in the activation dtype, HF op order (llama.py:177-184):
                            round(round(x*cos) + round(rotate_half(x)*sin)) — bit-exact with torch */
#define DUO_ROPE_FP32 2  /* cos/sin tables in fp32, fp32 math, one rounding (flashinfer semantics,
                            flashinfer_utils.py:29-59, with accurate trig)                  */
#define DUO_ROPE_SKIP_Q 0x100 /* OR-able flag: leave q untouched (it was rotated by an earlier call on the
                                 same buffer); k is still rotated on its way into the caches */

/*
 * Geometry + buffers of ONE decoder layer's KV cache, head-major ("heads first, then tokens"):
 *
 *   full_k/full_v : [batch][n_full  ][full_cap            ][head_dim]   retrieval heads
 *   ring_k/ring_v : [batch][n_stream][sink+recent+stage_cap][head_dim]  streaming heads:
 *                   slots [0,sink) = attention sinks, [sink,sink+recent) = ring of the most recent
 *                   tokens (token p lives in slot sink + (p-sink) % recent), and
 *                   [sink+recent, +stage_cap) = staging area holding the K/V of the chunk being
 *                   processed until duo_stream_commit moves its tail into the ring.
 *   KV heads are in the reference's reordered order: retrieval heads first
 *   (duo_attn/patch/utils.py:6-45); q-head i reads kv-head i / group.
 *
 *   DUO_KV_INT4: the staging area starts at slot round_up(sink+recent, 64) instead of sink+recent (the
 *   slots in between are unused) and full_cap / the ring slot count must be multiples of 8, so that every
 *   64-key tile and its scale/zero rows are 16-byte aligned.
 *   DUO_KV_INT4: the k/v tensors hold head_dim/2 bytes per row (high nibble = even element,
 *   demo/quantize_int4.cu:33-40,137) and *_scale / *_zero are fp16 [batch][heads][slots]
 *   (group_size == head_dim == 128, demo/int4_kv.py:140).
 */
typedef struct duo_layer_desc {
  void* full_k;
  void* full_v;
  void* ring_k;
  void* ring_v;
  void* full_k_scale; /* INT4 only, else NULL */
  void* full_k_zero;
  void* full_v_scale;
  void* full_v_zero;
  void* ring_k_scale;
  void* ring_k_zero;
  void* ring_v_scale;
  void* ring_v_zero;
  int64_t full_cap;  /* token capacity of the retrieval cache          */
  int32_t batch;
  int32_t n_full;    /* retrieval KV heads in this layer (0..n_kv)      */
  int32_t n_stream;  /* streaming KV heads (n_kv - n_full)              */
  int32_t group;     /* q heads per kv head                             */
  int32_t head_dim;  /* must be 128                                     */
  int32_t sink;
  int32_t recent;
  int32_t stage_cap; /* max tokens per call (prefill chunk capacity)    */
  int32_t dtype;     /* DUO_DT_*                                        */
  int32_t kv_format; /* DUO_KV_*                                        */
} duo_layer_desc;

/* Streaming/retrieval cache occupancy BEFORE the call (host integers; the caller advances them,
 * exactly like kv_seq_len_list / streaming_kv_seq_len_list in static_kv_cache.py:44-45).
 *   full_len : tokens in the retrieval cache
 *   total    : tokens seen so far by the streaming heads (== full_len unless evict_last ran)
 *   lo       : oldest token position still valid in the ring (>= sink); ring content is
 *              positions [lo, total) ∩ [sink, ∞), sinks are positions [0, min(total, sink)).   */
typedef struct duo_cache_state {
  int64_t full_len;
  int64_t total;
  int64_t lo;
  /* Optional (may be NULL): device array {full_len, total, lo}.  When set, duo_rope_append, duo_attention (chunks
   * of at most DUO_DECODE_MAX_Q tokens) and duo_stream_commit read the occupancy from DEVICE memory at kernel
   * start, so a captured CUDA graph of a decode step can be replayed while the context grows; the host values
   * above must then be upper-bound-consistent (they size the launch and drive the capacity checks).
   * duo_state_advance moves the device copy forward after a step. */
  const int64_t* device_state;
  /* Sequence sharding of the retrieval heads across tensor-parallel ranks (scope row f1; all zero = off).  With
   * seq_world > 1 the layer's full_k/full_v hold only the block-cyclic slice of rank seq_rank — token position p lives
   * on rank (p / seq_block) % seq_world at local row (p / (seq_block*seq_world)) * seq_block + p % seq_block, so a
   * slice is ordered by position and balanced at any length — while full_len keeps counting GLOBAL tokens (full_cap
   * is the local capacity).  duo_rope_append then appends only the positions this rank owns; duo_attention_seq
   * attends the local slice.  Streaming heads are not sharded.  The reference shards by head only
   * (duo_attn/utils.py:151-179). */
  int32_t seq_rank, seq_world, seq_block, seq_reserved;
} duo_cache_state;

typedef struct duo_layer duo_layer; /* opaque: desc + pre-encoded TMA descriptors (host memory) */

DUO_API int duo_layer_create(const duo_layer_desc* desc, duo_layer** out);
DUO_API void duo_layer_destroy(duo_layer* layer);

/* Bytes of scratch duo_attention needs for this geometry (split-KV partials + arrival counters).
 * The buffer must be zero-initialised once; the kernels leave the counters zeroed.            */
DUO_API size_t duo_workspace_bytes(int32_t batch, int32_t n_kv_heads, int32_t group, int32_t max_q_len);

/*
 * RoPE + KV append for one chunk of q_len tokens (all batch rows).
 *   qkv       : [batch][q_len][(n_q + 2 n_kv) * head_dim] fused projection output, row stride
 *               qkv_row_stride elements; q is rotated IN PLACE, k is rotated on its way into the
 *               caches, v is copied (INT4: both quantised, K1 semantics).
 *   cos, sin  : [q_len][head_dim] tables (dtype per rope_mode), shared by all batch rows.
 * Retrieval heads' K/V go to full_{k,v}[.., full_len + t, :]; streaming heads' K/V go to the
 * staging slots ring_{k,v}[.., sink+recent + t, :].
 * Returns DUO_EOVERFLOW if full_len + q_len > full_cap or q_len > stage_cap.
 */
DUO_API int duo_rope_append(const duo_layer* layer, const duo_cache_state* st, void* qkv, int64_t qkv_row_stride,
                    const void* cos, const void* sin, int32_t rope_mode, int32_t q_len, void* stream);

/*
 * Mixed-head attention for one chunk, both head classes in one launch.  Must follow
 * duo_rope_append for the same chunk and state.
 *   q   : [batch][q_len][n_q][head_dim], token stride q_row_stride elements (the rotated q inside qkv)
 *   out : [batch][q_len][n_q][head_dim] contiguous
 * Retrieval q-heads attend keys [0, full_len + t] (bottom-right causal); streaming q-heads attend
 * the valid sink+ring slots plus staged chunk tokens [0, t].  softmax scale `scale`
 * (1/sqrt(head_dim) in the reference), fp32 softmax, P rounded to the activation dtype before PV.
 * q_len <= DUO_DECODE_MAX_Q uses the split-KV bandwidth kernel, larger chunks the tensor-core
 * prefill kernel.
 */
#define DUO_DECODE_MAX_Q 16
DUO_API int duo_attention(const duo_layer* layer, const duo_cache_state* st, const void* q, int64_t q_row_stride,
                  void* out, int32_t q_len, float scale, void* workspace, size_t workspace_bytes,
                  void* stream);

/*
 * One launch for a whole decode-sized chunk (group * q_len <= 16): duo_rope_append + duo_attention +
 * duo_stream_commit fused — RoPE of q in registers, RoPE(k) / v of the new tokens written straight to their final
 * cache rows (retrieval: full_len + t; streaming: sink / ring slot, no staging round trip) and attended from a tile
 * built in shared memory.  Replaces, per decoder layer and decode step, what the reference does in
 * duo_attn/patch/llama.py:347-362 (RoPE), :353-362 + static_kv_cache.py:109-125 (append), :364-421 (attention) and
 * :423-425 + static_kv_cache.py:127-167 (streaming compaction).  `qkv` as for duo_rope_append but NOT modified;
 * cos / sin / rope_mode as for duo_rope_append (no DUO_ROPE_SKIP_Q); rows must be 16-byte aligned.
 * INT4 caches (group * q_len <= DUO_DECODE_MAX_Q_INT4, fp16 activations): the K1 quantisation of the new K / V
 * (demo/quantize_int4.cu:73-144, done by the reference in int4_kv.py:261-371 before every attention call) is part of the
 * same launch — the CTA that owns the end of a head's key range rotates and quantises the new rows into the cache, reads
 * them back with the rest of its keys, and commits the streaming ring when it has drained its pipeline; cache content
 * and outputs are bit-identical to duo_rope_append + duo_attention + duo_stream_commit.  The very first chunk of a
 * sequence must still go through the three calls on an fp16 layer: the reference attends the raw K / V there
 * (demo/w8a8kv4_llama.py:229-238).
 */
#define DUO_DECODE_MAX_Q_INT4 8
DUO_API int duo_decode_fused(const duo_layer* layer, const duo_cache_state* st, const void* qkv, int64_t qkv_row_stride,
                             const void* cos, const void* sin, int32_t rope_mode, void* out, int32_t q_len, float scale,
                             void* workspace, size_t workspace_bytes, void* stream);

/* Diagnostic twin of duo_attention that always takes the mma.sync (bandwidth) kernel family, also for
 * chunk shapes duo_attention hands to the tcgen05 prefill kernel.  16-bit KV only.  Used by the parity
 * tests to cross-check the two kernel families against each other. */
DUO_API int duo_attention_mma(const duo_layer* layer, const duo_cache_state* st, const void* q, int64_t q_row_stride,
                              void* out, int32_t q_len, float scale, void* workspace, size_t workspace_bytes,
                              void* stream);

/* device_state += n tokens: full_len += n, total += n, lo = max(lo, total - recent, sink) (one tiny kernel). */
DUO_API int duo_state_advance(int64_t* device_state, int32_t n, int32_t sink, int32_t recent, void* stream);
/* device_state := {full_len, total, lo}, stream-ordered (the three integers travel as kernel arguments, so repeated
 * evict_last()/clear() calls need no staging buffer and cannot race with earlier copies still in flight). */
DUO_API int duo_state_set(int64_t* device_state, int64_t full_len, int64_t total, int64_t lo, void* stream);

/* Move the tail of the staged chunk into sink/ring slots (call after duo_attention). */
DUO_API int duo_stream_commit(const duo_layer* layer, const duo_cache_state* st, int32_t q_len, void* stream);

/*
 * Stand-alone INT4 group quantisation, group_size == 128 == row length.
 *   in    : fp16 rows, row r at in + r*in_row_stride (elements)
 *   packed: [rows][64] u8, scale/zero: fp16 [rows]
 */
DUO_API int duo_quant_int4(const void* in, int64_t in_row_stride, int64_t rows, void* packed, void* scale,
                   void* zero, void* stream);
DUO_API int duo_dequant_int4(const void* packed, const void* scale, const void* zero, int64_t rows, void* out,
                     void* stream);

/*
 * Caller-side glue between the GEMMs of a decoder layer (the "next" row f3 of the scope table), HF arithmetic:
 *   duo_add_rmsnorm : h = residual + x (if residual != NULL; h is written to out_res if out_res != NULL),
 *                     out_norm = weight * T( h_fp32 * rsqrt(mean(h_fp32^2) + eps) )      [LlamaRMSNorm; replaces
 *                     flashinfer_rmsnorm_forward, duo_attn/patch/flashinfer_utils.py:9-16]
 *   duo_silu_mul    : out = T( T(silu(gate)) * up ) for gate_up = [rows][gate(inter) | up(inter)]
 * rows x hidden / rows x inter, contiguous, hidden and inter multiples of 8.
 */
DUO_API int duo_add_rmsnorm(const void* x, const void* residual, const void* weight, void* out_norm, void* out_res,
                            int64_t rows, int32_t hidden, float eps, int32_t dtype, void* stream);
DUO_API int duo_silu_mul(const void* gate_up, void* out, int64_t rows, int32_t inter, int32_t dtype, void* stream);

/*
 * Building blocks of the sequence-sharded decode: a retrieval head's
 * cache is split by position over several layers / ranks; each slice is attended separately and the slices are
 * combined with the online-softmax merge (what flash_attn_func computes over the whole cache in one call,
 * duo_attn/patch/llama.py:393-399, is recovered exactly up to fp32 rounding).
 *   duo_attention_partial : every query row (q as for duo_attention, already rotated, group * q_len <= 16) attends rows
 *       [0, n_keys) of EVERY retrieval head of `layer` (no causal offset, streaming heads are not computed);
 *       out_o  : fp32 [batch][q_len][n_q_heads][128]  normalised output of the slice (retrieval-head rows only)
 *       out_lse: fp32 [batch][q_len][n_q_heads]       log2-domain log-sum-exp: max*scale*log2(e) + log2(sum); -inf
 *                                                      for an empty slice
 *   duo_merge_partials    : out[tok][h][:] = sum_p 2^(lse_p - max) o_p / sum_p 2^(lse_p - max) for h < heads_used;
 *       o_parts fp32 [n_parts][tokens][heads_total][128], lse_parts fp32 [n_parts][tokens][heads_total],
 *       out: activation dtype [tokens][heads_total][128] (rows of heads >= heads_used are left untouched).
 */
DUO_API int duo_attention_partial(const duo_layer* layer, int64_t n_keys, const void* q, int64_t q_row_stride,
                                  float* out_o, float* out_lse, int32_t q_len, float scale, void* workspace,
                                  size_t workspace_bytes, void* stream);
/*
 * duo_attention_seq: one decode-sized chunk (group * q_len <= 16) of the sequence-sharded layout described at
 * duo_cache_state.  Retrieval q-heads attend the local slice (token t sees the local rows of positions
 * <= full_len + t) and report out_o / out_lse exactly as duo_attention_partial does; streaming q-heads (replicated on
 * every rank) are computed normally into `out` ([batch][q_len][n_q][128], rows of retrieval heads untouched).
 * Follow with duo_seq_merge, which exchanges the partials between the ranks and writes the retrieval rows of `out`.
 */
DUO_API int duo_attention_seq(const duo_layer* layer, const duo_cache_state* st, const void* q, int64_t q_row_stride,
                              void* out, float* out_o, float* out_lse, int32_t q_len, float scale, void* workspace,
                              size_t workspace_bytes, void* stream);
/* duo_decode_fused for a sequence-sharded cache, ONE new token per batch row: RoPE + append (by the owner of the position
 * only) + slice attention (partials as duo_attention_seq) + streaming heads + ring commit in one launch; follow with
 * duo_seq_merge. */
DUO_API int duo_decode_fused_seq(const duo_layer* layer, const duo_cache_state* st, const void* qkv, int64_t qkv_row_stride,
                                 const void* cos, const void* sin, int32_t rope_mode, void* out, float* out_o,
                                 float* out_lse, float scale, void* workspace, size_t workspace_bytes, void* stream);
DUO_API int duo_merge_partials(const float* o_parts, const float* lse_parts, int32_t n_parts, int64_t tokens,
                               int32_t heads_total, int32_t heads_used, void* out, int32_t dtype, void* stream);

/*
 * Head-parallel TP: the per-layer exchange step of the reference's
 * tensor-parallel sharding (duo_attn/utils.py:174-176, "sum" of the row-parallel o_proj / down_proj partials) as a
 * one-shot all-reduce over NVLink peer memory, fused with the residual add + RMSNorm that consumes it:
 *   out_res  = T( T(sum over ranks of partial) + residual )         (residual may be NULL: no add)
 *   out_norm = weight * T( out_res_fp32 * rsqrt(mean(out_res_fp32^2) + eps) )
 * rows <= max_rows (decode and small chunks; latency-bound sizes).  The caller owns all memory:
 *   data[r]  : rank r's receive buffer, duo_comm_data_bytes() bytes, zero-initialised, mapped on EVERY rank
 *              (CUDA IPC / torch symmetric memory; data[rank] is the local one); 16-byte aligned
 *   flags[r] : rank r's flag words, duo_comm_flag_bytes() bytes, zero-initialised, mapped on every rank
 *   local_state : local device int32[max_rows + 1], zero-initialised (call epochs; last word = error flag,
 *              set to 1 if a peer did not arrive within ~3 s - the result of that call is then undefined)
 * Every rank must issue the same sequence of calls (same rows) on one stream each.  Graph-capturable.
 */
typedef struct duo_comm duo_comm;
typedef struct duo_comm_desc {
  void* data[8];
  void* flags[8];
  void* local_state;
  int32_t rank, world;   /* 2 <= world <= 8 */
  int32_t hidden;        /* row length in elements, multiple of 8, <= 16384 */
  int32_t max_rows;      /* <= 64 */
  int32_t dtype;         /* DUO_DT_* */
} duo_comm_desc;
DUO_API size_t duo_comm_data_bytes(int32_t world, int32_t hidden, int32_t max_rows, int32_t dtype);
DUO_API size_t duo_comm_flag_bytes(int32_t world, int32_t max_rows);
DUO_API int duo_comm_create(const duo_comm_desc* desc, duo_comm** out);
DUO_API void duo_comm_destroy(duo_comm* comm);
DUO_API int duo_allreduce_add_rmsnorm(const duo_comm* comm, const void* partial, const void* residual,
                                      const void* weight, void* out_norm, void* out_res, int32_t rows, float eps,
                                      void* stream);

/*
 * duo_seq_merge: the exchange step of the sequence-sharded decode — one kernel that pushes this rank's (O, lse) partial
 * rows into every rank's receive slots over NVLink peer memory (same push / release-flag / acquire-wait protocol and
 * buffer ownership rules as duo_allreduce_add_rmsnorm), then merges the `world` partials of every row in rank order
 * (bit-identical on all ranks) with the online-softmax rule of duo_merge_partials and writes
 * out[tok][h][:] for h < heads_used in the activation dtype.  tokens * heads_used <= max_rows.
 *   part_o fp32 [tokens][heads_total][128], part_lse fp32 [tokens][heads_total] (as written by duo_attention_seq)
 *   data[r]: duo_seqcomm_data_bytes() bytes, flags[r]: duo_seqcomm_flag_bytes() bytes, zero-initialised, mapped on every
 *   rank; local_state: device int32[max_rows + 1] zero-initialised (epochs; last word = peer-timeout error flag).
 */
typedef struct duo_seqcomm duo_seqcomm;
typedef struct duo_seqcomm_desc {
  void* data[8];
  void* flags[8];
  void* local_state;
  int32_t rank, world;  /* 2 <= world <= 8 */
  int32_t max_rows;     /* <= 512 */
} duo_seqcomm_desc;
DUO_API size_t duo_seqcomm_data_bytes(int32_t world, int32_t max_rows);
DUO_API size_t duo_seqcomm_flag_bytes(int32_t world, int32_t max_rows);
DUO_API int duo_seqcomm_create(const duo_seqcomm_desc* desc, duo_seqcomm** out);
DUO_API void duo_seqcomm_destroy(duo_seqcomm* comm);
DUO_API int duo_seq_merge(const duo_seqcomm* comm, const float* part_o, const float* part_lse, void* out, int32_t tokens,
                          int32_t heads_total, int32_t heads_used, int32_t dtype, void* stream);

DUO_API const char* duo_last_error_string(void);
DUO_API int duo_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DUO_B200_H */
