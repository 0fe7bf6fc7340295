// INT4-KV attention (placeholder)
#include "duo_common.cuh"
namespace duo {
int launch_attn_int4(const duo_layer*, const duo_cache_state*, const void*, long long, void*, int, float, void*, size_t,
                     cudaStream_t) {
  set_error("INT4 attention kernel not built");
  return DUO_EINVAL;
}
}  // namespace duo
