// tcgen05 prefill kernel (placeholder until the UMMA path lands)
#include "duo_common.cuh"
namespace duo {
bool tc_prefill_supported(const duo_layer*, const duo_cache_state*, int) { return false; }
int launch_attn_tc(const duo_layer*, const duo_cache_state*, const void*, long long, void*, int, float, cudaStream_t) {
  set_error("tcgen05 prefill kernel not built");
  return DUO_EINVAL;
}
}  // namespace duo
