// tcgen05 / TMEM / TMA kernel family (placeholder until the kernels land).
#include "common.cuh"

namespace mlstm {
bool tensor_supported(const mlstm_b200_shape&) { return false; }
size_t tensor_workspace_bytes(const mlstm_b200_shape&, int) { return 0; }
int tensor_fw(const mlstm_b200_fw_args&, cudaStream_t) {
  set_error("tensor path not built");
  return MLSTM_B200_EUNSUPPORTED;
}
int tensor_bw(const mlstm_b200_bw_args&, cudaStream_t) {
  set_error("tensor path not built");
  return MLSTM_B200_EUNSUPPORTED;
}
}  // namespace mlstm
