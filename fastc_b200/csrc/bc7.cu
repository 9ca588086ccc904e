// placeholder until the BC7 kernels land (fails loudly, no fallback)
#include "kernels.h"
namespace fastc {
cudaError_t bc7_upload_tables() { return cudaSuccess; }
void bc7_free_workspace(Bc7Workspace &) {}
cudaError_t launch_bc7(Bc7Workspace &, const void *, uint32_t, uint32_t, uint32_t, uint32_t, void *, int, uint64_t,
                       uint32_t, uint32_t, cudaStream_t, uint32_t *) {
  return cudaErrorNotSupported;
}
cudaError_t bc7_count_solid(Bc7Workspace &, const void *, uint32_t, uint32_t, uint32_t, cudaStream_t, uint32_t *) {
  return cudaErrorNotSupported;
}
cudaError_t bc7_read_counters(Bc7Workspace &, uint64_t *, uint64_t *) { return cudaErrorNotSupported; }
}  // namespace fastc
