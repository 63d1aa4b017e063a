// K1 tensor-core engine (tcgen05 3xTF32) -- placeholder until the kernel lands.
#include "common.cuh"
namespace osc {
int knn_tc_supported(int64_t, int, int) { return 0; }
int launch_knn_tc(const float*, const float*, const float*, const float*, int64_t, int64_t, int64_t,
                  int64_t, int, int, int32_t*, float*, cudaStream_t) {
  return fail(OSC_ERR_UNSUPPORTED, "tensor-core kNN engine not built");
}
}  // namespace osc
