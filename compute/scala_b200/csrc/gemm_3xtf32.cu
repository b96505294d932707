// gemm_3xtf32.cu — placeholder until the tcgen05 pipeline lands (next commit).
#include "builtin_kernels.h"
#include "common.h"
namespace cc {
bool gemm_available() { return false; }
int launch_gemm_3xtf32(const float*, const float*, float*, int64_t, int64_t, int64_t, const GemmWorkspace&, int, TensorMapEncodeFn, cudaStream_t) {
  fail(CC_ERR_UNSUPPORTED, "tcgen05 contraction not built");
}
}  // namespace cc
