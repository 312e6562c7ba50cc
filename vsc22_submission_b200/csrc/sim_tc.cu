// Tensor-core similarity scoring (fp32-equivalent split-bf16 tcgen05 path).  Placeholder: reports
// "not handled" so that index.cu uses the exact SIMT kernel until this path lands.
#include "host_util.h"
#include "kernels.h"

namespace vscb200 {
int scores_tc(const float*, const float*, float*, int64_t, int64_t, int, int64_t, bool, const float*, const float*,
              cudaStream_t, bool* handled) {
  *handled = false;
  return VSCB200_OK;
}
}  // namespace vscb200
