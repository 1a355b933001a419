#include "common.cuh"
#include "tc_conv.h"

namespace b200seg {
bool tc_conv_supported(const b200seg_conv_desc*, int) { return false; }
int tc_conv_run(const b200seg_conv_desc*, int, const void*, const void*, const float*, const void*,
                void*, cudaStream_t) {
  set_error("tcgen05 conv path not built");
  return B200SEG_ERR_UNSUPPORTED;
}
size_t tc_wgrad_extra_workspace(const b200seg_conv_desc*) { return 0; }
}  // namespace b200seg
