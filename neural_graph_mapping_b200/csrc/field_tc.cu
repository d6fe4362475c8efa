// fp16-operand tensor-core (tcgen05) field / fused render path -- placeholder until the kernel lands.
#include "common.cuh"

namespace ngm {

bool field_tc_supported(const NgmFieldDesc& fd, const char** why) {
  if (why) *why = "tcgen05 kernel not built in this revision";
  return false;
}
size_t field_tc_workspace_bytes(const NgmFieldDesc&, int) { return 0; }
int launch_field_fwd_tc(const NgmFieldFwdArgs&, cudaStream_t) {
  set_error("tcgen05 kernel not built in this revision");
  return NGM_ERR_UNSUPPORTED;
}
int launch_render_fused_tc(const NgmRenderArgs&, cudaStream_t) {
  set_error("tcgen05 kernel not built in this revision");
  return NGM_ERR_UNSUPPORTED;
}

}  // namespace ngm
