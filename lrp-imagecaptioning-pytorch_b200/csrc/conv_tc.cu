// placeholder until the tcgen05 kernel lands
#include "lrpx_common.cuh"
extern "C" int lrpx_tc_conv(const lrpx_tc_conv_args* args, void* stream) {
  lrpx::set_error("lrpx_tc_conv: not built yet");
  return LRPX_E_UNSUPPORTED;
}
