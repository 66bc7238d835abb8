// placeholder, replaced by the tcgen05 implementation
#include "common.cuh"
using namespace vspw;
extern "C" int vspw_conv2d_tc_supported(const vspw_conv_desc* d) { (void)d; return 0; }
extern "C" int vspw_conv2d_fwd_tc(const vspw_conv_desc*, const uint16_t*, const uint16_t*, const uint16_t*, const uint16_t*, const float*, float*, void*) { set_error("tc path not built"); return VSPW_ERR_UNSUPPORTED; }
extern "C" int vspw_conv2d_dgrad_tc(const vspw_conv_desc*, const uint16_t*, const uint16_t*, const uint16_t*, const uint16_t*, float*, void*) { set_error("tc path not built"); return VSPW_ERR_UNSUPPORTED; }
extern "C" int vspw_conv2d_wgrad_tc(const vspw_conv_desc*, const uint16_t*, const uint16_t*, const uint16_t*, const uint16_t*, float*, void*) { set_error("tc path not built"); return VSPW_ERR_UNSUPPORTED; }
