// placeholder: tcgen05 path not built yet
#include "nef_conv.cuh"
extern "C" int nef_gconv_fwd_simt(const NefConvDesc* d, nef_stream_t s);
extern "C" int nef_gconv_wgrad_simt(const NefWgradDesc* d, nef_stream_t s);
extern "C" int nef_gconv_fwd_tc(const NefConvDesc* d, nef_stream_t s) { return nef_gconv_fwd_simt(d, s); }
extern "C" int nef_gconv_wgrad_tc(const NefWgradDesc* d, nef_stream_t s) { return nef_gconv_wgrad_simt(d, s); }
extern "C" int nef_tc_init(void) { return 0; }

NEF_DEFINE_EXACT_SETTER(nef_set_exact_tc)
