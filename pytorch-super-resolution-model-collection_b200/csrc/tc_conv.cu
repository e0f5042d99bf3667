// placeholder until the tcgen05 kernels land
#include "srb_common.cuh"
namespace srb {
bool tc_conv_supported(const Geom &, const T4 &, const T4 &, bool) { return false; }
size_t tc_conv_ws_bytes(const Geom &) { return 0; }
int tc_conv_gather(const Geom &, const T4 &, const float *, bool, const T4 &, const Epi &, void *, size_t, cudaStream_t) {
  set_error("tensor path not built"); return SRB_EUNSUPPORTED; }
bool tc_wgrad_supported(const Geom &, const T4 &, const T4 &) { return false; }
size_t tc_wgrad_ws_bytes(const Geom &) { return 0; }
int tc_conv_wgrad(const Geom &, const T4 &, const T4 &, float *, float *, float, int, void *, size_t, cudaStream_t) {
  set_error("tensor path not built"); return SRB_EUNSUPPORTED; }
}
