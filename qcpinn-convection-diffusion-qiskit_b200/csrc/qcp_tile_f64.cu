// qcpinn_b200 -- engine T kernels, float64 / complex128 instantiations (see qcp_tile.cuh).
#include "qcp_tile.cuh"

namespace qcp {
namespace tl {

template <>
int tl_launch<double>(int LB, int S, bool backward, const TlArgs& a, int grid, size_t smem, cudaStream_t s) {
#define TL_CALL(K, T, LBV, SV, DGV, BW) tl_launch_one(&K<T, LBV, SV, DGV>, a, grid, tl_warps(BW) * 32, smem, s, #K)
  TL_INSTANTIATE(double, 4, false)
  TL_INSTANTIATE(double, 4, true)
#undef TL_CALL
  set_error("engine T: no float64 kernel for %d local bits", LB);
  return 1;
}

}  // namespace tl
}  // namespace qcp
