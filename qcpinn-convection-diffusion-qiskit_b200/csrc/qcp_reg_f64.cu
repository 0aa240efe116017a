// qcpinn_b200 -- engine R kernels, float64 / complex128 instantiations (see qcp_reg.cuh).
// WV = 2: two warps per stream vector (six lane bits), used at n = 10.
#include "qcp_reg.cuh"

namespace qcp {
namespace rg {

template <>
int rg_launch<double>(int LB, int WV, int S, bool backward, const RgArgs& a, int grid, size_t smem, cudaStream_t s) {
#define RG_CALL(K, T, LBV, SV, WVV) rg_launch_one(&K<T, LBV, SV, WVV>, a, grid, rg_warps(SV, WVV) * 32, smem, s, #K)
  RG_INSTANTIATE(double, 4, 1)
  RG_INSTANTIATE(double, 4, 2)
#undef RG_CALL
  set_error("engine R: no float64 kernel for %d local bits / %d warps per vector", LB, WV);
  return 1;
}

template <>
int rg_occupancy<double>(int LB, int WV, int S, bool backward, size_t smem, int* blocks_per_sm) {
#define RG_CALL(K, T, LBV, SV, WVV) rg_occ_one(&K<T, LBV, SV, WVV>, rg_warps(SV, WVV) * 32, smem, blocks_per_sm)
  RG_INSTANTIATE(double, 4, 1)
  RG_INSTANTIATE(double, 4, 2)
#undef RG_CALL
  *blocks_per_sm = 0;
  return 1;
}

}  // namespace rg
}  // namespace qcp
