// qcpinn_b200 -- float instantiation of the fused per-point kernels (see qcp_point.cuh).
#include "qcp_point.cuh"

namespace qcp {
template int launch_solver_forward<float>(int, int, int, const SolverArgs&, int, cudaStream_t);
template int launch_solver_backward<float>(int, int, int, const SolverArgs&, int, cudaStream_t);
template int launch_layer_forward<float>(int, int, const LayerArgs&, int, cudaStream_t);
template int launch_layer_backward<float>(int, int, const LayerArgs&, int, cudaStream_t);
template size_t solver_backward_smem<float>(int, int, int);
template int solver_split_grids<float>(int, int, int, int, int, SplitGrids*);
template int launch_solver_backward_split<float>(int, int, int, const SolverArgs&, const SplitGrids&, void*, void*, void*, cudaStream_t, cudaEvent_t, cudaEvent_t);
template int solver_backward_max_grid<float>(int, int, int, int, int);
}  // namespace qcp
