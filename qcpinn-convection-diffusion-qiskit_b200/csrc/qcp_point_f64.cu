// qcpinn_b200 -- double instantiation of the fused per-point kernels (see qcp_point.cuh).
#include "qcp_point.cuh"

namespace qcp {
template int launch_solver_forward<double>(int, int, int, const SolverArgs&, int, cudaStream_t);
template int launch_solver_backward<double>(int, int, int, const SolverArgs&, int, cudaStream_t);
template int launch_layer_forward<double>(int, int, const LayerArgs&, int, cudaStream_t);
template int launch_layer_backward<double>(int, int, const LayerArgs&, int, cudaStream_t);
template size_t solver_backward_smem<double>(int, int, int);
template int solver_split_grids<double>(int, int, int, int, int, SplitGrids*);
template int launch_solver_backward_split<double>(int, int, int, const SolverArgs&, const SplitGrids&, void*, void*, void*, cudaStream_t, cudaEvent_t, cudaEvent_t);
template int solver_backward_max_grid<double>(int, int, int, int, int);
}  // namespace qcp
