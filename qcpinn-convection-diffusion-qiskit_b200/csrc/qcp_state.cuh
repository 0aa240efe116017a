// qcpinn_b200 -- host-side interface of engine L (qcp_state.cu, qcp_mlp.cu); internal header.
#pragma once

#include "qcp_common.cuh"

namespace qcp {

constexpr int kMaxQubitsSv = 16;      // largest statevector engine L handles (2^16 amplitudes)
constexpr int kMaxQubitsReg = 10;     // largest statevector engine R keeps in registers (qcp_reg.cuh)

struct SvLaunch {
  int n, enc, n_ops, n_theta, grid;
  const GateOp* ops;
  const double2* consts;
  const void* theta;
  void* ws;                 // saved-jet workspace [2][n*S][B]
  long long B;
  void* slab;               // per-CTA state storage in global memory, or null => shared memory
  double* theta_partials;   // [grid][n_theta]   (backward)
  void* grad_theta;         // [n_theta]          (backward)
};

size_t sv_state_bytes_rt(int dtype, int n, int S);
size_t sv_fixed_smem_rt(int dtype, int n, int S, int n_ops, int n_theta);
int sv_run(int dtype, int S, bool backward, const SvLaunch& L, cudaStream_t s);

// engine R (qcp_reg.cu): register-resident statevectors for 5 <= n <= 10 (float64 at n = 10: two warps per vector)
struct RegPlan;
int reg_supported(int n, int dtype);
RegPlan* reg_create(int n, int enc, int dtype, const GateOp* host_ops, int n_ops, int n_theta,
                    int n_consts, const GateOp* d_ops, const double2* d_consts, int num_sms);
void reg_destroy(RegPlan* r);
int reg_prepare(RegPlan* r, const void* d_theta, cudaStream_t s);       // phase tables of the diagonal blocks
int reg_describe(const RegPlan* r, char* buf, int len);
long long reg_state_elems(const RegPlan* r, long long B, int S);       // saved final psi, elements of T
int reg_run(RegPlan* r, int S, bool backward, void* ws, long long B, void* state, void* grad_theta,
            cudaStream_t s);

// engine T (qcp_tile.cu): tiled statevector sweeps for 11 <= n <= 16
struct TilePlan;
int tile_supported(int n, int dtype);
TilePlan* tile_create(int n, int enc, int dtype, const GateOp* host_ops, int n_ops, int n_theta,
                      int n_consts, const GateOp* d_ops, const double2* d_consts, int num_sms);
void tile_destroy(TilePlan* r);
int tile_prepare(TilePlan* r, const void* d_theta, cudaStream_t s);
int tile_num_sweeps(const TilePlan* r);
int tile_describe(const TilePlan* r, char* buf, int len);
long long tile_state_elems(const TilePlan* r, long long B, int S);     // saved final psi, elements of T
int tile_run(TilePlan* r, int S, bool backward, void* ws, long long B, void* state, void* grad_theta,
             cudaStream_t s);

// generic-n MLP stages (one thread per point, jets through the workspace)
struct MlpLaunch {
  int n, H, grid;
  SolverArgs args;          // X, weights, u/r/streams, gu/gr, gX, ws, B, pde; partials per stage
};
size_t mlp_smem_bytes(int dtype, int n, int H, int nacc);
int mlp_pre_forward(int dtype, int S, const MlpLaunch& L, cudaStream_t s);
int mlp_post_forward(int dtype, int S, const MlpLaunch& L, cudaStream_t s);
int mlp_post_backward(int dtype, int S, const MlpLaunch& L, cudaStream_t s);
int mlp_pre_backward(int dtype, int S, const MlpLaunch& L, cudaStream_t s);
int mlp_backward_grid(int dtype, int S, int n, int H, int num_sms);

}  // namespace qcp
