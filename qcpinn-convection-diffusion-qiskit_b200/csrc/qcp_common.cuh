// qcpinn_b200 -- shared declarations of the sm_100a QCPINN hot path.
//
// Internal header (not part of the C-ABI; see include/qcpinn_b200.h for that).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/qcpinn_b200.h"

namespace qcp {

constexpr int kMaxQubitsFused = 4;   // engine S: observables pre-multiplied into a feature matrix
constexpr int kCStride = 4;          // feature-matrix row stride (outputs padded to 4 for LDS.128)
constexpr int kMaxHidden = 128;      // hidden width of the pre/post MLP held in shared memory
constexpr int kWarpsPerBlock = 4;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kMaxBlocksPerSm = 8;   // persistent grids never exceed this many blocks per SM
// The forward of a call with a workspace also saves the tanh values of both hidden layers and the
// split backward reads them back instead of evaluating tanh again: in float64 a tanh costs ~35
// FP64-pipe instructions (a third of an adjoint step per hidden unit), in float32 it is cheaper
// than the extra 8 H bytes per point of traffic (measured: float64 +5.6 % points/s, float32 -4 %).
template <typename T> struct SaveAct { static constexpr bool value = false; };
template <> struct SaveAct<double> { static constexpr bool value = true; };
// ... per mode: with one stream per point (value mode, S = 1) a hidden unit's adjoint is ~15 FMAs, so
// re-evaluating its tanh (16 FP64-pipe instructions with the exp table) is cheaper than 16 bytes
// of workspace traffic per unit and point (QCP_SAVE_ACT_VALUE=1 saves them like residual mode does)
#ifndef QCP_SAVE_ACT_VALUE
#define QCP_SAVE_ACT_VALUE 1
#endif
template <typename T, int S> struct SaveActS {
  static constexpr bool value = SaveAct<T>::value && (S != 1 || QCP_SAVE_ACT_VALUE != 0);
};
constexpr int kStageRows = 32;       // rows of the per-warp gradient staging tile
constexpr int kStagePitch = 33;      // +1 padding: column sums are bank-conflict free

// One gate of the batch-shared circuit program (ansatz layers, Haar blocks, final Hadamard).
struct GateOp {
  int32_t kind;   // qcp_gate_kind
  int32_t a;      // wire (1q) / control (controlled) / first wire (U4)
  int32_t b;      // target (controlled) / second wire (U4) / -1
  int32_t p;      // flat index into theta (rotations) or into consts (U4) or -1
};

struct PdeCoeffs {   // r = ct u_t + cx u_x + cy u_y + cxx u_xx + cyy u_yy   (nn/pde.py:60-71)
  double ct, cx, cy, cxx, cyy;
};

// Device-side view of everything the per-point kernels need.  Pointers are typed by the launch.
struct SolverArgs {
  const void* X;       // [B,3]
  const void* w1;      // [H,3]
  const void* b1;      // [H]
  const void* w2;      // [n,H]
  const void* b2;      // [n]
  const void* w3;      // [H,n]
  const void* b3;      // [H]
  const void* w4;      // [1,H]
  const void* b4;      // [1]
  const void* C;       // [F, kCStride] feature matrix (plan-owned)
  void* u;             // [B]        forward out
  void* r;             // [B]        forward out (residual mode) or null
  void* streams;       // [B,6]      optional: all Taylor streams of u (tests / eval) or null
  const void* gu;      // [B]        backward in (may be null => 0)
  const void* gr;      // [B]        backward in (may be null => 0)
  const void* gs;      // [B,6]      backward in: cotangent of all six Taylor streams (or null)
  void* gX;            // [B,3]      backward out or null
  void* partials;      // [grid, nacc] backward out (plan-owned)
  void* ws;            // saved-jet workspace [2][n*S][B] or null (forward: save; split backward: use)
  long long B;
  int H;
  int io_f32;          // 1: X, u, r, streams, gu, gr, gX are float32 whatever the plan dtype
  PdeCoeffs pde;
};

struct LayerArgs {     // stand-alone DVQuantumLayer
  const void* z;       // [B,n]
  const void* C;       // [F, kCStride]
  void* q;             // [n,B]  (reference orientation, nn/DVQuantumLayer.py:154)
  const void* gq;      // [n,B]
  void* gz;            // [B,n]
  void* partials;      // [grid, F*n]
  long long B;
};

__host__ __device__ constexpr int ipow3(int k) { return k <= 0 ? 1 : 3 * ipow3(k - 1); }

__host__ __device__ constexpr int num_features(int n, int enc) {
  return enc == QCP_ENC_AMPLITUDE ? n * (n + 1) / 2 : ipow3(n);
}

// accumulator segment sizes (also the staging / reduce order, see reduce_solver_kernel)
__host__ __device__ constexpr int nacc_post(int n, int H) { return 1 + H * (n + 2); }
__host__ __device__ constexpr int nacc_contract(int n, int enc) { return num_features(n, enc) * n; }
__host__ __device__ constexpr int nacc_pre(int n, int H) { return n + H * (4 + n); }

// Accumulator ("kernel order") layout of the backward kernel, see qcp_point.cuh.
__host__ __device__ constexpr int nacc_solver(int n, int enc, int H) {
  return 1 + H * (n + 2) + num_features(n, enc) * n + n + H * (4 + n);
}

struct SplitGrids {
  int post, contract, pre;
};

void set_error(const char* fmt, ...);

// per-dtype launchers (qcp_point_f32.cu / qcp_point_f64.cu)
template <typename T>
int launch_solver_forward(int n, int enc, int mode, const SolverArgs& a, int grid, cudaStream_t s);
template <typename T>
int launch_solver_backward(int n, int enc, int mode, const SolverArgs& a, int grid, cudaStream_t s);
template <typename T>
int launch_layer_forward(int n, int enc, const LayerArgs& a, int grid, cudaStream_t s);
template <typename T>
int launch_layer_backward(int n, int enc, const LayerArgs& a, int grid, cudaStream_t s);
template <typename T>
int solver_split_grids(int n, int enc, int mode, int H, int num_sms, SplitGrids* g);
template <typename T>
int launch_solver_backward_split(int n, int enc, int mode, const SolverArgs& a, const SplitGrids& g,
                                 void* part_post, void* part_contract, void* part_pre,
                                 cudaStream_t s, cudaEvent_t after_contract = nullptr,
                                 cudaEvent_t after_post = nullptr);
template <typename T>
size_t solver_backward_smem(int n, int enc, int H);
template <typename T>
int solver_backward_max_grid(int n, int enc, int mode, int H, int num_sms);

}  // namespace qcp
