/* qcpinn_b200 -- C-ABI of the B200 (sm_100a) QCPINN training hot path.
 *
 * This is the drop-in boundary: the reference's hot path is pure Python that calls PennyLane's
 * ``default.qubit`` through a QNode (reference nn/DVQuantumLayer.py:143-154) and differentiates it
 * with nested torch autograd (reference nn/pde.py:53-72).  A maintainer of the reference would
 * bind these entry points with ctypes from ``DVQuantumLayer.forward`` /
 * ``DVPDESolver.forward`` / ``nn.pde.diffusion_operator``.
 *
 * Conventions
 *   - every function returns 0 on success; otherwise qcp_last_error() describes the failure
 *     (the Python host raises RuntimeError with that text).
 *   - all data pointers are DEVICE pointers borrowed from the caller (contiguous, 16-byte
 *     aligned); the library never frees or retains them beyond the call.  ``stream`` is a
 *     cudaStream_t passed as void* (torch's current stream); all work is stream-ordered.
 *   - ``dtype`` fixes the arithmetic AND the element type of every data pointer of a plan:
 *     QCP_F32 (float / complex64-equivalent) or QCP_F64 (double / complex128-equivalent).
 *   - a plan is not thread-safe; one host thread per process, one process per GPU.
 *   - there is NO CPU fallback: every entry point fails when no CUDA device is usable.
 */
#ifndef QCPINN_B200_H_
#define QCPINN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qcp_plan qcp_plan_t;

/* Gate kinds of the batch-shared circuit program.  Semantics follow PennyLane's public gate
 * definitions as used by reference nn/DVQuantumLayer.py:246-371 (wire 0 = most significant bit). */
enum qcp_gate_kind {
  QCP_GATE_RX = 0,   /* (wire, -, theta_index)            */
  QCP_GATE_RY = 1,
  QCP_GATE_RZ = 2,
  QCP_GATE_CRX = 3,  /* (control, target, theta_index)    */
  QCP_GATE_CRZ = 4,
  QCP_GATE_CNOT = 5, /* (control, target, -1)             */
  QCP_GATE_H = 6,    /* (wire, -, -1)                     */
  QCP_GATE_U4 = 7,   /* (wire_hi, wire_lo, const_index): fixed 4x4 unitary (the Haar blocks,
                        reference nn/DVQuantumLayer.py:203-209) */
  /* Per-SAMPLE gates inside the program ("jet gates"): the angle is scale * z[input], z = the
   * layer's per-point input vector, and the six Taylor streams are coupled through the gate's
   * derivatives.  They serve the data re-uploading circuit family of reference
   * hybrid_testing/CG_HQPINN_IBMtest_16qubits.py:217-235 (RY(x_i) encoding, RZ(0.5 x_j) re-upload
   * in every layer).  Fields: (wire, input index, scale in quarters: 4 = 1.0, 2 = 0.5). */
  QCP_GATE_RY_IN = 8,
  QCP_GATE_RZ_IN = 9,
  QCP_GATE_CZ = 10   /* (wire, wire, -1): diag(1, 1, 1, -1), reference :230-234 */
};

enum qcp_encoding {
  QCP_ENC_ANGLE = 0,     /* AngleEmbedding(rotation="X"), reference nn/DVQuantumLayer.py:182   */
  QCP_ENC_AMPLITUDE = 1, /* AmplitudeEmbedding(normalize=True, pad_with=0), reference :178-180 */
  QCP_ENC_NONE = 2       /* start from |0...0>: the program itself holds the per-sample gates */
};

enum qcp_dtype { QCP_F32 = 0, QCP_F64 = 1 };

enum qcp_mode {
  QCP_MODE_VALUE = 1,    /* u only            (DVPDESolver.forward, reference nn/DVPDESolver.py:81) */
  QCP_MODE_RESIDUAL = 6  /* u, u_t, u_x, u_y, u_xx, u_yy Taylor streams (reference nn/pde.py:53-72)  */
};

/* Pre / post MLP tensors of DVPDESolver (reference nn/DVPDESolver.py:28-51), torch layouts. */
typedef struct qcp_mlp {
  void* w1; /* [H,3]  preprocessor.0.weight  */
  void* b1; /* [H]    preprocessor.0.bias    */
  void* w2; /* [n,H]  preprocessor.2.weight  */
  void* b2; /* [n]    preprocessor.2.bias    */
  void* w3; /* [H,n]  postprocessor.0.weight */
  void* b3; /* [H]    postprocessor.0.bias   */
  void* w4; /* [1,H]  postprocessor.2.weight */
  void* b4; /* [1]    postprocessor.2.bias   */
} qcp_mlp_t;

const char* qcp_last_error(void);
int qcp_version(void);

/* Number of usable CUDA devices (0 => every other call fails). */
int qcp_device_count(void);

/* Plan = compiled circuit program + device workspaces for one (n, encoding, dtype, hidden).
 * ``ops`` is a HOST array [n_ops][4] of (kind, a, b, p); ``consts`` a HOST array of n_consts 4x4
 * complex128 matrices, row-major, interleaved (re, im).  Replaces the QNode construction of
 * reference nn/DVQuantumLayer.py:143-149. */
int qcp_plan_create(qcp_plan_t** out, int n_qubits, int encoding, int dtype, int hidden,
                    const int32_t* ops, int n_ops, const double* consts, int n_consts, int n_theta);
int qcp_plan_destroy(qcp_plan_t* plan);
int qcp_plan_num_features(const qcp_plan_t* plan);

/* Which kernels a plan runs on: QCP_ENGINE_FEATURE (n <= 4: observables pre-multiplied into a real
 * feature matrix), QCP_ENGINE_REGISTER (5 <= n <= 10: per-sample statevectors in
 * registers, one warp per Taylor stream), QCP_ENGINE_TILED (up to 16 qubits: per-sample statevectors
 * in an HBM/L2-resident slab, swept tile by tile through registers) or QCP_ENGINE_GLOBAL (the
 * gate-by-gate fallback, selected with QCP_ENGINE=L in the environment). */
enum qcp_engine { QCP_ENGINE_FEATURE = 0, QCP_ENGINE_GLOBAL = 1, QCP_ENGINE_REGISTER = 2, QCP_ENGINE_TILED = 3 };
int qcp_plan_engine(const qcp_plan_t* plan);

/* One-line description of the compiled plan (engine, layout, sweep / op counts) for logs and tests;
 * written to buf (HOST, NUL terminated, at most len bytes). */
int qcp_plan_describe(const qcp_plan_t* plan, char* buf, int len);

/* Element type of the caller-facing arrays of the solver entry points (X, u, r, streams, grad_u,
 * grad_r, grad_X).  Default = the plan dtype.  A QCP_F64 plan (n <= 4) may be switched to QCP_F32
 * I/O: arithmetic, weights, gradients and the saved jets stay float64, but the float32 tensors of
 * DVPDESolver (reference nn/DVPDESolver.py:96 casts to float32 anyway) are read and written
 * directly, with no cast kernels around the calls. */
int qcp_plan_set_io_dtype(qcp_plan_t* plan, int io_dtype);

/* Statevector engines (n >= 5): by default the forward of a call with a ``save`` workspace also keeps
 * the final psi streams there (2^n complex per stream and point, see qcp_solver_workspace_elems) and
 * the backward starts from them; enabled = 0 trades that memory for a recomputed forward. */
int qcp_plan_set_state_save(qcp_plan_t* plan, int enabled);

/* Evaluate the batch-shared part of the circuit for the current angles ``theta`` [n_theta]:
 * V(theta) = H_last . Haar . ansatz layers, O_i = V^dag Z_i V, and the real feature matrix C with
 * <Z_i>(z) = sum_s C[i,s] phi_s(z).  Must precede forward calls whenever theta changed. */
int qcp_prepare(qcp_plan_t* plan, const void* theta, void* stream);

/* Debug / test: copy the feature matrix to the HOST as double [n_qubits][num_features]. */
int qcp_feature_matrix(qcp_plan_t* plan, double* out_host, void* stream);

/* DVQuantumLayer.forward (reference nn/DVQuantumLayer.py:151-154): z [B,n] -> q [n,B]. */
int qcp_layer_forward(qcp_plan_t* plan, const void* z, long long batch, void* q, void* stream);

/* Its reverse mode: grad_q [n,B] -> grad_z [B,n] (may be NULL) and grad_theta [n_theta]. */
int qcp_layer_backward(qcp_plan_t* plan, const void* theta, const void* z, const void* grad_q,
                       long long batch, void* grad_z, void* grad_theta, void* stream);

/* Fused DVPDESolver forward (reference nn/DVPDESolver.py:81-110) and, in QCP_MODE_RESIDUAL, the
 * convection-diffusion residual of reference nn/pde.py:53-72 in Taylor mode:
 *   r = c[0] u_t + c[1] u_x + c[2] u_y + c[3] u_xx + c[4] u_yy        (coeffs = HOST double[5]).
 * X [B,3] -> u [B], r [B] (NULL in value mode), streams [B,6] (optional, may be NULL). */
int qcp_solver_forward(qcp_plan_t* plan, const qcp_mlp_t* weights, const void* X, long long batch,
                       int mode, const double* coeffs, void* u, void* r, void* streams,
                       void* save, void* stream);

/* Elements (of the plan dtype) of the optional ``save`` workspace of a (batch, mode) call:
 * 2 * n_qubits * mode * batch (+ (2 * hidden + 2 * n_qubits) * batch for float64 plans with n <= 4).  When ``save`` is given, the forward
 * stores the Taylor jets of the pre-MLP outputs z and of the expectation values q in it
 * (component-major, coalesced), in that case also the tanh values of both hidden layers and the
 * sin / cos of the encoding angles, and the
 * backward runs as three lean kernels over them instead of recomputing the forward. */
long long qcp_solver_workspace_elems(const qcp_plan_t* plan, long long batch, int mode);

/* Reverse mode of the above: given grad_u / grad_r [B] (either may be NULL) write (overwrite, not
 * accumulate) the gradient of every weight tensor, of theta [n_theta], and optionally of X [B,3]
 * (value mode only; NULL otherwise).  ``save`` is the workspace filled by the matching forward call
 * (it is consumed: overwritten with cotangents) or NULL to recompute the forward from X inside one
 * fused kernel.  Replaces loss.backward() through the nested-autograd graph (reference
 * trainer/diffusion_train.py:81). */
int qcp_solver_backward(qcp_plan_t* plan, const qcp_mlp_t* weights, const void* theta,
                        const void* X, const void* grad_u, const void* grad_r, long long batch,
                        int mode, const double* coeffs, void* save, const qcp_mlp_t* grads,
                        void* grad_theta, void* grad_X, void* stream);

/* The same reverse mode with a DEFERRED reduction (n <= 4): the three model calls of one train step
 * (reference trainer/diffusion_train.py:40-43) launch their adjoint kernels with _add() and share ONE
 * partial-sum reduction + ONE theta-gradient kernel in _finish(), which writes the summed gradients.
 * qcp_solver_backward() == _begin(); _add(); _finish(). */
int qcp_solver_backward_begin(qcp_plan_t* plan);
int qcp_solver_backward_add(qcp_plan_t* plan, const qcp_mlp_t* weights, const void* X,
                            const void* grad_u, const void* grad_r, long long batch, int mode,
                            const double* coeffs, void* save, void* grad_X, void* stream);
int qcp_solver_backward_finish(qcp_plan_t* plan, const void* theta, const qcp_mlp_t* grads,
                               void* grad_theta, void* stream);

/* Between two qcp_solver_backward_add() calls on DIFFERENT streams: make ``stream`` wait until the
 * post-MLP adjoint kernel of the call added last has finished (no-op if that call launched none).
 * The train step uses it to hold the IC/BC adjoints (low-priority side stream) back until the
 * residual chain's contraction adjoint is ready to start, so they fill that kernel's tail instead
 * of slipping in front of it. */
int qcp_solver_backward_after_post(qcp_plan_t* plan, void* stream);

/* Reverse mode for operators that are NONLINEAR in the Taylor streams (reference nn/pde.py:2-25,
 * navier_stokes_2D_operator: products like u * u_x): the caller differentiates its own residual
 * with respect to the six streams of qcp_solver_forward(..., streams) and passes
 * grad_streams [B,6] = d loss / d (u, u_t, u_x, u_y, u_xx, u_yy); outputs as in
 * qcp_solver_backward().  n <= 4; ``save`` as above (residual-mode workspace or NULL). */
int qcp_solver_backward_streams(qcp_plan_t* plan, const qcp_mlp_t* weights, const void* theta,
                                const void* X, const void* grad_streams, long long batch,
                                void* save, const qcp_mlp_t* grads, void* grad_theta, void* stream);

/* Sampler.sample() tail fused in one kernel (reference data/diffusion_dataset.py:12-38): maps
 * uniform random numbers rnd [n,3] (float32, from torch.rand) into the box lo_hi = HOST float[6]
 * (lo[3], hi[3]) -> X [n,3], and evaluates the analytic target y [n]: kind 0 = solution u,
 * kind 1 = forcing r(diffusion, v_x, v_y).  float32 like the reference. */
int qcp_sample_targets(const float* rnd, long long n, const float* lo_hi, int kind,
                       double diffusion, double v_x, double v_y, float* X, float* y, void* stream);

/* One MSELoss term of the train objective (reference trainer/diffusion_train.py:44-47) and its
 * cotangent in one launch: *loss_slot += mean((pred - target)^2) (DEVICE double, zeroed by the
 * caller), grad[i] = 2 * weight * (pred[i] - target[i]) / n.  float32 like the module tensors. */
int qcp_mse_seed(const float* pred, const float* target, long long n, double weight, float* grad,
                 double* loss_slot, void* stream);

/* Finish the flat float32 gradient [n_grad | n_extra scalars] of a data-parallel step in one launch:
 * scale everything by pre_scale (1 / world size after the sum all-reduce), then clip the gradient
 * part to max_norm exactly like torch.nn.utils.clip_grad_norm_ (reference :85). */
int qcp_clip_grads(float* flat, int n_grad, int n_extra, double pre_scale, double max_norm,
                   void* stream);

/* Gradient vector of the adjoint kernels (``dtype`` = the plan's, n_grad values in ``parameters()``
 * order) -> the optimizer's float32 flat buffer, with the objective behind it: flat[n_grad] =
 * w_r terms[0] + w_bc terms[1] + w_ic terms[2] (reference trainer/diffusion_train.py:48), flat[n_grad
 * + 1 .. + 3] = terms (DEVICE doubles written by qcp_mse_seed).  One launch instead of five. */
int qcp_pack_step(const void* grads, int dtype, int n_grad, const double* terms, double w_r,
                  double w_bc, double w_ic, float* flat, void* stream);

/* ``scheduler.step(loss)`` of the reference loop (trainer/diffusion_train.py:88-89;
 * torch.optim.lr_scheduler.ReduceLROnPlateau, nn/DVPDESolver.py:62) without a host round trip: one
 * thread repeats the scheduler's double arithmetic on DEVICE state = double[QCP_PLATEAU_STATE]
 * {best, num_bad_epochs, cooldown_counter, last_epoch, recorded, reductions}, multiplies the
 * device-resident float32 learning rate *lr by ``factor`` when patience runs out, and (history !=
 * NULL) appends *metric to history[recorded++] while recorded < history_cap (the loss history the
 * host drains later).  mode_max / threshold_abs select the scheduler's mode / threshold_mode. */
#define QCP_PLATEAU_STATE 6
int qcp_plateau_step(const float* metric, double* state, float* lr, float* history,
                     long long history_cap, int mode_max, int threshold_abs, double threshold,
                     double factor, long long patience, long long cooldown, double min_lr,
                     double eps, void* stream);

/* Data-parallel gradient exchange fused with the clip, over NVLink peer memory (one node; SURVEY.md
 * section 8e; the clip is reference trainer/diffusion_train.py:85).  ``peer_bufs`` = HOST array of
 * ``world`` DEVICE pointers, entry r = rank r's symmetric buffer of qcp_peer_allreduce_floats(n_grad +
 * n_extra, world) floats, zero-filled before the first call and mapped for peer access on this device
 * (torch.distributed._symmetric_memory).  One single-CTA launch per rank: store flat[0 .. n_grad +
 * n_extra) into every rank's buffer, exchange sequence-numbered flags (``seq`` = DEVICE counter, 0 at
 * the start, the same number of calls on every rank), sum the ranks' vectors in rank order (bit-
 * identical result on every rank), divide by world and clip the first n_grad values to max_norm like
 * torch.nn.utils.clip_grad_norm_ -- in place.  A peer that does not arrive within timeout_s sets
 * *error (DEVICE int, 1 + its rank) and the kernel returns. */
long long qcp_peer_allreduce_floats(int n_values, int world);
int qcp_peer_allreduce_clip(float* flat, int n_grad, int n_extra, const void* const* peer_bufs, int rank,
                            int world, unsigned int* seq, double max_norm, int* error,
                            double timeout_s, void* stream);

/* Host-only self check of the statevector-engine planners (no CUDA device needed; used by the CPU
 * tests): plans the gate list like qcp_plan_create() would for (n_qubits >= 5, dtype), runs the
 * logical circuit and the planned physical program on the CPU over a random statevector and
 * returns the largest amplitude difference.  consts = 4x4 complex128 matrices, theta = HOST doubles. */
int qcp_debug_check_plan(int n_qubits, int dtype, const int32_t* ops, int n_ops, const double* consts,
                         int n_consts, const double* theta, int n_theta, double* max_err,
                         int* n_phys_ops, int* n_sweeps);

/* FMA-pipe micro-benchmark used as the roofline denominator: runs
 * ``iters`` dependent-chain FMA rounds on every SM and returns achieved FLOP/s. */
int qcp_bench_fma(int dtype, int iters, double* flops_per_s, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* QCPINN_B200_H_ */
