// qcpinn_b200 -- fused per-collocation-point kernels ("engine S", n <= 4 qubits).
//
// One thread owns one collocation point and carries its Taylor jets in registers through
//   pre MLP (3 -> H -> n, tanh)                      reference nn/DVPDESolver.py:28-43
//   encoding features phi(z)                         reference nn/DVQuantumLayer.py:177-182
//   <Z_i> = sum_s C[i,s] phi_s(z)                    reference nn/DVQuantumLayer.py:184-214
//   post MLP (n -> H -> 1, tanh)                     reference nn/DVPDESolver.py:45-51
//   residual r = ct u_t + cx u_x + cy u_y + cxx u_xx + cyy u_yy     reference nn/pde.py:60-71
//
// The batch-shared part of the circuit (ansatz layers, Haar blocks, final Hadamard, the Z_i
// observables) was pre-multiplied by qcp_prepare() into the real feature matrix C (Heisenberg
// picture), so no per-sample statevector is needed: for angle encoding the input state is a
// product state whose density matrix is  (x)_j (I + y_j Y + w_j Z)/2  with y_j = -sin z_j,
// w_j = cos z_j, hence phi(z) = (x)_j (1, y_j, w_j) (3^n real features); for amplitude encoding
// phi_ab = f_a f_b / |f|^2 (n(n+1)/2 features).
//
// Backward, two ways:
//   * split (default for training): the forward saves the jets of z (pre-MLP output) and q (<Z_i>)
//     in a component-major workspace; post_backward / contract_backward / pre_backward then run as
//     three lean kernels, each with its own register budget, handing cotangent jets through the
//     same workspace.  The S = 6 contraction adjoint can run as two sub-jet passes (qcp_jet.cuh).
//   * fused: one kernel recomputes the forward from X (12 B/point), no workspace.
// Per-point parameter-gradient contributions are summed across the warp through a padded
// shared-memory staging tile, per block in shared accumulators, and across blocks by
// reduce_solver_kernel (deterministic: no atomics anywhere).
#pragma once

#include <type_traits>

#include "qcp_common.cuh"
#include "qcp_jet.cuh"

namespace qcp {

// ------------------------------------------------------------------------------------------
// occupancy knobs: min resident blocks per SM handed to __launch_bounds__ (tuned on B200 with
// tools/kbench.py; the register caps they imply are checked with -Xptxas -v)
// ------------------------------------------------------------------------------------------
// Encoded as decimal digits  FWD FUSED POST CONTRACT PRE TWOPASS ROLL  (one macro per kernel family
// so variants can be built with a single -D).
#ifndef QCP_TUNE_F64_6
#define QCP_TUNE_F64_6 1122310
#endif
#ifndef QCP_TUNE_F32_6
#define QCP_TUNE_F32_6 2332401
#endif
#ifndef QCP_TUNE_F64_1
#define QCP_TUNE_F64_1 3243400
#endif
#ifndef QCP_TUNE_F32_1
#define QCP_TUNE_F32_1 4444401
#endif

// roll the loop over the output qubit i inside the contraction adjoint (the body is ~170 FMAs; with
// i unrolled the float64 residual kernel is 21 k SASS instructions = 336 KB, far beyond the
// instruction caches: `no_instruction` was its second largest stall).  digits: F64 F32
// split contraction adjoint in two-pass mode: stream the sub-jet components through the workspace
#ifndef QCP_STREAM_PASSES
#define QCP_STREAM_PASSES 1
#endif
// block-wide barrier at every A-half iteration of the streamed contraction adjoint: the four warps
// of a block then run the (huge, fully unrolled) body together and share its instruction fetches
#ifndef QCP_LOCKSTEP
#define QCP_LOCKSTEP 0
#endif
// value mode (S = 1, the IC / BC points): points per thread of the post / pre MLP adjoints.  Their
// per-point arithmetic is ~15 FMAs per hidden unit, so with one point per thread the kernels were
// bound by the gradient staging (LDS + STS + DADD per parameter and point: LSU pipe 47 %, FP64 pipe
// 26 %, profiles/r02_*): K points per thread add their contributions in registers first and stage
// once, and the weight rows are fetched once for K points.
#ifndef QCP_VALUE_PPT
#define QCP_VALUE_PPT 4
#endif
#ifndef QCP_VALUE_PPT_CONTRACT
#define QCP_VALUE_PPT_CONTRACT 2      // the contraction adjoint carries more state per point
#endif
#ifndef QCP_PUT_AUTO
#define QCP_PUT_AUTO 1
#endif
#ifndef QCP_ROLL_I
#define QCP_ROLL_I 0
#endif
template <typename T>
struct RollI { static constexpr bool value = ((QCP_ROLL_I) % 10) != 0; };
template <>
struct RollI<double> { static constexpr bool value = (((QCP_ROLL_I) / 10) % 10) != 0; };

template <int V>
struct TuneValues {
  static constexpr int kFwd = (V / 1000000) % 10, kFused = (V / 100000) % 10,
                       kPost = (V / 10000) % 10, kContract = (V / 1000) % 10, kPre = (V / 100) % 10;
  // S = 6 contraction adjoint as two sub-jet passes (needed when 6-component jets overflow the
  // 255-register budget, i.e. in float64)
  static constexpr int kPasses = (V / 10) % 10;     // 0: one pass, 1: (t,x | y), 2: (t | x | y)
  static constexpr bool kTwoPass = kPasses != 0;
  // keep the loop over the A-half feature index rolled (9x less code, runtime-selected factors:
  // wins in float32 where the unrolled body thrashed the instruction cache; loses in float64)
  static constexpr int kRollMode = V % 10;   // 0: unrolled, 1: rolled, 2: outer trit rolled only
  static constexpr bool kRollA = kRollMode != 0;
};
template <typename T, int S>
struct Tune : TuneValues<1111100> {};
template <>
struct Tune<double, 6> : TuneValues<QCP_TUNE_F64_6> {};
template <>
struct Tune<float, 6> : TuneValues<QCP_TUNE_F32_6> {};
template <>
struct Tune<double, 1> : TuneValues<QCP_TUNE_F64_1> {};
template <>
struct Tune<float, 1> : TuneValues<QCP_TUNE_F32_1> {};

// ------------------------------------------------------------------------------------------
// feature-matrix image in shared memory
//   angle:     Ct[(a*NQ + i)*FBP + b] = C[i, a*FB + b]  (rows over b padded to a multiple of 4, so
//              one (a, i) row is a couple of LDS.128 broadcasts)
//   amplitude: C4[s][4]               = C[i, s]         (i padded to 4)
// ------------------------------------------------------------------------------------------
template <int NQ>
struct AngleShape {
  static constexpr int NA = (NQ + 1) / 2;
  static constexpr int NB = NQ - NA;
  static constexpr int FA = ipow3(NA);
  static constexpr int FB = ipow3(NB);
  static constexpr int FBP = (FB + 3) & ~3;
};

__host__ __device__ constexpr int c_elems(int n, int enc) {
  return enc == QCP_ENC_AMPLITUDE
             ? 4 * (n * (n + 1) / 2)
             : ipow3((n + 1) / 2) * n * ((ipow3(n - (n + 1) / 2) + 3) & ~3);
}

template <typename T, int NQ, int ENC>
__device__ __forceinline__ void load_C(T* sC, const T* Cg) {
  if constexpr (ENC == QCP_ENC_ANGLE) {
    using A = AngleShape<NQ>;
    for (int e = threadIdx.x; e < A::FA * NQ * A::FBP; e += blockDim.x) {
      const int b = e % A::FBP, i = (e / A::FBP) % NQ, a = e / (A::FBP * NQ);
      sC[e] = b < A::FB ? Cg[(a * A::FB + b) * kCStride + i] : T(0);
    }
  } else {
    for (int e = threadIdx.x; e < c_elems(NQ, ENC); e += blockDim.x) sC[e] = Cg[e];
  }
}

// ------------------------------------------------------------------------------------------
// shared-memory image of the MLP weights
// ------------------------------------------------------------------------------------------
template <typename T>
struct SmemWeights {
  Vec4<T>* w1b;   // [H]  (w1[k,0], w1[k,1], w1[k,2], b1[k])
  Vec4<T>* w2t;   // [H]  w2[j,k], j < n (zero padded)
  Vec4<T>* w3;    // [H]  w3[k,i], i < n (zero padded)
  T* b3w4;        // [H][2] (b3[k], w4[k])
  T* b2;          // [4]
  T* b4;          // [4]  (only [0] used)
  T* etab;        // [64] 2^(j/64): exp table of the float64 tanh (unused in float32)
  T* C;           // feature matrix image (layout above)
};

template <typename T>
__host__ __device__ inline size_t smem_weights_bytes(int H, int celems) {
  return sizeof(T) * (size_t)(4 * H * 3 + 2 * H + 4 + 4 + kExpTabSize + celems);
}

template <typename T>
__device__ __forceinline__ SmemWeights<T> carve_weights(unsigned char* base, int H, int celems) {
  SmemWeights<T> s;
  T* p = reinterpret_cast<T*>(base);
  s.w1b = reinterpret_cast<Vec4<T>*>(p); p += 4 * H;
  s.w2t = reinterpret_cast<Vec4<T>*>(p); p += 4 * H;
  s.w3 = reinterpret_cast<Vec4<T>*>(p);  p += 4 * H;
  s.C = p;    p += celems;
  s.b3w4 = p; p += 2 * H;
  s.b2 = p;   p += 4;
  s.b4 = p;   p += 4;
  s.etab = p; p += kExpTabSize;
  return s;
}

template <typename T, int NQ, int ENC>
__device__ __forceinline__ void load_weights(const SmemWeights<T>& s, const SolverArgs& a) {
  const int H = a.H;
  const T* w1 = static_cast<const T*>(a.w1);
  const T* b1 = static_cast<const T*>(a.b1);
  const T* w2 = static_cast<const T*>(a.w2);
  const T* b2 = static_cast<const T*>(a.b2);
  const T* w3 = static_cast<const T*>(a.w3);
  const T* b3 = static_cast<const T*>(a.b3);
  const T* w4 = static_cast<const T*>(a.w4);
  const T* b4 = static_cast<const T*>(a.b4);
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    Vec4<T> v;
    v.v[0] = w1[k * 3 + 0]; v.v[1] = w1[k * 3 + 1]; v.v[2] = w1[k * 3 + 2]; v.v[3] = b1[k];
    s.w1b[k] = v;
    Vec4<T> t, u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      t.v[j] = j < NQ ? w2[j * H + k] : T(0);
      u.v[j] = j < NQ ? w3[k * NQ + j] : T(0);
    }
    s.w2t[k] = t;
    s.w3[k] = u;
    s.b3w4[2 * k] = b3[k];
    s.b3w4[2 * k + 1] = w4[k];
  }
  load_C<T, NQ, ENC>(s.C, static_cast<const T*>(a.C));
  if (threadIdx.x < 4) {
    s.b2[threadIdx.x] = threadIdx.x < NQ ? b2[threadIdx.x] : T(0);
    s.b4[threadIdx.x] = b4[0];
  }
  if (threadIdx.x < kExpTabSize) s.etab[threadIdx.x] = (T)exp2((double)threadIdx.x / kExpTabSize);
}

// ------------------------------------------------------------------------------------------
// pre MLP:  z_j = b2_j + sum_k w2[j,k] tanh(b1_k + w1[k,:] . X)       (S = 1 or 6)
// ------------------------------------------------------------------------------------------
template <typename T, int S>
__device__ __forceinline__ Jet<T, S> pre_activation(const Vec4<T>& w, const T (&X)[3]) {
  Jet<T, S> a;
  a.c[0] = fma(w.v[0], X[0], fma(w.v[1], X[1], fma(w.v[2], X[2], w.v[3])));
  if constexpr (S == 6) {
    a.c[1] = w.v[0]; a.c[2] = w.v[1]; a.c[3] = w.v[2]; a.c[4] = T(0); a.c[5] = T(0);
  }
  return a;
}

// `act` (optional): the tanh values of the hidden units are stored at act[k * B] (the caller has
// already offset the pointer by the point index), so that the split backward does not have to
// evaluate tanh again (in float64 a tanh is ~35 FP64-pipe instructions, a third of an adjoint step)
template <typename T, int NQ, int S>
__device__ __forceinline__ void pre_forward(const SmemWeights<T>& s, int H, const T (&X)[3],
                                            Jet<T, S> (&z)[NQ], T* act = nullptr, long long B = 0) {
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    jzero(z[j]);
    z[j].c[0] = s.b2[j];
  }
#pragma unroll 2
  for (int k = 0; k < H; ++k) {
    const Vec4<T> w = s.w1b[k];
    const Vec4<T> w2 = s.w2t[k];
    const Jet<T, S> a = pre_activation<T, S>(w, X);
    const T h0 = Math<T>::tanh_tab(a.c[0], s.etab);
    if (act) act[(size_t)k * B] = h0;
    const T f1 = fma(-h0, h0, T(1));
    const T f2 = T(-2) * h0 * f1;
    const Jet<T, S> h = jfunc(a, h0, f1, f2);
#pragma unroll
    for (int j = 0; j < NQ; ++j) jaxpy(z[j], w2.v[j], h);
  }
}

// ------------------------------------------------------------------------------------------
// post MLP:  u = b4 + sum_k w4[k] tanh(b3_k + w3[k,:] . q)
// ------------------------------------------------------------------------------------------
template <typename T, int NQ, int S>
__device__ __forceinline__ void post_forward(const SmemWeights<T>& s, int H,
                                             const Jet<T, S> (&q)[NQ], Jet<T, S>& u,
                                             T* act = nullptr, long long B = 0) {
  jzero(u);
  u.c[0] = s.b4[0];
#pragma unroll 2
  for (int k = 0; k < H; ++k) {
    const Vec4<T> w3 = s.w3[k];
    const T b3 = s.b3w4[2 * k], w4 = s.b3w4[2 * k + 1];
    Jet<T, S> p;
    jzero(p);
    p.c[0] = b3;
#pragma unroll
    for (int i = 0; i < NQ; ++i) jaxpy(p, w3.v[i], q[i]);
    const T g0 = Math<T>::tanh_tab(p.c[0], s.etab);
    if (act) act[(size_t)k * B] = g0;
    const T f1 = fma(-g0, g0, T(1));
    const T f2 = T(-2) * g0 * f1;
    const Jet<T, S> g = jfunc(p, g0, f1, f2);
    jaxpy(u, w4, g);
  }
}

// ------------------------------------------------------------------------------------------
// angle-encoding features.  Qubits [0, NA) form the "A" half (index a), qubits [NA, NQ) the "B"
// half (index b); trit 0 = I (factor 1), 1 = Y (y_j), 2 = Z (w_j); a = 3*s_0 + s_1 etc.
// The B-half products Q[b] are kept; A-half products P_a are formed on the fly.
// ------------------------------------------------------------------------------------------
template <typename T, int NQ, int S>
struct AngleFeat {
  using A = AngleShape<NQ>;
  T sn[NQ], cs[NQ];
  Jet<T, S> y[NQ], w[NQ];   // y = -sin z (Pauli-Y Bloch component), w = cos z (Pauli-Z)
  Jet<T, S> Q[A::FB];       // Q[0] unused (== 1)
};

// e[3*s0 + s1] over two qubits (or e[s0] over one)
template <typename T, int S, int K>
__device__ __forceinline__ void prod_build(Jet<T, S>* e, const Jet<T, S>* y, const Jet<T, S>* w) {
  if constexpr (K == 1) {
    e[1] = y[0]; e[2] = w[0];
  } else if constexpr (K == 2) {
    e[1] = y[1]; e[2] = w[1]; e[3] = y[0]; e[6] = w[0];
    e[4] = jmul(y[0], y[1]); e[5] = jmul(y[0], w[1]);
    e[7] = jmul(w[0], y[1]); e[8] = jmul(w[0], w[1]);
  }
}

template <typename T, int S, int K>
__device__ __forceinline__ void prod_pull(const Jet<T, S>* eb, const Jet<T, S>* y,
                                          const Jet<T, S>* w, Jet<T, S>* yb, Jet<T, S>* wb) {
  if constexpr (K == 1) {
    jadd(yb[0], eb[1]); jadd(wb[0], eb[2]);
  } else if constexpr (K == 2) {
    jadd(yb[1], eb[1]); jadd(wb[1], eb[2]); jadd(yb[0], eb[3]); jadd(wb[0], eb[6]);
    jmul_pull_acc(yb[0], eb[4], y[1]); jmul_pull_acc(yb[1], eb[4], y[0]);
    jmul_pull_acc(yb[0], eb[5], w[1]); jmul_pull_acc(wb[1], eb[5], y[0]);
    jmul_pull_acc(wb[0], eb[7], y[1]); jmul_pull_acc(yb[1], eb[7], w[0]);
    jmul_pull_acc(wb[0], eb[8], w[1]); jmul_pull_acc(wb[1], eb[8], w[0]);
  }
}

// runtime-selected Bloch factor of one qubit: trit 0 -> 1, 1 -> y, 2 -> w.  `s` is warp-uniform
// (it comes from the rolled loop over the A-half feature index), so these selects never diverge.
template <typename T, int S>
__device__ __forceinline__ Jet<T, S> bloch_select(int s, const Jet<T, S>& y, const Jet<T, S>& w) {
  Jet<T, S> r;
#pragma unroll
  for (int c = 0; c < S; ++c) r.c[c] = s == 1 ? y.c[c] : (s == 2 ? w.c[c] : (c == 0 ? T(1) : T(0)));
  return r;
}

// A-half product P_a, formed on the fly inside the ROLLED loop over a (keeping that loop rolled
// cuts the kernel's code size ~9x: the unrolled version thrashed the instruction cache)
template <typename T, int NQ, int S>
struct PA {
  using A = AngleShape<NQ>;
  Jet<T, S> f0, f1, p;
  int s0, s1;
  __device__ __forceinline__ void set(const AngleFeat<T, NQ, S>& f, int a) {
    if constexpr (A::NA == 1) set2(f, a, 0);
    else set2(f, a / 3, a % 3);
  }
  // (s0, s1) = trits of the first / second A-half qubit (s1 ignored when the half has one qubit)
  __device__ __forceinline__ void set2(const AngleFeat<T, NQ, S>& f, int t0, int t1) {
    s0 = t0; s1 = t1;
    if constexpr (A::NA == 1) {
      p = bloch_select<T, S>(s0, f.y[0], f.w[0]);
    } else {
      f0 = bloch_select<T, S>(s0, f.y[0], f.w[0]);
      f1 = bloch_select<T, S>(s1, f.y[1], f.w[1]);
      p = jmul(f0, f1);
    }
  }
  // pull the cotangent of P_a into the A-half Bloch cotangents
  __device__ __forceinline__ void pull(const Jet<T, S>& pb, Jet<T, S>* yb, Jet<T, S>* wb) const {
    if constexpr (A::NA == 1) {
      if (s0 == 1) jadd(yb[0], pb);
      else if (s0 == 2) jadd(wb[0], pb);
    } else {
      Jet<T, S> b0, b1;
      jzero(b0); jzero(b1);
      jmul_pull_acc(b0, pb, f1);
      jmul_pull_acc(b1, pb, f0);
      if (s0 == 1) jadd(yb[0], b0);
      else if (s0 == 2) jadd(wb[0], b0);
      if (s1 == 1) jadd(yb[1], b1);
      else if (s1 == 2) jadd(wb[1], b1);
    }
  }
};

// loop over the A-half feature index a = 3*s0 + s1 (or a = s0 for a one-qubit half), rolled /
// unrolled per Tune<>::kRollMode; body(a, s0, s1)
struct NoOp { __device__ __forceinline__ void operator()() const {} };

template <typename T, int NQ, int S, int MODE = -1, typename Body = NoOp, typename After = NoOp>
__device__ __forceinline__ void for_each_a(Body&& body, After&& after_rolled = After{}) {
  // after_rolled(): runs at the end of every iteration of a ROLLED loop (lets the caller bring
  // loop-carried state, e.g. the staging row counter, back to a compile-time known value)
  using A = AngleShape<NQ>;
  constexpr int mode = MODE >= 0 ? MODE : Tune<T, (S == 1 ? 1 : 6)>::kRollMode;
  if constexpr (A::NA == 1) {
    if constexpr (mode == 1) {
#pragma unroll 1
      for (int a = 0; a < 3; ++a) { body(a, a, 0); after_rolled(); }
    } else {
#pragma unroll
      for (int a = 0; a < 3; ++a) body(a, a, 0);
    }
  } else if constexpr (mode == 1) {
#pragma unroll 1
    for (int a = 0; a < 9; ++a) { body(a, a / 3, a % 3); after_rolled(); }
  } else if constexpr (mode == 2) {
#pragma unroll 1
    for (int t0 = 0; t0 < 3; ++t0) {
#pragma unroll
      for (int t1 = 0; t1 < 3; ++t1) body(3 * t0 + t1, t0, t1);
      after_rolled();
    }
  } else {
#pragma unroll
    for (int a = 0; a < 9; ++a) body(a, a / 3, a % 3);
  }
}

template <typename T, int NQ>
__device__ __forceinline__ void angle_sincos(const T (&z0)[NQ], T (&sn)[NQ], T (&cs)[NQ]) {
#pragma unroll
  for (int j = 0; j < NQ; ++j) Math<T>::sincos_(z0[j], &sn[j], &cs[j]);
}

// Bloch jets and B-half products from z jets and precomputed sin / cos of their values
template <typename T, int NQ, int S>
__device__ __forceinline__ void angle_features(const Jet<T, S> (&z)[NQ], const T (&sn)[NQ],
                                               const T (&cs)[NQ], AngleFeat<T, NQ, S>& f) {
  using A = AngleShape<NQ>;
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    f.sn[j] = sn[j];
    f.cs[j] = cs[j];
    f.y[j] = jfunc(z[j], -sn[j], -cs[j], sn[j]);
    f.w[j] = jfunc(z[j], cs[j], -sn[j], -cs[j]);
  }
  prod_build<T, S, A::NB>(f.Q, f.y + A::NA, f.w + A::NA);
}

template <typename T, int NQ, int S>
__device__ __forceinline__ void angle_forward(const Jet<T, S> (&z)[NQ], AngleFeat<T, NQ, S>& f) {
  T z0[NQ], sn[NQ], cs[NQ];
#pragma unroll
  for (int j = 0; j < NQ; ++j) z0[j] = z[j].c[0];
  angle_sincos<T, NQ>(z0, sn, cs);
  angle_features<T, NQ, S>(z, sn, cs, f);
}

// coefficients C[i, a, 0..FB) of one (a, i) row
template <typename T, int NQ>
struct CRow {
  using A = AngleShape<NQ>;
  Vec4<T> v[A::FBP / 4];
  __device__ __forceinline__ void load(const T* sC, int a, int i) {
    const Vec4<T>* p = reinterpret_cast<const Vec4<T>*>(sC + (a * NQ + i) * A::FBP);
#pragma unroll
    for (int k = 0; k < A::FBP / 4; ++k) v[k] = p[k];
  }
  __device__ __forceinline__ T at(int b) const { return v[b >> 2].v[b & 3]; }
};

// t = sum_b C[i, a, b] Q[b]   (Q[0] == 1)
template <typename T, int NQ, int S>
__device__ __forceinline__ Jet<T, S> contract_row(const CRow<T, NQ>& c, const Jet<T, S>* Q) {
  using A = AngleShape<NQ>;
  Jet<T, S> t;
  jzero(t);
  t.c[0] = c.at(0);
#pragma unroll
  for (int b = 1; b < A::FB; ++b) jaxpy(t, c.at(b), Q[b]);
  return t;
}

template <typename T, int NQ, int S>
__device__ __forceinline__ void angle_contract(const T* sC, const AngleFeat<T, NQ, S>& f,
                                               Jet<T, S> (&q)[NQ]) {
  using A = AngleShape<NQ>;
#pragma unroll
  for (int i = 0; i < NQ; ++i) jzero(q[i]);
  auto body = [&](int a, int t0, int t1) {
    PA<T, NQ, S> pa;
    pa.set2(f, t0, t1);
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      CRow<T, NQ> c;
      c.load(sC, a, i);
      const Jet<T, S> t = contract_row<T, NQ, S>(c, f.Q);
      jmul_acc(q[i], pa.p, t);
    }
  };
  for_each_a<T, NQ, S>(body);
}

// ------------------------------------------------------------------------------------------
// amplitude-encoding features: phi_ab = f_a f_b / sum_j f_j^2, a <= b (row-major pair index)
// ------------------------------------------------------------------------------------------
template <typename T, int NQ, int S>
struct AmpFeat {
  static constexpr int F = NQ * (NQ + 1) / 2;
  Jet<T, S> nrm;       // sum f_j^2
  Jet<T, S> inv;       // 1 / nrm
  T i1, i2, i3;        // derivatives of 1/x at nrm
  Jet<T, S> phi[F];
};

__host__ __device__ constexpr int pair_index(int a, int b, int n) {
  return a * n - a * (a - 1) / 2 + (b - a);
}

template <typename T, int NQ, int S>
__device__ __forceinline__ void amp_forward(const Jet<T, S> (&z)[NQ], AmpFeat<T, NQ, S>& f) {
  jzero(f.nrm);
#pragma unroll
  for (int j = 0; j < NQ; ++j) jmul_acc(f.nrm, z[j], z[j]);
  const T r = T(1) / f.nrm.c[0];
  f.i1 = -r * r;
  f.i2 = T(2) * r * r * r;
  f.i3 = T(-6) * r * r * r * r;
  f.inv = jfunc(f.nrm, r, f.i1, f.i2);
#pragma unroll
  for (int a = 0; a < NQ; ++a)
#pragma unroll
    for (int b = a; b < NQ; ++b) f.phi[pair_index(a, b, NQ)] = jmul(jmul(z[a], z[b]), f.inv);
}

template <typename T, int NQ, int S>
__device__ __forceinline__ void amp_contract(const T* sC, const AmpFeat<T, NQ, S>& f,
                                             Jet<T, S> (&q)[NQ]) {
  const Vec4<T>* C4 = reinterpret_cast<const Vec4<T>*>(sC);
#pragma unroll
  for (int i = 0; i < NQ; ++i) jzero(q[i]);
#pragma unroll
  for (int s = 0; s < AmpFeat<T, NQ, S>::F; ++s) {
    const Vec4<T> c = C4[s];
#pragma unroll
    for (int i = 0; i < NQ; ++i) jaxpy(q[i], c.v[i], f.phi[s]);
  }
}

// ------------------------------------------------------------------------------------------
// warp-level gradient staging: every lane deposits its per-point contribution to a parameter in
// row `cnt` of a [rows][33] tile; flush() lets lane r sum row r (bank-conflict free thanks to the
// +1 pitch) into this warp's private accumulator array.  Accumulator order == push order.
// ------------------------------------------------------------------------------------------
template <typename T>
struct Stager {
  T* tile;
  T* acc;
  int lane;
  int cnt;    // rows currently staged (warp uniform)
  int base;   // accumulator index of row 0 (warp uniform)

  __device__ __forceinline__ void begin() { cnt = 0; base = 0; }
  __device__ __forceinline__ void put(T v) {
    tile[cnt * kStagePitch + lane] = v;
    ++cnt;
  }
  __device__ __forceinline__ void flush() {
    __syncwarp();
    if (lane < cnt) {
      const T* p = tile + lane * kStagePitch;
      T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        s0 += p[j]; s1 += p[j + 1]; s2 += p[j + 2]; s3 += p[j + 3];
      }
      acc[base + lane] += (s0 + s1) + (s2 + s3);
    }
    __syncwarp();
    base += cnt;
    cnt = 0;
  }
  // make room for `rows` more rows
  __device__ __forceinline__ void reserve(int rows) {
    if (cnt + rows > kStageRows) flush();
  }
  // put() that closes the tile exactly when it is full (in fully unrolled code the row counter is a
  // compile-time constant, so the test folds away): tiles of 32 rows instead of 27 in the
  // contraction adjoint, i.e. 11 instead of 12 row-sum rounds per pass
  __device__ __forceinline__ void put_auto(T v) {
    if (cnt == kStageRows) flush();
    put(v);
  }
};

// ------------------------------------------------------------------------------------------
// saved activations (workspace slot 2: act[2H][B], pre-MLP tanh values then post-MLP ones).  The
// reader keeps kActAhead values in flight so the HBM latency of act[k * B] hides behind the
// adjoint steps of the previous hidden units.
// ------------------------------------------------------------------------------------------
#ifndef QCP_ACT_AHEAD
#define QCP_ACT_AHEAD 4
#endif
constexpr int kActAhead = QCP_ACT_AHEAD;

template <typename T>
__device__ __forceinline__ void tanh_derivs_saved(T f0, T& f1, T& f2, T& f3) {
  f1 = fma(-f0, f0, T(1));
  f2 = T(-2) * f0 * f1;
  f3 = f1 * fma(T(6) * f0, f0, T(-2));
}

// for_hidden<T, SAVED>(H, act, B, body): body(k, saved tanh value of unit k) for k = 0..H-1;
// SAVED = false passes zeros that the body ignores (it evaluates tanh itself then)
template <typename T, bool SAVED, typename Body>
__device__ __forceinline__ void for_hidden(int H, const T* act, long long B, Body&& body) {
  if constexpr (!SAVED) {
    for (int k = 0; k < H; ++k) body(k, T(0));
    return;
  }
  T nxt[kActAhead];
#pragma unroll
  for (int j = 0; j < kActAhead; ++j) nxt[j] = j < H ? act[(size_t)j * B] : T(0);
  for (int k0 = 0; k0 < H; k0 += kActAhead) {
    T cur[kActAhead];
#pragma unroll
    for (int j = 0; j < kActAhead; ++j) {
      cur[j] = nxt[j];
      const int kn = k0 + kActAhead + j;
      nxt[j] = kn < H ? act[(size_t)kn * B] : T(0);
    }
#pragma unroll
    for (int j = 0; j < kActAhead; ++j)
      if (k0 + j < H) body(k0 + j, cur[j]);
  }
}

// ------------------------------------------------------------------------------------------
// reverse sweeps
// ------------------------------------------------------------------------------------------
template <typename T, int NQ, int S, bool SAVED = false>
__device__ __forceinline__ void post_backward(const SmemWeights<T>& s, int H,
                                              const Jet<T, S> (&q)[NQ], const Jet<T, S>& ub,
                                              Jet<T, S> (&qb)[NQ], Stager<T>& st,
                                              const T* act = nullptr, long long B = 0) {
#pragma unroll
  for (int i = 0; i < NQ; ++i) jzero(qb[i]);
  for_hidden<T, SAVED>(H, act, B, [&](int k, T saved) {
    const Vec4<T> w3 = s.w3[k];
    const T b3 = s.b3w4[2 * k], w4 = s.b3w4[2 * k + 1];
    Jet<T, S> p;
    jzero(p);
    p.c[0] = b3;
#pragma unroll
    for (int i = 0; i < NQ; ++i) jaxpy(p, w3.v[i], q[i]);
    T g0, f1, f2, f3;
    if constexpr (SAVED) { g0 = saved; tanh_derivs_saved(g0, f1, f2, f3); }
    else tanh_derivs(p.c[0], g0, f1, f2, f3);
    const Jet<T, S> g = jfunc(p, g0, f1, f2);
    Jet<T, S> gb;
#pragma unroll
    for (int c = 0; c < S; ++c) gb.c[c] = w4 * ub.c[c];
    Jet<T, S> pb;
    jzero(pb);
    jfunc_pull_acc(pb, gb, p, f1, f2, f3);
    st.reserve(NQ + 2);
#pragma unroll
    for (int i = 0; i < NQ; ++i) st.put(jdot(pb, q[i]));   // d w3[k,i]
    st.put(pb.c[0]);                                      // d b3[k]
    st.put(jdot(ub, g));                                  // d w4[k]
#pragma unroll
    for (int i = 0; i < NQ; ++i) jaxpy(qb[i], w3.v[i], pb);
  });
}

// qb[i] for a runtime (warp-uniform) i without indexing the register array
template <typename T, int NQ, int S>
__device__ __forceinline__ Jet<T, S> jet_pick(const Jet<T, S> (&v)[NQ], int i) {
  Jet<T, S> r = v[0];
#pragma unroll
  for (int k = 1; k < NQ; ++k) {
    if (i == k) r = v[k];
  }
  return r;
}

// Adjoint of features + contraction for one jet layout.  Pushes d C in (a, i, b) order and ADDS
// the pulled-back cotangent into zb (so sub-jet passes can accumulate).
template <typename T, int NQ, int S, int MODE = -1, bool LOCK = false>
__device__ __forceinline__ void angle_backward(const T* sC, const Jet<T, S> (&z)[NQ],
                                               const AngleFeat<T, NQ, S>& f,
                                               const Jet<T, S> (&qb)[NQ], Jet<T, S> (&zb)[NQ],
                                               Stager<T>& st) {
  using A = AngleShape<NQ>;
  Jet<T, S> yb[NQ], wb[NQ], Qb[A::FB];
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    jzero(yb[j]);
    jzero(wb[j]);
  }
#pragma unroll
  for (int b = 0; b < A::FB; ++b) jzero(Qb[b]);

  auto body = [&](int a, int t0, int t1) {
    if constexpr (LOCK) __syncthreads();
    PA<T, NQ, S> pa;
    pa.set2(f, t0, t1);
    Jet<T, S> pab;
    jzero(pab);
    auto step = [&](int i, const Jet<T, S>& qbi) {
      CRow<T, NQ> c;
      c.load(sC, a, i);
      // D = pullback of qb_i through the multiplication by P_a
      Jet<T, S> D;
      jzero(D);
      jmul_pull_acc(D, qbi, pa.p);
      const Jet<T, S> t = contract_row<T, NQ, S>(c, f.Q);
      jmul_pull_acc(pab, qbi, t);
      if constexpr (QCP_PUT_AUTO != 0 && sizeof(T) == 4) {
        // float32 (rolled A-half loop, runtime row counter anyway): full 32-row tiles, -1.6 %.
        // In float64 the same change made the unrolled kernel 70 % slower (the row counter stops
        // being a compile-time constant across the two passes), so it keeps reserve().
        st.put_auto(D.c[0]);                               // d C[i, a, 0]  (Q[0] == 1)
#pragma unroll
        for (int b = 1; b < A::FB; ++b) {
          st.put_auto(jdot(D, f.Q[b]));                    // d C[i, a, b]
          jaxpy(Qb[b], c.at(b), D);
        }
      } else {
        st.reserve(A::FB);
        st.put(D.c[0]);                                    // d C[i, a, 0]  (Q[0] == 1)
#pragma unroll
        for (int b = 1; b < A::FB; ++b) {
          st.put(jdot(D, f.Q[b]));                         // d C[i, a, b]
          jaxpy(Qb[b], c.at(b), D);
        }
      }
    };
    if constexpr (RollI<T>::value && S != 1) {
#pragma unroll 1
      for (int i = 0; i < NQ; ++i) step(i, jet_pick<T, NQ, S>(qb, i));   // i is warp uniform
    } else {
#pragma unroll
      for (int i = 0; i < NQ; ++i) step(i, qb[i]);
    }
    pa.pull(pab, yb, wb);
  };
  // rolled variants: close the staging tile at the end of every rolled iteration, so that inside
  // the iteration the row counter is a compile-time constant again (immediate-offset STS, no
  // per-put address arithmetic or reserve() branches)
  st.flush();
  for_each_a<T, NQ, S, MODE>(body, [&]() { st.flush(); });
  prod_pull<T, S, A::NB>(Qb, f.y + A::NA, f.w + A::NA, yb + A::NA, wb + A::NA);
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    // y = -sin z : f' = -cos, f'' = sin, f''' = cos ;  w = cos z : f' = -sin, f'' = -cos, f''' = sin
    jfunc_pull_acc(zb[j], yb[j], z[j], -f.cs[j], f.sn[j], f.cs[j]);
    jfunc_pull_acc(zb[j], wb[j], z[j], -f.sn[j], -f.cs[j], f.sn[j]);
  }
}

// Value mode (S = 1) with K points per thread: the d C contributions of the K points are added in
// registers before they are staged (same push order as angle_backward), the C rows are fetched once.
template <typename T, int NQ, int K>
__device__ __forceinline__ void angle_backward_value(const T* sC, const Jet<T, 1> (&z)[K][NQ],
                                                     const AngleFeat<T, NQ, 1> (&f)[K],
                                                     const Jet<T, 1> (&qb)[K][NQ],
                                                     Jet<T, 1> (&zb)[K][NQ], Stager<T>& st) {
  using A = AngleShape<NQ>;
  constexpr int S = 1;
  Jet<T, S> yb[K][NQ], wb[K][NQ], Qb[K][A::FB];
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      jzero(yb[k][j]);
      jzero(wb[k][j]);
    }
#pragma unroll
    for (int b = 0; b < A::FB; ++b) jzero(Qb[k][b]);
  }
  auto body = [&](int a, int t0, int t1) {
    PA<T, NQ, S> pa[K];
    Jet<T, S> pab[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      pa[k].set2(f[k], t0, t1);
      jzero(pab[k]);
    }
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      CRow<T, NQ> c;
      c.load(sC, a, i);
      T dc[A::FB];
#pragma unroll
      for (int b = 0; b < A::FB; ++b) dc[b] = T(0);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        Jet<T, S> D;
        jzero(D);
        jmul_pull_acc(D, qb[k][i], pa[k].p);
        const Jet<T, S> t = contract_row<T, NQ, S>(c, f[k].Q);
        jmul_pull_acc(pab[k], qb[k][i], t);
        dc[0] += D.c[0];
#pragma unroll
        for (int b = 1; b < A::FB; ++b) {
          dc[b] = fma(D.c[0], f[k].Q[b].c[0], dc[b]);
          jaxpy(Qb[k][b], c.at(b), D);
        }
      }
      st.reserve(A::FB);
#pragma unroll
      for (int b = 0; b < A::FB; ++b) st.put(dc[b]);          // d C[i, a, b]
    }
#pragma unroll
    for (int k = 0; k < K; ++k) pa[k].pull(pab[k], yb[k], wb[k]);
  };
  st.flush();
  for_each_a<T, NQ, S>(body, [&]() { st.flush(); });
#pragma unroll
  for (int k = 0; k < K; ++k) {
    prod_pull<T, S, A::NB>(Qb[k], f[k].y + A::NA, f[k].w + A::NA, yb[k] + A::NA, wb[k] + A::NA);
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      jfunc_pull_acc(zb[k][j], yb[k][j], z[k][j], -f[k].cs[j], f[k].sn[j], f[k].cs[j]);
      jfunc_pull_acc(zb[k][j], wb[k][j], z[k][j], -f[k].sn[j], -f[k].cs[j], f[k].sn[j]);
    }
  }
}

template <typename T, int NQ, int S>
__device__ __forceinline__ void amp_backward(const T* sC, const Jet<T, S> (&z)[NQ],
                                             const AmpFeat<T, NQ, S>& f,
                                             const Jet<T, S> (&qb)[NQ], Jet<T, S> (&zb)[NQ],
                                             Stager<T>& st) {
  const Vec4<T>* C4 = reinterpret_cast<const Vec4<T>*>(sC);
  Jet<T, S> invb;
  jzero(invb);
#pragma unroll
  for (int a = 0; a < NQ; ++a) {
#pragma unroll
    for (int b = a; b < NQ; ++b) {
      const int s = pair_index(a, b, NQ);
      const Vec4<T> c = C4[s];
      Jet<T, S> phib;
      jzero(phib);
      st.reserve(NQ);
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        st.put(jdot(qb[i], f.phi[s]));                     // d C[i, (a,b)]
        jaxpy(phib, c.v[i], qb[i]);
      }
      const Jet<T, S> m = jmul(z[a], z[b]);
      Jet<T, S> mb;
      jzero(mb);
      jmul_pull_acc(mb, phib, f.inv);
      jmul_pull_acc(invb, phib, m);
      jmul_pull_acc(zb[a], mb, z[b]);
      jmul_pull_acc(zb[b], mb, z[a]);
    }
  }
  Jet<T, S> nb;
  jzero(nb);
  jfunc_pull_acc(nb, invb, f.nrm, f.i1, f.i2, f.i3);
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    jmul_pull_acc(zb[j], nb, z[j]);
    jmul_pull_acc(zb[j], nb, z[j]);
  }
}

// features + contraction adjoint of one point; zb is overwritten
template <typename T, int NQ, int ENC, int S>
__device__ __forceinline__ void feature_backward(const T* sC, const Jet<T, S> (&z)[NQ],
                                                 const Jet<T, S> (&qb)[NQ], Jet<T, S> (&zb)[NQ],
                                                 Stager<T>& st) {
#pragma unroll
  for (int j = 0; j < NQ; ++j) jzero(zb[j]);
  if constexpr (ENC == QCP_ENC_AMPLITUDE) {
    AmpFeat<T, NQ, S> f;
    amp_forward<T, NQ, S>(z, f);
    amp_backward<T, NQ, S>(sC, z, f, qb, zb, st);
  } else if constexpr (S == 6 && Tune<T, S>::kTwoPass) {
    // register-light passes over sub-jets, e.g. X = (u, u_t | u_x | u_xx), Y = (u | u_y | u_yy).
    // Every pullback is linear in the cotangent, so the value-component cotangent may be carried
    // by the first pass alone; the partial parameter-gradient pushes of all passes add up in acc.
    T z0[NQ], sn[NQ], cs[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) z0[j] = z[j].c[0];
    angle_sincos<T, NQ>(z0, sn, cs);
    st.flush();                  // all passes must start at the same accumulator index
    const int base0 = st.base;
    // one pass over the full-jet components idx[0..SS) (idx[0] == 0 is the value)
    auto pass = [&](auto tag, const int (&idx)[decltype(tag)::value], bool carry_value) {
      constexpr int SS = decltype(tag)::value;
      Jet<T, SS> zs[NQ], qs[NQ], zbs[NQ];
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
#pragma unroll
        for (int c = 0; c < SS; ++c) {
          zs[j].c[c] = z[j].c[idx[c]];
          qs[j].c[c] = qb[j].c[idx[c]];
        }
        if (!carry_value) qs[j].c[0] = T(0);
        jzero(zbs[j]);
      }
      AngleFeat<T, NQ, SS> f;
      angle_features<T, NQ, SS>(zs, sn, cs, f);
      angle_backward<T, NQ, SS>(sC, zs, f, qs, zbs, st);
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        zb[j].c[0] += zbs[j].c[0];
#pragma unroll
        for (int c = 1; c < SS; ++c) zb[j].c[idx[c]] = zbs[j].c[c];
      }
      st.flush();
      st.base = base0;           // the next pass adds onto the same accumulator segment
    };
    if constexpr (Tune<T, S>::kPasses == 1) {
      const int ix[4] = {0, 1, 2, 4}, iy[3] = {0, 3, 5};
      pass(std::integral_constant<int, 4>{}, ix, true);
      pass(std::integral_constant<int, 3>{}, iy, false);
    } else {
      const int it[2] = {0, 1}, ix[3] = {0, 2, 4}, iy[3] = {0, 3, 5};
      pass(std::integral_constant<int, 3>{}, ix, true);
      pass(std::integral_constant<int, 3>{}, iy, false);
      pass(std::integral_constant<int, 2>{}, it, false);
    }
    st.base = base0 + nacc_contract(NQ, ENC);
  } else {
    AngleFeat<T, NQ, S> f;
    angle_forward<T, NQ, S>(z, f);
    angle_backward<T, NQ, S>(sC, z, f, qb, zb, st);
  }
}

// Two-pass S = 6 contraction adjoint of the SPLIT kernel, streaming through the workspace: each
// sub-jet pass loads just its own components of z (slot 0) and qb (slot 1) and stores its own
// components of zb (slot 0) right away, so no full 6-component jets stay live across the passes
// (the register-array version keeps z, qb and zb = 72 doubles alive and spills ~1 KB per thread).
// Slot 0 is overwritten component by component: a pass only overwrites derivative components it
// has loaded itself; the value component (needed by every pass for nothing but sin / cos, which
// are taken first) is written by the last pass from the sum of the partial value cotangents.
template <typename T, int NQ, int MODE = -1>
__device__ __forceinline__ void angle_backward_ws(const T* sC, T* ws, long long B, long long p,
                                                  bool valid, Stager<T>& st, const T* trig = nullptr) {
  constexpr int S = 6;
  T* zrow = ws + p;                                   // slot 0, component-major
  const T* qrow = ws + (size_t)NQ * S * B + p;        // slot 1
  T sn[NQ], cs[NQ], zb0[NQ];
  if (SaveAct<T>::value && trig) {
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      sn[j] = trig[(size_t)j * B];
      cs[j] = trig[(size_t)(NQ + j) * B];
      zb0[j] = T(0);
    }
  } else {
    T z0[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) { z0[j] = zrow[(size_t)(j * S) * B]; zb0[j] = T(0); }
    angle_sincos<T, NQ>(z0, sn, cs);
  }
  st.flush();                  // all passes must start at the same accumulator index
  const int base0 = st.base;
  auto pass = [&](auto tag, const int (&idx)[decltype(tag)::value], bool carry_value, bool last) {
    constexpr int SS = decltype(tag)::value;
    Jet<T, SS> zs[NQ], qs[NQ], zbs[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
#pragma unroll
      for (int c = 0; c < SS; ++c) {
        zs[j].c[c] = zrow[(size_t)(j * S + idx[c]) * B];
        qs[j].c[c] = valid ? qrow[(size_t)(j * S + idx[c]) * B] : T(0);
      }
      if (!carry_value) qs[j].c[0] = T(0);
      jzero(zbs[j]);
    }
    AngleFeat<T, NQ, SS> f;
    angle_features<T, NQ, SS>(zs, sn, cs, f);
    angle_backward<T, NQ, SS, MODE, (QCP_LOCKSTEP != 0)>(sC, zs, f, qs, zbs, st);
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      zb0[j] += zbs[j].c[0];
      if (valid) {
#pragma unroll
        for (int c = 1; c < SS; ++c) zrow[(size_t)(j * S + idx[c]) * B] = zbs[j].c[c];
        if (last) zrow[(size_t)(j * S) * B] = zb0[j];
      }
    }
    st.flush();
    st.base = base0;           // the next pass adds onto the same accumulator segment
  };
  if constexpr (Tune<T, S>::kPasses == 1) {
    const int ix[4] = {0, 1, 2, 4}, iy[3] = {0, 3, 5};
    pass(std::integral_constant<int, 4>{}, ix, true, false);
    pass(std::integral_constant<int, 3>{}, iy, false, true);
  } else {
    const int it[2] = {0, 1}, ix[3] = {0, 2, 4}, iy[3] = {0, 3, 5};
    pass(std::integral_constant<int, 3>{}, ix, true, false);
    pass(std::integral_constant<int, 3>{}, iy, false, false);
    pass(std::integral_constant<int, 2>{}, it, false, true);
  }
  st.base = base0 + nacc_contract(NQ, QCP_ENC_ANGLE);
}

// `trig` (optional, angle encoding): sin z_j / cos z_j are stored at trig[j * B] / trig[(NQ + j) * B]
// (pointer already offset by the point), so the contraction adjoint need not evaluate four
// double-precision sincos per point again (workspace slot 3, float64 plans: SaveAct<T>)
template <typename T, int NQ, int ENC, int S>
__device__ __forceinline__ void feature_forward(const T* sC, const Jet<T, S> (&z)[NQ],
                                                Jet<T, S> (&q)[NQ], T* trig = nullptr,
                                                long long B = 0) {
  if constexpr (ENC == QCP_ENC_ANGLE) {
    AngleFeat<T, NQ, S> f;
    angle_forward<T, NQ, S>(z, f);
    if (trig) {
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        trig[(size_t)j * B] = f.sn[j];
        trig[(size_t)(NQ + j) * B] = f.cs[j];
      }
    }
    angle_contract<T, NQ, S>(sC, f, q);
  } else {
    AmpFeat<T, NQ, S> f;
    amp_forward<T, NQ, S>(z, f);
    amp_contract<T, NQ, S>(sC, f, q);
  }
}

template <typename T, int NQ, int S, bool SAVED = false>
__device__ __forceinline__ void pre_backward(const SmemWeights<T>& s, int H, const T (&X)[3],
                                             const Jet<T, S> (&zb)[NQ], T (&Xb)[3],
                                             Stager<T>& st, const T* act = nullptr,
                                             long long B = 0) {
  st.reserve(NQ);
#pragma unroll
  for (int j = 0; j < NQ; ++j) st.put(zb[j].c[0]);          // d b2[j]
  Xb[0] = Xb[1] = Xb[2] = T(0);
  for_hidden<T, SAVED>(H, act, B, [&](int k, T saved) {
    const Vec4<T> w = s.w1b[k];
    const Vec4<T> w2 = s.w2t[k];
    const Jet<T, S> a = pre_activation<T, S>(w, X);
    T h0, f1, f2, f3;
    if constexpr (SAVED) { h0 = saved; tanh_derivs_saved(h0, f1, f2, f3); }
    else tanh_derivs(a.c[0], h0, f1, f2, f3);
    const Jet<T, S> h = jfunc(a, h0, f1, f2);
    Jet<T, S> hb;
    jzero(hb);
#pragma unroll
    for (int j = 0; j < NQ; ++j) jaxpy(hb, w2.v[j], zb[j]);
    Jet<T, S> ab;
    jzero(ab);
    jfunc_pull_acc(ab, hb, a, f1, f2, f3);
    st.reserve(4 + NQ);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      T g = ab.c[0] * X[d];
      if constexpr (S == 6) g += ab.c[1 + d];               // a.c[1+d] IS w1[k,d]
      st.put(g);                                           // d w1[k,d]
      Xb[d] = fma(ab.c[0], w.v[d], Xb[d]);
    }
    st.put(ab.c[0]);                                       // d b1[k]
#pragma unroll
    for (int j = 0; j < NQ; ++j) st.put(jdot(zb[j], h));    // d w2[j,k]
  });
}

// ------------------------------------------------------------------------------------------
// value mode with K points per thread (S = 1: plain scalars, no jets).  Accumulator order is the
// one of post_backward / pre_backward, so the reduction and the fused kernel see no difference.
// `act` points at this MLP's saved tanh row 0 (unused when !SAVED); pts[] are the point indices.
// ------------------------------------------------------------------------------------------
template <typename T, int NQ, int K, bool SAVED>
__device__ __forceinline__ void post_backward_value(const SmemWeights<T>& s, int H,
                                                    const T (&q)[K][NQ], const T (&ub)[K],
                                                    T (&qb)[K][NQ], Stager<T>& st, const T* act,
                                                    long long B, const long long (&pts)[K]) {
#pragma unroll
  for (int j = 0; j < K; ++j)
#pragma unroll
    for (int i = 0; i < NQ; ++i) qb[j][i] = T(0);
  T nxt[K];
#pragma unroll
  for (int j = 0; j < K; ++j) nxt[j] = SAVED ? act[pts[j]] : T(0);
  for (int k = 0; k < H; ++k) {
    const Vec4<T> w3 = s.w3[k];
    const T b3 = s.b3w4[2 * k], w4 = s.b3w4[2 * k + 1];
    T g0[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      g0[j] = nxt[j];
      if constexpr (SAVED) {
        if (k + 1 < H) nxt[j] = act[(size_t)(k + 1) * B + pts[j]];
      } else {
        T pre = b3;
#pragma unroll
        for (int i = 0; i < NQ; ++i) pre = fma(w3.v[i], q[j][i], pre);
        g0[j] = Math<T>::tanh_tab(pre, s.etab);
      }
    }
    T dw3[NQ], db3 = T(0), dw4 = T(0);
#pragma unroll
    for (int i = 0; i < NQ; ++i) dw3[i] = T(0);
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const T f1 = fma(-g0[j], g0[j], T(1));
      const T pb = w4 * ub[j] * f1;
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        dw3[i] = fma(pb, q[j][i], dw3[i]);
        qb[j][i] = fma(w3.v[i], pb, qb[j][i]);
      }
      db3 += pb;
      dw4 = fma(ub[j], g0[j], dw4);
    }
    st.reserve(NQ + 2);
#pragma unroll
    for (int i = 0; i < NQ; ++i) st.put(dw3[i]);             // d w3[k,i]
    st.put(db3);                                           // d b3[k]
    st.put(dw4);                                           // d w4[k]
  }
}

template <typename T, int NQ, int K, bool SAVED>
__device__ __forceinline__ void pre_backward_value(const SmemWeights<T>& s, int H,
                                                   const T (&X)[K][3], const T (&zb)[K][NQ],
                                                   T (&Xb)[K][3], Stager<T>& st, const T* act,
                                                   long long B, const long long (&pts)[K]) {
  st.reserve(NQ);
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    T sum = T(0);
#pragma unroll
    for (int j = 0; j < K; ++j) sum += zb[j][i];
    st.put(sum);                                           // d b2[i]
  }
#pragma unroll
  for (int j = 0; j < K; ++j) Xb[j][0] = Xb[j][1] = Xb[j][2] = T(0);
  T nxt[K];
#pragma unroll
  for (int j = 0; j < K; ++j) nxt[j] = SAVED ? act[pts[j]] : T(0);
  for (int k = 0; k < H; ++k) {
    const Vec4<T> w = s.w1b[k];
    const Vec4<T> w2 = s.w2t[k];
    T h0[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      h0[j] = nxt[j];
      if constexpr (SAVED) {
        if (k + 1 < H) nxt[j] = act[(size_t)(k + 1) * B + pts[j]];
      } else {
        const T pre = fma(w.v[0], X[j][0], fma(w.v[1], X[j][1], fma(w.v[2], X[j][2], w.v[3])));
        h0[j] = Math<T>::tanh_tab(pre, s.etab);
      }
    }
    T dw1[3] = {T(0), T(0), T(0)}, db1 = T(0), dw2[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) dw2[i] = T(0);
#pragma unroll
    for (int j = 0; j < K; ++j) {
      T hb = T(0);
#pragma unroll
      for (int i = 0; i < NQ; ++i) hb = fma(w2.v[i], zb[j][i], hb);
      const T ab = hb * fma(-h0[j], h0[j], T(1));
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        dw1[d] = fma(ab, X[j][d], dw1[d]);
        Xb[j][d] = fma(ab, w.v[d], Xb[j][d]);
      }
      db1 += ab;
#pragma unroll
      for (int i = 0; i < NQ; ++i) dw2[i] = fma(zb[j][i], h0[j], dw2[i]);
    }
    st.reserve(4 + NQ);
#pragma unroll
    for (int d = 0; d < 3; ++d) st.put(dw1[d]);              // d w1[k,d]
    st.put(db1);                                           // d b1[k]
#pragma unroll
    for (int i = 0; i < NQ; ++i) st.put(dw2[i]);             // d w2[i,k]
  }
}

// ------------------------------------------------------------------------------------------
// saved-jet workspace: ws[slot][j*S + c][B] (component-major => coalesced across the warp).
// slot 0: z jets (pre-MLP output), overwritten by their cotangents zb;
// slot 1: q jets (<Z_i> streams),  overwritten by their cotangents qb.
// ------------------------------------------------------------------------------------------
template <typename T, int NQ, int S>
__device__ __forceinline__ void ws_store(T* ws, long long B, int slot, long long p,
                                         const Jet<T, S> (&v)[NQ]) {
  T* base = ws + (size_t)slot * NQ * S * B + p;
#pragma unroll
  for (int j = 0; j < NQ; ++j)
#pragma unroll
    for (int c = 0; c < S; ++c) base[(size_t)(j * S + c) * B] = v[j].c[c];
}

template <typename T, int NQ, int S>
__device__ __forceinline__ void ws_load(const T* ws, long long B, int slot, long long p,
                                        Jet<T, S> (&v)[NQ]) {
  const T* base = ws + (size_t)slot * NQ * S * B + p;
#pragma unroll
  for (int j = 0; j < NQ; ++j)
#pragma unroll
    for (int c = 0; c < S; ++c) v[j].c[c] = base[(size_t)(j * S + c) * B];
}

// L2 prefetch of the rows a LATER iteration of the persistent loop will load (this thread's element
// of `rows` consecutive component rows of one slot): the loads at the top of an iteration are used
// immediately, so without it every iteration exposes one full HBM latency
#ifndef QCP_PREFETCH
#define QCP_PREFETCH 1
#endif
template <typename T>
__device__ __forceinline__ void ws_prefetch(const T* row0, long long B, int rows) {
#if QCP_PREFETCH
  // float64 only: -0.4 % at 4 194 304 points, -3.4 % at 524 288; float32 measured +1 % (its
  // iterations are short enough for the hardware to cover the latency across warps)
  if constexpr (sizeof(T) == 8) {
#pragma unroll 4
    for (int r = 0; r < rows; ++r)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(row0 + (size_t)r * B));
  }
#endif
}

// slot 2 of the workspace: saved tanh values act[2H][B] behind the two jet slots
template <typename T, int NQ, int S>
__device__ __forceinline__ T* ws_act(T* ws, long long B) {
  return ws + (size_t)2 * NQ * S * B;
}
template <typename T, int NQ, int S>
__device__ __forceinline__ const T* ws_act(const T* ws, long long B) {
  return ws + (size_t)2 * NQ * S * B;
}

// slot 3: saved sin / cos of the pre-MLP outputs, trig[2 NQ][B], behind the 2 H tanh rows
template <typename T, int NQ, int S>
__device__ __forceinline__ T* ws_trig(T* ws, long long B, int H) {
  return ws + (size_t)(2 * NQ * S + 2 * H) * B;
}

// cotangent seed of the outputs: (grad_u, grad_r) -> jet cotangent of u
template <typename T, int S, typename TIO = T>
__device__ __forceinline__ Jet<T, S> seed_cotangent(const SolverArgs& a, long long p, bool valid) {
  Jet<T, S> ub;
  jzero(ub);
  if (valid) {
    const TIO* gug = static_cast<const TIO*>(a.gu);
    const TIO* grg = static_cast<const TIO*>(a.gr);
    if (gug) ub.c[0] = (T)gug[p];
    if constexpr (S == 6) {
      const TIO* gsg = static_cast<const TIO*>(a.gs);
      if (gsg) {
        // differentiable-streams mode (nn/pde.py operators that are nonlinear in the streams):
        // the caller's autograd supplies d loss / d (u, u_t, u_x, u_y, u_xx, u_yy) per point
#pragma unroll
        for (int c = 0; c < 6; ++c) ub.c[c] += (T)gsg[6 * p + c];
      }
      if (grg) {
        const T g = (T)grg[p];
        ub.c[1] = T(a.pde.ct) * g; ub.c[2] = T(a.pde.cx) * g; ub.c[3] = T(a.pde.cy) * g;
        ub.c[4] = T(a.pde.cxx) * g; ub.c[5] = T(a.pde.cyy) * g;
      }
    }
  }
  return ub;
}

template <typename T>
__device__ __forceinline__ Stager<T> make_stager(T* acc_all, T* tile_all, int nacc) {
  Stager<T> st;
  const int warp = threadIdx.x >> 5;
  st.lane = threadIdx.x & 31;
  st.acc = acc_all + (size_t)warp * nacc;
  st.tile = tile_all + (size_t)warp * kStageRows * kStagePitch;
  return st;
}

template <typename T>
__device__ __forceinline__ void write_partials(const T* acc_all, int nacc, T* partials) {
  __syncthreads();
  T* out = partials + (size_t)blockIdx.x * nacc;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) {
    T s = T(0);
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) s += acc_all[(size_t)w * nacc + i];
    out[i] = s;
  }
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
// TIO = element type of the caller-facing arrays (X, u, r, streams, grad_u, grad_r, grad_X): the
// plan dtype T, or float when a float64 plan serves the float32 module interface directly (saves
// the cast kernels around every call; the arithmetic and the saved jets stay in T).
template <typename T, int NQ, int ENC, int S, typename TIO>
__global__ void __launch_bounds__(kThreads, Tune<T, S>::kFwd)
solver_forward_kernel(const SolverArgs a) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  const int H = a.H;
  const SmemWeights<T> sw = carve_weights<T>(smem_raw, H, c_elems(NQ, ENC));
  load_weights<T, NQ, ENC>(sw, a);
  __syncthreads();

  const TIO* Xg = static_cast<const TIO*>(a.X);
  TIO* ug = static_cast<TIO*>(a.u);
  TIO* rg = static_cast<TIO*>(a.r);
  TIO* sg = static_cast<TIO*>(a.streams);
  T* wsg = static_cast<T*>(a.ws);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < a.B; p += stride) {
    T X[3] = {(T)Xg[3 * p], (T)Xg[3 * p + 1], (T)Xg[3 * p + 2]};
    Jet<T, S> z[NQ], q[NQ], u;
    T* act = (SaveActS<T, S>::value && wsg) ? ws_act<T, NQ, S>(wsg, a.B) + p : nullptr;
    pre_forward<T, NQ, S>(sw, H, X, z, act, a.B);
    if (wsg) ws_store<T, NQ, S>(wsg, a.B, 0, p, z);
    feature_forward<T, NQ, ENC, S>(sw.C, z, q,
                                   (SaveAct<T>::value && wsg) ? ws_trig<T, NQ, S>(wsg, a.B, H) + p : nullptr,
                                   a.B);
    if (wsg) ws_store<T, NQ, S>(wsg, a.B, 1, p, q);
    post_forward<T, NQ, S>(sw, H, q, u, act ? act + (size_t)H * a.B : nullptr, a.B);
    ug[p] = (TIO)u.c[0];
    if constexpr (S == 6) {
      if (rg) {
        rg[p] = (TIO)(T(a.pde.ct) * u.c[1] + T(a.pde.cx) * u.c[2] + T(a.pde.cy) * u.c[3] +
                      T(a.pde.cxx) * u.c[4] + T(a.pde.cyy) * u.c[5]);
      }
      if (sg) {
#pragma unroll
        for (int c = 0; c < 6; ++c) sg[6 * p + c] = (TIO)u.c[c];
      }
    }
  }
}

// fused backward: recompute the forward from X, then the whole reverse sweep
template <typename T, int NQ, int ENC, int S, typename TIO>
__global__ void __launch_bounds__(kThreads, Tune<T, S>::kFused)
solver_backward_kernel(const SolverArgs a) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  constexpr int CE = c_elems(NQ, ENC);
  const int H = a.H;
  const int nacc = nacc_solver(NQ, ENC, H);
  const SmemWeights<T> sw = carve_weights<T>(smem_raw, H, CE);
  T* acc_all = reinterpret_cast<T*>(smem_raw + ((smem_weights_bytes<T>(H, CE) + 31) & ~size_t(31)));
  T* tile_all = acc_all + (size_t)kWarpsPerBlock * nacc;
  load_weights<T, NQ, ENC>(sw, a);
  for (int i = threadIdx.x; i < kWarpsPerBlock * nacc; i += blockDim.x) acc_all[i] = T(0);
  __syncthreads();
  Stager<T> st = make_stager<T>(acc_all, tile_all, nacc);

  const TIO* Xg = static_cast<const TIO*>(a.X);
  TIO* gXg = static_cast<TIO*>(a.gX);
  const long long stride = (long long)gridDim.x * blockDim.x;
  // whole warps iterate together (staging needs every lane); out-of-range lanes carry zero seeds
  const long long Bpad = (a.B + 31) & ~31LL;
  for (long long p0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; p0 < Bpad; p0 += stride) {
    const bool valid = p0 < a.B;
    const long long p = valid ? p0 : a.B - 1;
    T X[3] = {(T)Xg[3 * p], (T)Xg[3 * p + 1], (T)Xg[3 * p + 2]};
    const Jet<T, S> ub = seed_cotangent<T, S, TIO>(a, p, valid);
    st.begin();
    st.put(ub.c[0]);                                       // d b4
    Jet<T, S> z[NQ], q[NQ], qb[NQ], zb[NQ];
    pre_forward<T, NQ, S>(sw, H, X, z);
    feature_forward<T, NQ, ENC, S>(sw.C, z, q);
    post_backward<T, NQ, S>(sw, H, q, ub, qb, st);
    st.flush();                  // segment boundary: d C starts at a known accumulator index
    feature_backward<T, NQ, ENC, S>(sw.C, z, qb, zb, st);
    st.flush();
    st.base = nacc_post(NQ, H) + nacc_contract(NQ, ENC);
    T Xb[3];
    pre_backward<T, NQ, S>(sw, H, X, zb, Xb, st);
    st.flush();
    if (gXg && valid) {
      gXg[3 * p] = (TIO)Xb[0]; gXg[3 * p + 1] = (TIO)Xb[1]; gXg[3 * p + 2] = (TIO)Xb[2];
    }
  }
  write_partials<T>(acc_all, nacc, static_cast<T*>(a.partials));
}

// ---- split backward (the forward saved its jets) --------------------------------------------
//   post_backward_kernel     q (slot 1), grad_u/grad_r  -> qb (slot 1);  d b4, d w3, d b3, d w4
//   contract_backward_kernel z (slot 0), qb (slot 1)    -> zb (slot 0);  d C
//   pre_backward_kernel      X, zb (slot 0)             -> grad_X;       d b2, d w1, d b1, d w2
// Their accumulators are consecutive segments of the fused kernel's accumulator order.
template <typename T, int NQ, int ENC, int S, typename TIO>
__global__ void __launch_bounds__(kThreads, Tune<T, S>::kPost)
post_backward_kernel(const SolverArgs a) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  constexpr int CE = c_elems(NQ, ENC);
  const int H = a.H;
  const int nacc = nacc_post(NQ, H);
  const SmemWeights<T> sw = carve_weights<T>(smem_raw, H, CE);
  T* acc_all = reinterpret_cast<T*>(smem_raw + ((smem_weights_bytes<T>(H, CE) + 31) & ~size_t(31)));
  T* tile_all = acc_all + (size_t)kWarpsPerBlock * nacc;
  load_weights<T, NQ, ENC>(sw, a);
  for (int i = threadIdx.x; i < kWarpsPerBlock * nacc; i += blockDim.x) acc_all[i] = T(0);
  __syncthreads();
  Stager<T> st = make_stager<T>(acc_all, tile_all, nacc);

  T* wsg = static_cast<T*>(a.ws);
  if constexpr (S == 1 && QCP_VALUE_PPT > 1) {
    constexpr int K = QCP_VALUE_PPT;
    const long long chunk = (long long)blockDim.x * K;
    const T* act = ws_act<T, NQ, S>(wsg, a.B) + (size_t)H * a.B;
    for (long long b0 = (long long)blockIdx.x * chunk; b0 < a.B; b0 += (long long)gridDim.x * chunk) {
      long long pts[K];
      bool ok[K];
      T q[K][NQ], qb[K][NQ], ub[K];
      T db4 = T(0);
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const long long p0 = b0 + (long long)j * blockDim.x + threadIdx.x;
        ok[j] = p0 < a.B;
        pts[j] = ok[j] ? p0 : a.B - 1;
        ub[j] = seed_cotangent<T, 1, TIO>(a, pts[j], ok[j]).c[0];
        db4 += ub[j];
#pragma unroll
        for (int i = 0; i < NQ; ++i) q[j][i] = wsg[(size_t)(NQ + i) * a.B + pts[j]];   // slot 1
      }
      st.begin();
      st.put(db4);                                         // d b4
      post_backward_value<T, NQ, K, SaveActS<T, 1>::value>(sw, H, q, ub, qb, st, act, a.B, pts);
      st.flush();
#pragma unroll
      for (int j = 0; j < K; ++j)
        if (ok[j]) {
#pragma unroll
          for (int i = 0; i < NQ; ++i) wsg[(size_t)(NQ + i) * a.B + pts[j]] = qb[j][i];
        }
    }
    write_partials<T>(acc_all, nacc, static_cast<T*>(a.partials));
    return;
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long Bpad = (a.B + 31) & ~31LL;
  for (long long p0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; p0 < Bpad; p0 += stride) {
    const bool valid = p0 < a.B;
    const long long p = valid ? p0 : a.B - 1;
    const Jet<T, S> ub = seed_cotangent<T, S, TIO>(a, p, valid);
    Jet<T, S> q[NQ], qb[NQ];
    ws_load<T, NQ, S>(wsg, a.B, 1, p, q);
    if (p0 + stride < a.B) ws_prefetch<T>(wsg + (size_t)NQ * S * a.B + p0 + stride, a.B, NQ * S);
    st.begin();
    st.put(ub.c[0]);                                       // d b4
    post_backward<T, NQ, S, SaveActS<T, S>::value>(sw, H, q, ub, qb, st,
                                  ws_act<T, NQ, S>(wsg, a.B) + (size_t)H * a.B + p, a.B);
    st.flush();
    if (valid) ws_store<T, NQ, S>(wsg, a.B, 1, p, qb);
  }
  write_partials<T>(acc_all, nacc, static_cast<T*>(a.partials));
}

// MODE >= 0 overrides the A-half loop rolling of Tune<> (the small-batch variant, see
// kRollSmallIters below)
template <typename T, int NQ, int ENC, int S, int MODE = -1>
__global__ void __launch_bounds__(kThreads, Tune<T, S>::kContract)
contract_backward_kernel(const SolverArgs a) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  constexpr int CE = c_elems(NQ, ENC);
  constexpr int nacc = nacc_contract(NQ, ENC);
  T* sC = reinterpret_cast<T*>(smem_raw);
  T* acc_all = sC + ((CE + 3) & ~3);
  T* tile_all = acc_all + (size_t)kWarpsPerBlock * nacc;
  load_C<T, NQ, ENC>(sC, static_cast<const T*>(a.C));
  for (int i = threadIdx.x; i < kWarpsPerBlock * nacc; i += blockDim.x) acc_all[i] = T(0);
  __syncthreads();
  Stager<T> st = make_stager<T>(acc_all, tile_all, nacc);

  T* wsg = static_cast<T*>(a.ws);
  if constexpr (S == 1 && ENC == QCP_ENC_ANGLE && QCP_VALUE_PPT_CONTRACT > 1) {
    constexpr int K = QCP_VALUE_PPT_CONTRACT;
    const long long chunk = (long long)blockDim.x * K;
    for (long long b0 = (long long)blockIdx.x * chunk; b0 < a.B; b0 += (long long)gridDim.x * chunk) {
      long long pts[K];
      bool ok[K];
      Jet<T, 1> z[K][NQ], qb[K][NQ], zb[K][NQ];
      AngleFeat<T, NQ, 1> f[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const long long p0 = b0 + (long long)k * blockDim.x + threadIdx.x;
        ok[k] = p0 < a.B;
        pts[k] = ok[k] ? p0 : a.B - 1;
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
          z[k][j].c[0] = wsg[(size_t)j * a.B + pts[k]];                               // slot 0
          qb[k][j].c[0] = ok[k] ? wsg[(size_t)(NQ + j) * a.B + pts[k]] : T(0);        // slot 1
          jzero(zb[k][j]);
        }
        if constexpr (SaveAct<T>::value) {
          const T* trig = ws_trig<T, NQ, 1>(wsg, a.B, a.H) + pts[k];
          T sn[NQ], cs[NQ];
#pragma unroll
          for (int j = 0; j < NQ; ++j) {
            sn[j] = trig[(size_t)j * a.B];
            cs[j] = trig[(size_t)(NQ + j) * a.B];
          }
          angle_features<T, NQ, 1>(z[k], sn, cs, f[k]);
        } else {
          angle_forward<T, NQ, 1>(z[k], f[k]);
        }
      }
      st.begin();
      angle_backward_value<T, NQ, K>(sC, z, f, qb, zb, st);
      st.flush();
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (ok[k]) {
#pragma unroll
          for (int j = 0; j < NQ; ++j) wsg[(size_t)j * a.B + pts[k]] = zb[k][j].c[0];
        }
    }
    write_partials<T>(acc_all, nacc, static_cast<T*>(a.partials));
    return;
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long Bpad = (a.B + 31) & ~31LL;
  // the loop bound is uniform over the block (warps past the end run with valid == false), so the
  // body may contain block-wide barriers
  for (long long b0 = (long long)blockIdx.x * blockDim.x; b0 < Bpad; b0 += stride) {
    const long long p0 = b0 + threadIdx.x;
    const bool valid = p0 < a.B;
    const long long p = valid ? p0 : a.B - 1;
    if (p0 + stride < a.B) ws_prefetch<T>(wsg + p0 + stride, a.B, 2 * NQ * S);   // z and qb rows
    st.begin();
    if constexpr (ENC == QCP_ENC_ANGLE && S == 6 && Tune<T, S>::kTwoPass && QCP_STREAM_PASSES) {
      angle_backward_ws<T, NQ, MODE>(sC, wsg, a.B, p, valid, st,
                                     SaveAct<T>::value ? ws_trig<T, NQ, S>(wsg, a.B, a.H) + p : nullptr);
      st.flush();
    } else {
      Jet<T, S> z[NQ], qb[NQ], zb[NQ];
      ws_load<T, NQ, S>(wsg, a.B, 0, p, z);
      ws_load<T, NQ, S>(wsg, a.B, 1, p, qb);
      if (!valid) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) jzero(qb[i]);
      }
      feature_backward<T, NQ, ENC, S>(sC, z, qb, zb, st);
      st.flush();
      if (valid) ws_store<T, NQ, S>(wsg, a.B, 0, p, zb);
    }
  }
  write_partials<T>(acc_all, nacc, static_cast<T*>(a.partials));
}

template <typename T, int NQ, int ENC, int S, typename TIO>
__global__ void __launch_bounds__(kThreads, Tune<T, S>::kPre)
pre_backward_kernel(const SolverArgs a) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  constexpr int CE = c_elems(NQ, ENC);
  const int H = a.H;
  const int nacc = nacc_pre(NQ, H);
  const SmemWeights<T> sw = carve_weights<T>(smem_raw, H, CE);
  T* acc_all = reinterpret_cast<T*>(smem_raw + ((smem_weights_bytes<T>(H, CE) + 31) & ~size_t(31)));
  T* tile_all = acc_all + (size_t)kWarpsPerBlock * nacc;
  load_weights<T, NQ, ENC>(sw, a);
  for (int i = threadIdx.x; i < kWarpsPerBlock * nacc; i += blockDim.x) acc_all[i] = T(0);
  __syncthreads();
  Stager<T> st = make_stager<T>(acc_all, tile_all, nacc);

  const TIO* Xg = static_cast<const TIO*>(a.X);
  TIO* gXg = static_cast<TIO*>(a.gX);
  const T* wsg = static_cast<const T*>(a.ws);
  if constexpr (S == 1 && QCP_VALUE_PPT > 1) {
    constexpr int K = QCP_VALUE_PPT;
    const long long chunk = (long long)blockDim.x * K;
    const T* act = ws_act<T, NQ, S>(wsg, a.B);
    for (long long b0 = (long long)blockIdx.x * chunk; b0 < a.B; b0 += (long long)gridDim.x * chunk) {
      long long pts[K];
      bool ok[K];
      T X[K][3], zb[K][NQ], Xb[K][3];
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const long long p0 = b0 + (long long)j * blockDim.x + threadIdx.x;
        ok[j] = p0 < a.B;
        pts[j] = ok[j] ? p0 : a.B - 1;
#pragma unroll
        for (int d = 0; d < 3; ++d) X[j][d] = (T)Xg[3 * pts[j] + d];
#pragma unroll
        for (int i = 0; i < NQ; ++i) zb[j][i] = ok[j] ? wsg[(size_t)i * a.B + pts[j]] : T(0);   // slot 0
      }
      st.begin();
      pre_backward_value<T, NQ, K, SaveActS<T, 1>::value>(sw, H, X, zb, Xb, st, act, a.B, pts);
      st.flush();
      if (gXg) {
#pragma unroll
        for (int j = 0; j < K; ++j)
          if (ok[j]) {
#pragma unroll
            for (int d = 0; d < 3; ++d) gXg[3 * pts[j] + d] = (TIO)Xb[j][d];
          }
      }
    }
    write_partials<T>(acc_all, nacc, static_cast<T*>(a.partials));
    return;
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long Bpad = (a.B + 31) & ~31LL;
  for (long long p0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; p0 < Bpad; p0 += stride) {
    const bool valid = p0 < a.B;
    const long long p = valid ? p0 : a.B - 1;
    T X[3] = {(T)Xg[3 * p], (T)Xg[3 * p + 1], (T)Xg[3 * p + 2]};
    Jet<T, S> zb[NQ];
    ws_load<T, NQ, S>(wsg, a.B, 0, p, zb);
    if (p0 + stride < a.B) ws_prefetch<T>(wsg + p0 + stride, a.B, NQ * S);
    if (!valid) {
#pragma unroll
      for (int j = 0; j < NQ; ++j) jzero(zb[j]);
    }
    T Xb[3];
    st.begin();
    pre_backward<T, NQ, S, SaveActS<T, S>::value>(sw, H, X, zb, Xb, st,
                                              ws_act<T, NQ, S>(wsg, a.B) + p, a.B);
    st.flush();
    if (gXg && valid) {
      gXg[3 * p] = (TIO)Xb[0]; gXg[3 * p + 1] = (TIO)Xb[1]; gXg[3 * p + 2] = (TIO)Xb[2];
    }
  }
  write_partials<T>(acc_all, nacc, static_cast<T*>(a.partials));
}

// ---- stand-alone quantum layer (DVQuantumLayer.forward / its reverse mode) ----------------
template <typename T, int NQ, int ENC>
__global__ void __launch_bounds__(kThreads)
layer_forward_kernel(const LayerArgs a) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  T* sC = reinterpret_cast<T*>(smem_raw);
  load_C<T, NQ, ENC>(sC, static_cast<const T*>(a.C));
  __syncthreads();
  const T* zg = static_cast<const T*>(a.z);
  T* qg = static_cast<T*>(a.q);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < a.B; p += stride) {
    Jet<T, 1> z[NQ], q[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) z[j].c[0] = zg[p * NQ + j];
    feature_forward<T, NQ, ENC, 1>(sC, z, q);
#pragma unroll
    for (int i = 0; i < NQ; ++i) qg[(long long)i * a.B + p] = q[i].c[0];
  }
}

template <typename T, int NQ, int ENC>
__global__ void __launch_bounds__(kThreads)
layer_backward_kernel(const LayerArgs a) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  constexpr int CE = c_elems(NQ, ENC);
  constexpr int nacc = nacc_contract(NQ, ENC);
  T* sC = reinterpret_cast<T*>(smem_raw);
  T* acc_all = sC + ((CE + 3) & ~3);
  T* tile_all = acc_all + kWarpsPerBlock * nacc;
  load_C<T, NQ, ENC>(sC, static_cast<const T*>(a.C));
  for (int i = threadIdx.x; i < kWarpsPerBlock * nacc; i += blockDim.x) acc_all[i] = T(0);
  __syncthreads();
  Stager<T> st = make_stager<T>(acc_all, tile_all, nacc);

  const T* zg = static_cast<const T*>(a.z);
  const T* gq = static_cast<const T*>(a.gq);
  T* gz = static_cast<T*>(a.gz);
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long Bpad = (a.B + 31) & ~31LL;
  for (long long p0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; p0 < Bpad; p0 += stride) {
    const bool valid = p0 < a.B;
    const long long p = valid ? p0 : a.B - 1;
    Jet<T, 1> z[NQ], qb[NQ], zb[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      z[j].c[0] = zg[p * NQ + j];
      qb[j].c[0] = valid ? gq[(long long)j * a.B + p] : T(0);
    }
    st.begin();
    feature_backward<T, NQ, ENC, 1>(sC, z, qb, zb, st);
    st.flush();
    if (gz && valid) {
#pragma unroll
      for (int j = 0; j < NQ; ++j) gz[p * NQ + j] = zb[j].c[0];
    }
  }
  write_partials<T>(acc_all, nacc, static_cast<T*>(a.partials));
}

// ------------------------------------------------------------------------------------------
// host-side dispatch (instantiated per dtype in qcp_point_f32.cu / qcp_point_f64.cu)
// ------------------------------------------------------------------------------------------
template <typename T>
size_t tiles_bytes() {
  return sizeof(T) * (size_t)kWarpsPerBlock * kStageRows * kStagePitch;
}

template <typename T>
size_t weights_bytes_padded(int n, int enc, int H) {
  return (smem_weights_bytes<T>(H, c_elems(n, enc)) + 31) & ~size_t(31);
}

template <typename T>
size_t solver_forward_smem(int n, int enc, int H) {
  return smem_weights_bytes<T>(H, c_elems(n, enc));
}

template <typename T>
size_t solver_backward_smem_impl(int n, int enc, int H) {
  return weights_bytes_padded<T>(n, enc, H) +
         sizeof(T) * (size_t)kWarpsPerBlock * nacc_solver(n, enc, H) + tiles_bytes<T>();
}

template <typename T>
size_t c_only_smem(int n, int enc) {
  return sizeof(T) * ((size_t)((c_elems(n, enc) + 3) & ~3) +
                      (size_t)kWarpsPerBlock * nacc_contract(n, enc)) + tiles_bytes<T>();
}

// which: 0 post, 1 contract, 2 pre
template <typename T>
size_t split_smem(int which, int n, int enc, int H) {
  if (which == 0)
    return weights_bytes_padded<T>(n, enc, H) + sizeof(T) * (size_t)kWarpsPerBlock * nacc_post(n, H) +
           tiles_bytes<T>();
  if (which == 1) return c_only_smem<T>(n, enc);
  return weights_bytes_padded<T>(n, enc, H) + sizeof(T) * (size_t)kWarpsPerBlock * nacc_pre(n, H) +
         tiles_bytes<T>();
}

template <typename K, typename A>
int launch_checked(K kernel, int grid, size_t smem, cudaStream_t s, const char* what,
                   const A& args) {
  cudaError_t e = cudaSuccess;
  if (smem > 48 * 1024)
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(smem=%zu) failed: %s", what, smem, cudaGetErrorString(e));
    return 1;
  }
  kernel<<<grid, kThreads, smem, s>>>(args);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

template <typename K>
int blocks_per_sm(K kernel, size_t smem) {
  int per_sm = 0;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem) != cudaSuccess ||
      per_sm < 1)
    per_sm = 1;
  // the plan's partial-sum slots are sized for kMaxBlocksPerSm blocks per SM (qcp_plan.cu)
  if (per_sm > kMaxBlocksPerSm) per_sm = kMaxBlocksPerSm;
  return per_sm;
}

#define QCP_DISPATCH_NQ_ENC(N, E, ...)                                                \
  do {                                                                                \
    if ((N) == 2 && (E) == 0) { constexpr int NQ = 2, ENC = 0; __VA_ARGS__; }          \
    else if ((N) == 3 && (E) == 0) { constexpr int NQ = 3, ENC = 0; __VA_ARGS__; }     \
    else if ((N) == 4 && (E) == 0) { constexpr int NQ = 4, ENC = 0; __VA_ARGS__; }     \
    else if ((N) == 2 && (E) == 1) { constexpr int NQ = 2, ENC = 1; __VA_ARGS__; }     \
    else if ((N) == 3 && (E) == 1) { constexpr int NQ = 3, ENC = 1; __VA_ARGS__; }     \
    else if ((N) == 4 && (E) == 1) { constexpr int NQ = 4, ENC = 1; __VA_ARGS__; }     \
    else { set_error("fused engine supports 2..4 qubits, got n=%d enc=%d", (N), (E)); return 1; } \
  } while (0)

// QCP_IO(KERNEL, ...): pick the caller-facing element type (a.io_f32 => float, else T)
#define QCP_IO_LAUNCH(KERNEL, S_, GRID, SMEM, WHAT, ARGS)                                        \
  ((ARGS).io_f32 ? launch_checked(&KERNEL<T, NQ, ENC, S_, float>, GRID, SMEM, s, WHAT, ARGS)     \
                 : launch_checked(&KERNEL<T, NQ, ENC, S_, T>, GRID, SMEM, s, WHAT, ARGS))

template <typename T>
int launch_solver_forward(int n, int enc, int mode, const SolverArgs& a, int grid, cudaStream_t s) {
  const size_t smem = solver_forward_smem<T>(n, enc, a.H);
  QCP_DISPATCH_NQ_ENC(n, enc, {
    if (mode == QCP_MODE_RESIDUAL)
      return QCP_IO_LAUNCH(solver_forward_kernel, 6, grid, smem, "solver_forward<residual>", a);
    return QCP_IO_LAUNCH(solver_forward_kernel, 1, grid, smem, "solver_forward<value>", a);
  });
  return 1;
}

template <typename T>
int launch_solver_backward(int n, int enc, int mode, const SolverArgs& a, int grid, cudaStream_t s) {
  const size_t smem = solver_backward_smem_impl<T>(n, enc, a.H);
  QCP_DISPATCH_NQ_ENC(n, enc, {
    if (mode == QCP_MODE_RESIDUAL)
      return QCP_IO_LAUNCH(solver_backward_kernel, 6, grid, smem, "solver_backward<residual>", a);
    return QCP_IO_LAUNCH(solver_backward_kernel, 1, grid, smem, "solver_backward<value>", a);
  });
  return 1;
}

template <typename T>
int launch_layer_forward(int n, int enc, const LayerArgs& a, int grid, cudaStream_t s) {
  const size_t smem = sizeof(T) * (size_t)c_elems(n, enc);
  QCP_DISPATCH_NQ_ENC(n, enc, {
    return launch_checked(&layer_forward_kernel<T, NQ, ENC>, grid, smem, s, "layer_forward", a);
  });
  return 1;
}

template <typename T>
int launch_layer_backward(int n, int enc, const LayerArgs& a, int grid, cudaStream_t s) {
  const size_t smem = c_only_smem<T>(n, enc);
  QCP_DISPATCH_NQ_ENC(n, enc, {
    return launch_checked(&layer_backward_kernel<T, NQ, ENC>, grid, smem, s, "layer_backward", a);
  });
  return 1;
}

template <typename T>
int solver_split_grids(int n, int enc, int mode, int H, int num_sms, SplitGrids* g) {
  const size_t m0 = split_smem<T>(0, n, enc, H), m1 = split_smem<T>(1, n, enc, H),
               m2 = split_smem<T>(2, n, enc, H);
  QCP_DISPATCH_NQ_ENC(n, enc, {
    if (mode == QCP_MODE_RESIDUAL) {
      g->post = num_sms * blocks_per_sm(&post_backward_kernel<T, NQ, ENC, 6, T>, m0);
      g->contract = num_sms * blocks_per_sm(&contract_backward_kernel<T, NQ, ENC, 6>, m1);
      g->pre = num_sms * blocks_per_sm(&pre_backward_kernel<T, NQ, ENC, 6, T>, m2);
    } else {
      g->post = num_sms * blocks_per_sm(&post_backward_kernel<T, NQ, ENC, 1, T>, m0);
      g->contract = num_sms * blocks_per_sm(&contract_backward_kernel<T, NQ, ENC, 1>, m1);
      g->pre = num_sms * blocks_per_sm(&pre_backward_kernel<T, NQ, ENC, 1, T>, m2);
    }
    return 0;
  });
  return 1;
}

#ifndef QCP_ROLL_SMALL_ITERS
#define QCP_ROLL_SMALL_ITERS 20
#endif
constexpr long long kRollSmallIters = QCP_ROLL_SMALL_ITERS;
// the small-batch variant exists where the default is the unrolled two-pass float64 kernel
template <typename T, int NQ, int ENC>
struct ContractSmall {
  static constexpr bool value = std::is_same<T, double>::value && NQ == 4 && ENC == QCP_ENC_ANGLE &&
                                Tune<T, 6>::kTwoPass && Tune<T, 6>::kRollMode == 0 &&
                                (QCP_STREAM_PASSES != 0) && kRollSmallIters > 0;
};

template <typename T>
int launch_solver_backward_split(int n, int enc, int mode, const SolverArgs& a, const SplitGrids& g,
                                 void* part_post, void* part_contract, void* part_pre,
                                 cudaStream_t s, cudaEvent_t after_contract, cudaEvent_t after_post) {
  // after_contract (optional) is recorded between the contraction adjoint and the pre-MLP adjoint:
  // d C is complete there, so the caller can reduce it and run theta_grad next to pre_backward.
  // after_post (optional) is recorded behind the post-MLP adjoint: another chain's adjoints can be
  // held back until then (qcp_solver_backward_after_post)
  SolverArgs a0 = a, a1 = a, a2 = a;
  a0.partials = part_post; a1.partials = part_contract; a2.partials = part_pre;
  const size_t m0 = split_smem<T>(0, n, enc, a.H), m1 = split_smem<T>(1, n, enc, a.H),
               m2 = split_smem<T>(2, n, enc, a.H);
  QCP_DISPATCH_NQ_ENC(n, enc, {
    if (mode == QCP_MODE_RESIDUAL) {
      if (QCP_IO_LAUNCH(post_backward_kernel, 6, g.post, m0, "post_backward", a0)) return 1;
      if (after_post && cudaEventRecord(after_post, s) != cudaSuccess) { set_error("cudaEventRecord failed"); return 1; }
      // few iterations per thread: the fully unrolled body (336 KB of SASS in float64) is fetched
      // cold at every launch and the first pass over it costs as much as several warm ones, so a
      // variant with the outer A-half trit loop rolled (3x less code) wins below ~20 iterations
      // (measured: -4.7 % at 524 288 points, -13 % at 65 536, +2 % at 1 048 576, +8 % at 4 194 304)
      const long long iters = (a.B + (long long)g.contract * kThreads - 1) / ((long long)g.contract * kThreads);
      if (ContractSmall<T, NQ, ENC>::value && iters <= kRollSmallIters) {
        if (launch_checked(&contract_backward_kernel<T, NQ, ENC, 6, 2>, g.contract, m1, s, "contract_backward", a1)) return 1;
      } else if (launch_checked(&contract_backward_kernel<T, NQ, ENC, 6>, g.contract, m1, s, "contract_backward", a1)) return 1;
      if (after_contract && cudaEventRecord(after_contract, s) != cudaSuccess) { set_error("cudaEventRecord failed"); return 1; }
      return QCP_IO_LAUNCH(pre_backward_kernel, 6, g.pre, m2, "pre_backward", a2);
    }
    if (QCP_IO_LAUNCH(post_backward_kernel, 1, g.post, m0, "post_backward", a0)) return 1;
    if (after_post && cudaEventRecord(after_post, s) != cudaSuccess) { set_error("cudaEventRecord failed"); return 1; }
    if (launch_checked(&contract_backward_kernel<T, NQ, ENC, 1>, g.contract, m1, s, "contract_backward", a1)) return 1;
    if (after_contract && cudaEventRecord(after_contract, s) != cudaSuccess) { set_error("cudaEventRecord failed"); return 1; }
    return QCP_IO_LAUNCH(pre_backward_kernel, 1, g.pre, m2, "pre_backward", a2);
  });
  return 1;
}

template <typename T>
size_t solver_backward_smem(int n, int enc, int H) {
  return solver_backward_smem_impl<T>(n, enc, H);
}

template <typename T>
int solver_backward_max_grid(int n, int enc, int mode, int H, int num_sms) {
  const size_t smem = solver_backward_smem_impl<T>(n, enc, H);
  QCP_DISPATCH_NQ_ENC(n, enc, {
    if (mode == QCP_MODE_RESIDUAL)
      return num_sms * blocks_per_sm(&solver_backward_kernel<T, NQ, ENC, 6, T>, smem);
    return num_sms * blocks_per_sm(&solver_backward_kernel<T, NQ, ENC, 1, T>, smem);
  });
  return num_sms;
}

}  // namespace qcp
