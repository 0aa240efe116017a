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
// The backward kernel recomputes the forward from X (12 B/point) instead of saving ~1 KB/point of
// intermediates, runs the hand-derived pullbacks of every jet operation, and sums the per-point
// parameter-gradient contributions across the warp through a padded shared-memory staging tile.
#pragma once

#include "qcp_common.cuh"
#include "qcp_jet.cuh"

namespace qcp {

// ------------------------------------------------------------------------------------------
// shared-memory image of the weights
// ------------------------------------------------------------------------------------------
template <typename T>
struct SmemWeights {
  Vec4<T>* w1b;   // [H]  (w1[k,0], w1[k,1], w1[k,2], b1[k])
  Vec4<T>* w2t;   // [H]  w2[j,k], j < n (zero padded)
  Vec4<T>* w3;    // [H]  w3[k,i], i < n (zero padded)
  T* b3w4;        // [H][2] (b3[k], w4[k])
  T* b2;          // [4]
  T* b4;          // [4]  (only [0] used)
  Vec4<T>* C;     // [F]  C[i,s], i < n (zero padded)
};

template <typename T>
__host__ __device__ inline size_t smem_weights_bytes(int H, int F) {
  return sizeof(T) * (size_t)(4 * H * 3 + 2 * H + 4 + 4 + 4 * F);
}

template <typename T>
__device__ __forceinline__ SmemWeights<T> carve_weights(unsigned char* base, int H, int F) {
  SmemWeights<T> s;
  T* p = reinterpret_cast<T*>(base);
  s.w1b = reinterpret_cast<Vec4<T>*>(p); p += 4 * H;
  s.w2t = reinterpret_cast<Vec4<T>*>(p); p += 4 * H;
  s.w3 = reinterpret_cast<Vec4<T>*>(p);  p += 4 * H;
  s.C = reinterpret_cast<Vec4<T>*>(p);   p += 4 * F;
  s.b3w4 = p; p += 2 * H;
  s.b2 = p;   p += 4;
  s.b4 = p;   p += 4;
  return s;
}

template <typename T, int NQ>
__device__ __forceinline__ void load_weights(const SmemWeights<T>& s, const SolverArgs& a, int F) {
  const int H = a.H;
  const T* w1 = static_cast<const T*>(a.w1);
  const T* b1 = static_cast<const T*>(a.b1);
  const T* w2 = static_cast<const T*>(a.w2);
  const T* b2 = static_cast<const T*>(a.b2);
  const T* w3 = static_cast<const T*>(a.w3);
  const T* b3 = static_cast<const T*>(a.b3);
  const T* w4 = static_cast<const T*>(a.w4);
  const T* b4 = static_cast<const T*>(a.b4);
  const T* C = static_cast<const T*>(a.C);
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
  T* Cs = reinterpret_cast<T*>(s.C);
  for (int i = threadIdx.x; i < 4 * F; i += blockDim.x) Cs[i] = C[i];
  if (threadIdx.x < 4) {
    s.b2[threadIdx.x] = threadIdx.x < NQ ? b2[threadIdx.x] : T(0);
    s.b4[threadIdx.x] = b4[0];
  }
}

// ------------------------------------------------------------------------------------------
// pre MLP:  z_j = b2_j + sum_k w2[j,k] tanh(b1_k + w1[k,:] . X)
// ------------------------------------------------------------------------------------------
template <typename T, int NQ, int S>
__device__ __forceinline__ void pre_forward(const SmemWeights<T>& s, int H, const T (&X)[3],
                                            Jet<T, S> (&z)[NQ]) {
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    jzero(z[j]);
    z[j].c[0] = s.b2[j];
  }
#pragma unroll 2
  for (int k = 0; k < H; ++k) {
    const Vec4<T> w = s.w1b[k];
    const Vec4<T> w2 = s.w2t[k];
    Jet<T, S> a;
    a.c[0] = fma(w.v[0], X[0], fma(w.v[1], X[1], fma(w.v[2], X[2], w.v[3])));
    if constexpr (S == 6) {
      a.c[1] = w.v[0]; a.c[2] = w.v[1]; a.c[3] = w.v[2]; a.c[4] = T(0); a.c[5] = T(0);
    }
    const T h0 = Math<T>::tanh_(a.c[0]);
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
                                             const Jet<T, S> (&q)[NQ], Jet<T, S>& u) {
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
    const T g0 = Math<T>::tanh_(p.c[0]);
    const T f1 = fma(-g0, g0, T(1));
    const T f2 = T(-2) * g0 * f1;
    const Jet<T, S> g = jfunc(p, g0, f1, f2);
    jaxpy(u, w4, g);
  }
}

// ------------------------------------------------------------------------------------------
// angle-encoding features: per-qubit Bloch jets and their tensor products
// ------------------------------------------------------------------------------------------
template <typename T, int NQ, int S>
struct AngleFeat {
  static constexpr int NA = (NQ + 1) / 2;
  static constexpr int NB = NQ - NA;
  static constexpr int FA = ipow3(NA);
  static constexpr int FB = ipow3(NB);
  T sn[NQ], cs[NQ];
  Jet<T, S> y[NQ], w[NQ];   // y = -sin z (Pauli-Y Bloch component), w = cos z (Pauli-Z)
  Jet<T, S> P[FA];          // products over qubits [0, NA); P[0] unused (== 1)
  Jet<T, S> Q[FB];          // products over qubits [NA, NQ); Q[0] unused (== 1)
};

// e[3*s0 + s1] over two qubits (or e[s0] over one); trit 0 = I, 1 = Y, 2 = Z
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

template <typename T, int NQ, int S>
__device__ __forceinline__ void angle_forward(const Jet<T, S> (&z)[NQ], AngleFeat<T, NQ, S>& f) {
  using F = AngleFeat<T, NQ, S>;
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    Math<T>::sincos_(z[j].c[0], &f.sn[j], &f.cs[j]);
    f.y[j] = jfunc(z[j], -f.sn[j], -f.cs[j], f.sn[j]);
    f.w[j] = jfunc(z[j], f.cs[j], -f.sn[j], -f.cs[j]);
  }
  prod_build<T, S, F::NA>(f.P, f.y, f.w);
  prod_build<T, S, F::NB>(f.Q, f.y + F::NA, f.w + F::NA);
}

// t[i] = sum_b C[i, a*FB + b] Q[b]   (Q[0] == 1)
template <typename T, int NQ, int S, int FB>
__device__ __forceinline__ void contract_row(const Vec4<T>* sC, int a, const Jet<T, S>* Q,
                                             Jet<T, S> (&t)[NQ]) {
  const Vec4<T> c0 = sC[a * FB];
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    jzero(t[i]);
    t[i].c[0] = c0.v[i];
  }
#pragma unroll
  for (int b = 1; b < FB; ++b) {
    const Vec4<T> c = sC[a * FB + b];
#pragma unroll
    for (int i = 0; i < NQ; ++i) jaxpy(t[i], c.v[i], Q[b]);
  }
}

template <typename T, int NQ, int S>
__device__ __forceinline__ void angle_contract(const Vec4<T>* sC, const AngleFeat<T, NQ, S>& f,
                                               Jet<T, S> (&q)[NQ]) {
  using F = AngleFeat<T, NQ, S>;
#pragma unroll
  for (int a = 0; a < F::FA; ++a) {
    Jet<T, S> t[NQ];
    contract_row<T, NQ, S, F::FB>(sC, a, f.Q, t);
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      if (a == 0) q[i] = t[i];
      else jmul_acc(q[i], f.P[a], t[i]);
    }
  }
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
    for (int b = a; b < NQ; ++b)
      f.phi[a * NQ - a * (a - 1) / 2 + (b - a)] = jmul(jmul(z[a], z[b]), f.inv);
}

template <typename T, int NQ, int S>
__device__ __forceinline__ void amp_contract(const Vec4<T>* sC, const AmpFeat<T, NQ, S>& f,
                                             Jet<T, S> (&q)[NQ]) {
#pragma unroll
  for (int i = 0; i < NQ; ++i) jzero(q[i]);
#pragma unroll
  for (int s = 0; s < AmpFeat<T, NQ, S>::F; ++s) {
    const Vec4<T> c = sC[s];
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
    for (int r = lane; r < cnt; r += 32) {
      const T* p = tile + r * kStagePitch;
      T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        s0 += p[j]; s1 += p[j + 1]; s2 += p[j + 2]; s3 += p[j + 3];
      }
      acc[base + r] += (s0 + s1) + (s2 + s3);
    }
    __syncwarp();
    base += cnt;
    cnt = 0;
  }
  // make room for `rows` more rows
  __device__ __forceinline__ void reserve(int rows) {
    if (cnt + rows > kStageRows) flush();
  }
};

// ------------------------------------------------------------------------------------------
// reverse sweeps
// ------------------------------------------------------------------------------------------
template <typename T, int NQ, int S>
__device__ __forceinline__ void post_backward(const SmemWeights<T>& s, int H,
                                              const Jet<T, S> (&q)[NQ], const Jet<T, S>& ub,
                                              Jet<T, S> (&qb)[NQ], Stager<T>& st) {
#pragma unroll
  for (int i = 0; i < NQ; ++i) jzero(qb[i]);
  for (int k = 0; k < H; ++k) {
    const Vec4<T> w3 = s.w3[k];
    const T b3 = s.b3w4[2 * k], w4 = s.b3w4[2 * k + 1];
    Jet<T, S> p;
    jzero(p);
    p.c[0] = b3;
#pragma unroll
    for (int i = 0; i < NQ; ++i) jaxpy(p, w3.v[i], q[i]);
    T g0, f1, f2, f3;
    tanh_derivs(p.c[0], g0, f1, f2, f3);
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
  }
}

template <typename T, int NQ, int S>
__device__ __forceinline__ void angle_backward(const Vec4<T>* sC, const Jet<T, S> (&z)[NQ],
                                               const AngleFeat<T, NQ, S>& f,
                                               const Jet<T, S> (&qb)[NQ], Jet<T, S> (&zb)[NQ],
                                               Stager<T>& st) {
  using F = AngleFeat<T, NQ, S>;
  Jet<T, S> Pb[F::FA], Qb[F::FB];
#pragma unroll
  for (int a = 0; a < F::FA; ++a) jzero(Pb[a]);
#pragma unroll
  for (int b = 0; b < F::FB; ++b) jzero(Qb[b]);

#pragma unroll
  for (int a = 0; a < F::FA; ++a) {
    // D_i = pullback of qb_i through multiplication by P[a]; also P[a]'s own cotangent
    Jet<T, S> D[NQ];
    if (a == 0) {
#pragma unroll
      for (int i = 0; i < NQ; ++i) D[i] = qb[i];
    } else {
      Jet<T, S> t[NQ];
      contract_row<T, NQ, S, F::FB>(sC, a, f.Q, t);
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        jzero(D[i]);
        jmul_pull_acc(D[i], qb[i], f.P[a]);
        jmul_pull_acc(Pb[a], qb[i], t[i]);
      }
    }
    st.reserve(F::FB * NQ);
    {
#pragma unroll
      for (int i = 0; i < NQ; ++i) st.put(D[i].c[0]);      // d C[i, a, 0]  (Q[0] == 1)
    }
#pragma unroll
    for (int b = 1; b < F::FB; ++b) {
      const Vec4<T> c = sC[a * F::FB + b];
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        st.put(jdot(D[i], f.Q[b]));                        // d C[i, a, b]
        jaxpy(Qb[b], c.v[i], D[i]);
      }
    }
  }

  Jet<T, S> yb[NQ], wb[NQ];
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    jzero(yb[j]);
    jzero(wb[j]);
  }
  prod_pull<T, S, F::NA>(Pb, f.y, f.w, yb, wb);
  prod_pull<T, S, F::NB>(Qb, f.y + F::NA, f.w + F::NA, yb + F::NA, wb + F::NA);
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    jzero(zb[j]);
    // y = -sin z : f' = -cos, f'' = sin, f''' = cos ;  w = cos z : f' = -sin, f'' = -cos, f''' = sin
    jfunc_pull_acc(zb[j], yb[j], z[j], -f.cs[j], f.sn[j], f.cs[j]);
    jfunc_pull_acc(zb[j], wb[j], z[j], -f.sn[j], -f.cs[j], f.sn[j]);
  }
}

template <typename T, int NQ, int S>
__device__ __forceinline__ void amp_backward(const Vec4<T>* sC, const Jet<T, S> (&z)[NQ],
                                             const AmpFeat<T, NQ, S>& f,
                                             const Jet<T, S> (&qb)[NQ], Jet<T, S> (&zb)[NQ],
                                             Stager<T>& st) {
  Jet<T, S> invb;
  jzero(invb);
#pragma unroll
  for (int j = 0; j < NQ; ++j) jzero(zb[j]);
#pragma unroll
  for (int a = 0; a < NQ; ++a) {
#pragma unroll
    for (int b = a; b < NQ; ++b) {
      const int s = a * NQ - a * (a - 1) / 2 + (b - a);
      const Vec4<T> c = sC[s];
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

template <typename T, int NQ, int S>
__device__ __forceinline__ void pre_backward(const SmemWeights<T>& s, int H, const T (&X)[3],
                                             const Jet<T, S> (&zb)[NQ], T (&Xb)[3],
                                             Stager<T>& st) {
  st.reserve(NQ);
#pragma unroll
  for (int j = 0; j < NQ; ++j) st.put(zb[j].c[0]);          // d b2[j]
  Xb[0] = Xb[1] = Xb[2] = T(0);
  for (int k = 0; k < H; ++k) {
    const Vec4<T> w = s.w1b[k];
    const Vec4<T> w2 = s.w2t[k];
    Jet<T, S> a;
    a.c[0] = fma(w.v[0], X[0], fma(w.v[1], X[1], fma(w.v[2], X[2], w.v[3])));
    if constexpr (S == 6) {
      a.c[1] = w.v[0]; a.c[2] = w.v[1]; a.c[3] = w.v[2]; a.c[4] = T(0); a.c[5] = T(0);
    }
    T h0, f1, f2, f3;
    tanh_derivs(a.c[0], h0, f1, f2, f3);
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
  }
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
template <typename T, int NQ, int ENC, int S>
__global__ void __launch_bounds__(kThreads)
solver_forward_kernel(const SolverArgs a) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  constexpr int F = num_features(NQ, ENC);
  const int H = a.H;
  const SmemWeights<T> sw = carve_weights<T>(smem_raw, H, F);
  load_weights<T, NQ>(sw, a, F);
  __syncthreads();

  const T* Xg = static_cast<const T*>(a.X);
  T* ug = static_cast<T*>(a.u);
  T* rg = static_cast<T*>(a.r);
  T* sg = static_cast<T*>(a.streams);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < a.B; p += stride) {
    T X[3] = {Xg[3 * p], Xg[3 * p + 1], Xg[3 * p + 2]};
    Jet<T, S> z[NQ], q[NQ], u;
    pre_forward<T, NQ, S>(sw, H, X, z);
    if constexpr (ENC == QCP_ENC_ANGLE) {
      AngleFeat<T, NQ, S> f;
      angle_forward<T, NQ, S>(z, f);
      angle_contract<T, NQ, S>(sw.C, f, q);
    } else {
      AmpFeat<T, NQ, S> f;
      amp_forward<T, NQ, S>(z, f);
      amp_contract<T, NQ, S>(sw.C, f, q);
    }
    post_forward<T, NQ, S>(sw, H, q, u);
    ug[p] = u.c[0];
    if constexpr (S == 6) {
      if (rg) {
        rg[p] = T(a.pde.ct) * u.c[1] + T(a.pde.cx) * u.c[2] + T(a.pde.cy) * u.c[3] +
                T(a.pde.cxx) * u.c[4] + T(a.pde.cyy) * u.c[5];
      }
      if (sg) {
#pragma unroll
        for (int c = 0; c < 6; ++c) sg[6 * p + c] = u.c[c];
      }
    }
  }
}

template <typename T, int NQ, int ENC, int S>
__global__ void __launch_bounds__(kThreads)
solver_backward_kernel(const SolverArgs a) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  constexpr int F = num_features(NQ, ENC);
  const int H = a.H;
  const int nacc = nacc_solver(NQ, ENC, H);
  const SmemWeights<T> sw = carve_weights<T>(smem_raw, H, F);
  T* acc_all = reinterpret_cast<T*>(smem_raw + ((smem_weights_bytes<T>(H, F) + 31) & ~size_t(31)));
  T* tile_all = acc_all + (size_t)kWarpsPerBlock * nacc;
  load_weights<T, NQ>(sw, a, F);
  for (int i = threadIdx.x; i < kWarpsPerBlock * nacc; i += blockDim.x) acc_all[i] = T(0);
  __syncthreads();

  const int warp = threadIdx.x >> 5;
  Stager<T> st;
  st.lane = threadIdx.x & 31;
  st.acc = acc_all + (size_t)warp * nacc;
  st.tile = tile_all + (size_t)warp * kStageRows * kStagePitch;

  const T* Xg = static_cast<const T*>(a.X);
  const T* gug = static_cast<const T*>(a.gu);
  const T* grg = static_cast<const T*>(a.gr);
  T* gXg = static_cast<T*>(a.gX);
  const long long stride = (long long)gridDim.x * blockDim.x;
  // whole warps iterate together (staging needs every lane); out-of-range lanes carry zero seeds
  const long long Bpad = (a.B + 31) & ~31LL;
  for (long long p0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; p0 < Bpad; p0 += stride) {
    const bool valid = p0 < a.B;
    const long long p = valid ? p0 : a.B - 1;
    T X[3] = {Xg[3 * p], Xg[3 * p + 1], Xg[3 * p + 2]};
    Jet<T, S> ub;
    jzero(ub);
    if (valid) {
      if (gug) ub.c[0] = gug[p];
      if constexpr (S == 6) {
        if (grg) {
          const T g = grg[p];
          ub.c[1] = T(a.pde.ct) * g; ub.c[2] = T(a.pde.cx) * g; ub.c[3] = T(a.pde.cy) * g;
          ub.c[4] = T(a.pde.cxx) * g; ub.c[5] = T(a.pde.cyy) * g;
        }
      }
    }
    st.begin();
    st.put(ub.c[0]);                                       // d b4

    // ---- recompute forward ----
    Jet<T, S> z[NQ], q[NQ], qb[NQ], zb[NQ];
    pre_forward<T, NQ, S>(sw, H, X, z);
    T Xb[3];
    if constexpr (ENC == QCP_ENC_ANGLE) {
      AngleFeat<T, NQ, S> f;
      angle_forward<T, NQ, S>(z, f);
      angle_contract<T, NQ, S>(sw.C, f, q);
      post_backward<T, NQ, S>(sw, H, q, ub, qb, st);
      angle_backward<T, NQ, S>(sw.C, z, f, qb, zb, st);
    } else {
      AmpFeat<T, NQ, S> f;
      amp_forward<T, NQ, S>(z, f);
      amp_contract<T, NQ, S>(sw.C, f, q);
      post_backward<T, NQ, S>(sw, H, q, ub, qb, st);
      amp_backward<T, NQ, S>(sw.C, z, f, qb, zb, st);
    }
    pre_backward<T, NQ, S>(sw, H, X, zb, Xb, st);
    st.flush();
    if (gXg && valid) {
      gXg[3 * p] = Xb[0]; gXg[3 * p + 1] = Xb[1]; gXg[3 * p + 2] = Xb[2];
    }
  }
  __syncthreads();
  T* out = static_cast<T*>(a.partials) + (size_t)blockIdx.x * nacc;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) {
    T s = T(0);
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) s += acc_all[(size_t)w * nacc + i];
    out[i] = s;
  }
}

// ---- stand-alone quantum layer (DVQuantumLayer.forward / its reverse mode) ----------------
template <typename T, int NQ, int ENC>
__global__ void __launch_bounds__(kThreads)
layer_forward_kernel(const LayerArgs a) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  constexpr int F = num_features(NQ, ENC);
  Vec4<T>* sC = reinterpret_cast<Vec4<T>*>(smem_raw);
  {
    T* d = reinterpret_cast<T*>(sC);
    const T* C = static_cast<const T*>(a.C);
    for (int i = threadIdx.x; i < 4 * F; i += blockDim.x) d[i] = C[i];
  }
  __syncthreads();
  const T* zg = static_cast<const T*>(a.z);
  T* qg = static_cast<T*>(a.q);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < a.B; p += stride) {
    Jet<T, 1> z[NQ], q[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) z[j].c[0] = zg[p * NQ + j];
    if constexpr (ENC == QCP_ENC_ANGLE) {
      AngleFeat<T, NQ, 1> f;
      angle_forward<T, NQ, 1>(z, f);
      angle_contract<T, NQ, 1>(sC, f, q);
    } else {
      AmpFeat<T, NQ, 1> f;
      amp_forward<T, NQ, 1>(z, f);
      amp_contract<T, NQ, 1>(sC, f, q);
    }
#pragma unroll
    for (int i = 0; i < NQ; ++i) qg[(long long)i * a.B + p] = q[i].c[0];
  }
}

template <typename T, int NQ, int ENC>
__global__ void __launch_bounds__(kThreads)
layer_backward_kernel(const LayerArgs a) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  constexpr int F = num_features(NQ, ENC);
  constexpr int nacc = F * NQ;
  Vec4<T>* sC = reinterpret_cast<Vec4<T>*>(smem_raw);
  T* acc_all = reinterpret_cast<T*>(smem_raw) + 4 * F;
  T* tile_all = acc_all + kWarpsPerBlock * nacc;
  {
    T* d = reinterpret_cast<T*>(sC);
    const T* C = static_cast<const T*>(a.C);
    for (int i = threadIdx.x; i < 4 * F; i += blockDim.x) d[i] = C[i];
    for (int i = threadIdx.x; i < kWarpsPerBlock * nacc; i += blockDim.x) acc_all[i] = T(0);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  Stager<T> st;
  st.lane = threadIdx.x & 31;
  st.acc = acc_all + warp * nacc;
  st.tile = tile_all + (size_t)warp * kStageRows * kStagePitch;

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
    if constexpr (ENC == QCP_ENC_ANGLE) {
      AngleFeat<T, NQ, 1> f;
      angle_forward<T, NQ, 1>(z, f);
      angle_backward<T, NQ, 1>(sC, z, f, qb, zb, st);
    } else {
      AmpFeat<T, NQ, 1> f;
      amp_forward<T, NQ, 1>(z, f);
      amp_backward<T, NQ, 1>(sC, z, f, qb, zb, st);
    }
    st.flush();
    if (gz && valid) {
#pragma unroll
      for (int j = 0; j < NQ; ++j) gz[p * NQ + j] = zb[j].c[0];
    }
  }
  __syncthreads();
  T* out = static_cast<T*>(a.partials) + (size_t)blockIdx.x * nacc;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) {
    T s = T(0);
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) s += acc_all[w * nacc + i];
    out[i] = s;
  }
}

// ------------------------------------------------------------------------------------------
// host-side dispatch (instantiated per dtype in qcp_point_f32.cu / qcp_point_f64.cu)
// ------------------------------------------------------------------------------------------
template <typename T>
size_t solver_forward_smem(int n, int enc, int H) {
  return smem_weights_bytes<T>(H, num_features(n, enc));
}

template <typename T>
size_t solver_backward_smem_impl(int n, int enc, int H) {
  size_t w = (smem_weights_bytes<T>(H, num_features(n, enc)) + 31) & ~size_t(31);
  return w + sizeof(T) * ((size_t)kWarpsPerBlock * nacc_solver(n, enc, H) +
                          (size_t)kWarpsPerBlock * kStageRows * kStagePitch);
}

template <typename T>
size_t layer_backward_smem(int n, int enc) {
  const int F = num_features(n, enc);
  return sizeof(T) * ((size_t)4 * F + (size_t)kWarpsPerBlock * F * n +
                      (size_t)kWarpsPerBlock * kStageRows * kStagePitch);
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

#define QCP_DISPATCH_NQ_ENC(N, E, ...)                                       \
  do {                                                                        \
    if ((N) == 2 && (E) == 0) { constexpr int NQ = 2, ENC = 0; __VA_ARGS__; }         \
    else if ((N) == 3 && (E) == 0) { constexpr int NQ = 3, ENC = 0; __VA_ARGS__; }    \
    else if ((N) == 4 && (E) == 0) { constexpr int NQ = 4, ENC = 0; __VA_ARGS__; }    \
    else if ((N) == 2 && (E) == 1) { constexpr int NQ = 2, ENC = 1; __VA_ARGS__; }    \
    else if ((N) == 3 && (E) == 1) { constexpr int NQ = 3, ENC = 1; __VA_ARGS__; }    \
    else if ((N) == 4 && (E) == 1) { constexpr int NQ = 4, ENC = 1; __VA_ARGS__; }    \
    else { set_error("fused engine supports 2..4 qubits, got n=%d enc=%d", (N), (E)); return 1; } \
  } while (0)

template <typename T>
int launch_solver_forward(int n, int enc, int mode, const SolverArgs& a, int grid, cudaStream_t s) {
  const size_t smem = solver_forward_smem<T>(n, enc, a.H);
  QCP_DISPATCH_NQ_ENC(n, enc, {
    if (mode == QCP_MODE_RESIDUAL)
      return launch_checked(&solver_forward_kernel<T, NQ, ENC, 6>, grid, smem, s,
                            "solver_forward<residual>", a);
    return launch_checked(&solver_forward_kernel<T, NQ, ENC, 1>, grid, smem, s,
                          "solver_forward<value>", a);
  });
  return 1;
}

template <typename T>
int launch_solver_backward(int n, int enc, int mode, const SolverArgs& a, int grid, cudaStream_t s) {
  const size_t smem = solver_backward_smem_impl<T>(n, enc, a.H);
  QCP_DISPATCH_NQ_ENC(n, enc, {
    if (mode == QCP_MODE_RESIDUAL)
      return launch_checked(&solver_backward_kernel<T, NQ, ENC, 6>, grid, smem, s,
                            "solver_backward<residual>", a);
    return launch_checked(&solver_backward_kernel<T, NQ, ENC, 1>, grid, smem, s,
                          "solver_backward<value>", a);
  });
  return 1;
}

template <typename T>
int launch_layer_forward(int n, int enc, const LayerArgs& a, int grid, cudaStream_t s) {
  const size_t smem = sizeof(T) * 4 * (size_t)num_features(n, enc);
  QCP_DISPATCH_NQ_ENC(n, enc, {
    return launch_checked(&layer_forward_kernel<T, NQ, ENC>, grid, smem, s, "layer_forward",
                          a);
  });
  return 1;
}

template <typename T>
int launch_layer_backward(int n, int enc, const LayerArgs& a, int grid, cudaStream_t s) {
  const size_t smem = layer_backward_smem<T>(n, enc);
  QCP_DISPATCH_NQ_ENC(n, enc, {
    return launch_checked(&layer_backward_kernel<T, NQ, ENC>, grid, smem, s, "layer_backward",
                          a);
  });
  return 1;
}

template <typename T>
size_t solver_backward_smem(int n, int enc, int H) {
  return solver_backward_smem_impl<T>(n, enc, H);
}

template <typename T>
int solver_backward_max_grid(int n, int enc, int mode, int H, int num_sms) {
  int per_sm = 0;
  const size_t smem = solver_backward_smem_impl<T>(n, enc, H);
  cudaError_t e = cudaErrorInvalidValue;
  QCP_DISPATCH_NQ_ENC(n, enc, {
    if (mode == QCP_MODE_RESIDUAL) {
      auto k = &solver_backward_kernel<T, NQ, ENC, 6>;
      if (smem > 48 * 1024)
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kThreads, smem);
    } else {
      auto k = &solver_backward_kernel<T, NQ, ENC, 1>;
      if (smem > 48 * 1024)
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kThreads, smem);
    }
  });
  if (e != cudaSuccess || per_sm < 1) per_sm = 1;
  return per_sm * num_sms;
}

}  // namespace qcp
