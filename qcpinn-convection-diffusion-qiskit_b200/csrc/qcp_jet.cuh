// qcpinn_b200 -- truncated Taylor "jets" carried by one collocation point.
//
// The convection-diffusion residual (reference nn/pde.py:53-72) needs u, u_t, u_x, u_y, u_xx, u_yy.
// Instead of five nested autograd sweeps we push a small jet through every operation.  Every forward
// rule has a hand-derived pullback, so the backward kernels are the exact adjoint of the forward.
//
// Component layout of a jet with S components:
//   c[0]                         value
//   c[1 .. N1]                   first derivatives along directions WITHOUT a second derivative
//   c[1+N1 .. N1+N2]             first derivatives along directions WITH a second derivative
//   c[1+N1+N2 .. N1+2*N2]        the matching pure second derivatives
// S = 6 -> (N1, N2) = (1, 2): (u, u_t | u_x, u_y | u_xx, u_yy)   the full residual jet
// S = 4 -> (1, 1): (u, u_t | u_x | u_xx)   and   S = 3 -> (0, 1): (u | u_y | u_yy)
//          sub-jets that let the S = 6 adjoint run as two register-light passes
// S = 1 -> (0, 0): plain value (IC / BC points, reference trainer/diffusion_train.py:40-41)
#pragma once

#include <cuda_runtime.h>

namespace qcp {

__host__ __device__ constexpr int jet_n2(int S) { return S == 6 ? 2 : (S >= 3 ? 1 : 0); }
__host__ __device__ constexpr int jet_n1(int S) { return S - 1 - 2 * jet_n2(S); }

template <typename T, int S>
struct Jet {
  static constexpr int N2 = jet_n2(S);
  static constexpr int N1 = jet_n1(S);
  static constexpr int P0 = 1 + N1;        // first "paired" first-derivative slot
  static constexpr int E0 = 1 + N1 + N2;   // first second-derivative slot
  T c[S];
};

template <typename T, int S>
__device__ __forceinline__ void jzero(Jet<T, S>& a) {
#pragma unroll
  for (int i = 0; i < S; ++i) a.c[i] = T(0);
}

// z += w * h
template <typename T, int S>
__device__ __forceinline__ void jaxpy(Jet<T, S>& z, T w, const Jet<T, S>& h) {
#pragma unroll
  for (int i = 0; i < S; ++i) z.c[i] = fma(w, h.c[i], z.c[i]);
}

template <typename T, int S>
__device__ __forceinline__ void jadd(Jet<T, S>& z, const Jet<T, S>& h) {
#pragma unroll
  for (int i = 0; i < S; ++i) z.c[i] += h.c[i];
}

// sum over components of cotangent x tangent
template <typename T, int S>
__device__ __forceinline__ T jdot(const Jet<T, S>& a, const Jet<T, S>& b) {
  T s = a.c[0] * b.c[0];
#pragma unroll
  for (int i = 1; i < S; ++i) s = fma(a.c[i], b.c[i], s);
  return s;
}

// product rule:  (ab)_d = a_d b + a b_d ;  (ab)_dd = a_dd b + 2 a_d b_d + a b_dd
template <typename T, int S>
__device__ __forceinline__ Jet<T, S> jmul(const Jet<T, S>& a, const Jet<T, S>& b) {
  using J = Jet<T, S>;
  J r;
  r.c[0] = a.c[0] * b.c[0];
#pragma unroll
  for (int i = 1; i < J::E0; ++i) r.c[i] = fma(a.c[i], b.c[0], a.c[0] * b.c[i]);
#pragma unroll
  for (int k = 0; k < J::N2; ++k) {
    const int e = J::E0 + k, p = J::P0 + k;
    const T t = fma(a.c[e], b.c[0], a.c[0] * b.c[e]);
    r.c[e] = fma(T(2) * a.c[p], b.c[p], t);
  }
  return r;
}

// acc += a * b
template <typename T, int S>
__device__ __forceinline__ void jmul_acc(Jet<T, S>& acc, const Jet<T, S>& a, const Jet<T, S>& b) {
  using J = Jet<T, S>;
  acc.c[0] = fma(a.c[0], b.c[0], acc.c[0]);
#pragma unroll
  for (int i = 1; i < J::E0; ++i) acc.c[i] = fma(a.c[i], b.c[0], fma(a.c[0], b.c[i], acc.c[i]));
#pragma unroll
  for (int k = 0; k < J::N2; ++k) {
    const int e = J::E0 + k, p = J::P0 + k;
    const T t = fma(a.c[e], b.c[0], fma(a.c[0], b.c[e], acc.c[e]));
    acc.c[e] = fma(T(2) * a.c[p], b.c[p], t);
  }
}

// Pullback of c = a*b w.r.t. b:  bbar += (d c / d b)^T cbar, with the other factor a.
// Linear in cbar, so the partial cotangents of a two-pass adjoint can be pulled independently.
template <typename T, int S>
__device__ __forceinline__ void jmul_pull_acc(Jet<T, S>& bbar, const Jet<T, S>& cbar,
                                              const Jet<T, S>& a) {
  using J = Jet<T, S>;
  T s0 = fma(cbar.c[0], a.c[0], bbar.c[0]);
#pragma unroll
  for (int i = 1; i < S; ++i) s0 = fma(cbar.c[i], a.c[i], s0);
  bbar.c[0] = s0;
#pragma unroll
  for (int i = 1; i < J::P0; ++i) bbar.c[i] = fma(cbar.c[i], a.c[0], bbar.c[i]);
#pragma unroll
  for (int k = 0; k < J::N2; ++k) {
    const int e = J::E0 + k, p = J::P0 + k;
    bbar.c[p] = fma(cbar.c[p], a.c[0], fma(T(2) * cbar.c[e], a.c[p], bbar.c[p]));
    bbar.c[e] = fma(cbar.c[e], a.c[0], bbar.c[e]);
  }
}

// c = f(a) with f0 = f(a0), f1 = f'(a0), f2 = f''(a0)
template <typename T, int S>
__device__ __forceinline__ Jet<T, S> jfunc(const Jet<T, S>& a, T f0, T f1, T f2) {
  using J = Jet<T, S>;
  J r;
  r.c[0] = f0;
#pragma unroll
  for (int i = 1; i < J::E0; ++i) r.c[i] = f1 * a.c[i];
#pragma unroll
  for (int k = 0; k < J::N2; ++k) {
    const int e = J::E0 + k, p = J::P0 + k;
    r.c[e] = fma(f1, a.c[e], f2 * a.c[p] * a.c[p]);
  }
  return r;
}

// Pullback of c = f(a):  abar += (d c / d a)^T cbar  (f3 = third derivative, needed by the
// second-order rows).
template <typename T, int S>
__device__ __forceinline__ void jfunc_pull_acc(Jet<T, S>& abar, const Jet<T, S>& cbar,
                                               const Jet<T, S>& a, T f1, T f2, T f3) {
  using J = Jet<T, S>;
  T s0 = cbar.c[0] * f1;
#pragma unroll
  for (int i = 1; i < J::E0; ++i) s0 = fma(cbar.c[i] * f2, a.c[i], s0);
#pragma unroll
  for (int k = 0; k < J::N2; ++k) {
    const int e = J::E0 + k, p = J::P0 + k;
    s0 = fma(cbar.c[e], fma(f2, a.c[e], f3 * a.c[p] * a.c[p]), s0);
  }
#pragma unroll
  for (int i = 1; i < J::P0; ++i) abar.c[i] = fma(cbar.c[i], f1, abar.c[i]);
#pragma unroll
  for (int k = 0; k < J::N2; ++k) {
    const int e = J::E0 + k, p = J::P0 + k;
    abar.c[p] = fma(cbar.c[p], f1, fma(T(2) * cbar.c[e] * f2, a.c[p], abar.c[p]));
    abar.c[e] = fma(cbar.c[e], f1, abar.c[e]);
  }
  abar.c[0] += s0;
}

// ---- scalar math per dtype ---------------------------------------------------------------
template <typename T>
struct Math;

template <>
struct Math<float> {
  static __device__ __forceinline__ float tanh_(float x) { return tanhf(x); }
  static __device__ __forceinline__ void sincos_(float x, float* s, float* c) { sincosf(x, s, c); }
};

template <>
struct Math<double> {
  static __device__ __forceinline__ double tanh_(double x) { return tanh(x); }
  static __device__ __forceinline__ void sincos_(double x, double* s, double* c) { sincos(x, s, c); }
};

// tanh and its first three derivatives at x
template <typename T>
__device__ __forceinline__ void tanh_derivs(T x, T& f0, T& f1, T& f2, T& f3) {
  f0 = Math<T>::tanh_(x);
  f1 = fma(-f0, f0, T(1));
  f2 = T(-2) * f0 * f1;
  f3 = f1 * fma(T(6) * f0, f0, T(-2));
}

// 4-wide shared-memory vector (LDS.128 for float, 2 x LDS.128 for double)
template <typename T>
struct alignas(sizeof(T) * 4) Vec4 {
  T v[4];
};

}  // namespace qcp
