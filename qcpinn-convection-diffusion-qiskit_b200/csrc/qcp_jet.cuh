// qcpinn_b200 -- truncated Taylor "jets" carried by one collocation point.
//
// The convection-diffusion residual (reference nn/pde.py:53-72) needs u, u_t, u_x, u_y, u_xx, u_yy.
// Instead of five nested autograd sweeps we push a 6-component jet through every operation:
//
//   c[0] value | c[1] d/dt | c[2] d/dx | c[3] d/dy | c[4] d2/dx2 | c[5] d2/dy2
//
// (no second t-derivative is needed).  S = 1 degenerates to the plain value (IC / BC points,
// reference trainer/diffusion_train.py:40-41).  Every forward rule has a hand-derived pullback so
// the backward kernel is the exact adjoint of the six-stream forward.
#pragma once

#include <cuda_runtime.h>

namespace qcp {

template <typename T, int S>
struct Jet {
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
  Jet<T, S> r;
  r.c[0] = a.c[0] * b.c[0];
  if constexpr (S == 6) {
#pragma unroll
    for (int i = 1; i <= 3; ++i) r.c[i] = fma(a.c[i], b.c[0], a.c[0] * b.c[i]);
#pragma unroll
    for (int e = 4; e <= 5; ++e) {
      T t = fma(a.c[e], b.c[0], a.c[0] * b.c[e]);
      r.c[e] = fma(T(2) * a.c[e - 2], b.c[e - 2], t);
    }
  }
  return r;
}

// acc += a * b
template <typename T, int S>
__device__ __forceinline__ void jmul_acc(Jet<T, S>& acc, const Jet<T, S>& a, const Jet<T, S>& b) {
  acc.c[0] = fma(a.c[0], b.c[0], acc.c[0]);
  if constexpr (S == 6) {
#pragma unroll
    for (int i = 1; i <= 3; ++i) acc.c[i] = fma(a.c[i], b.c[0], fma(a.c[0], b.c[i], acc.c[i]));
#pragma unroll
    for (int e = 4; e <= 5; ++e) {
      T t = fma(a.c[e], b.c[0], fma(a.c[0], b.c[e], acc.c[e]));
      acc.c[e] = fma(T(2) * a.c[e - 2], b.c[e - 2], t);
    }
  }
}

// Pullback of c = a*b w.r.t. b:  bbar += (d c / d b)^T cbar, with the other factor a.
template <typename T, int S>
__device__ __forceinline__ void jmul_pull_acc(Jet<T, S>& bbar, const Jet<T, S>& cbar,
                                              const Jet<T, S>& a) {
  bbar.c[0] = fma(cbar.c[0], a.c[0], bbar.c[0]);
  if constexpr (S == 6) {
#pragma unroll
    for (int i = 1; i <= 5; ++i) bbar.c[0] = fma(cbar.c[i], a.c[i], bbar.c[0]);
    bbar.c[1] = fma(cbar.c[1], a.c[0], bbar.c[1]);
#pragma unroll
    for (int i = 2; i <= 3; ++i)
      bbar.c[i] = fma(cbar.c[i], a.c[0], fma(T(2) * cbar.c[i + 2], a.c[i], bbar.c[i]));
#pragma unroll
    for (int e = 4; e <= 5; ++e) bbar.c[e] = fma(cbar.c[e], a.c[0], bbar.c[e]);
  }
}

// c = f(a) with f0 = f(a0), f1 = f'(a0), f2 = f''(a0)
template <typename T, int S>
__device__ __forceinline__ Jet<T, S> jfunc(const Jet<T, S>& a, T f0, T f1, T f2) {
  Jet<T, S> r;
  r.c[0] = f0;
  if constexpr (S == 6) {
#pragma unroll
    for (int i = 1; i <= 3; ++i) r.c[i] = f1 * a.c[i];
#pragma unroll
    for (int e = 4; e <= 5; ++e) r.c[e] = fma(f1, a.c[e], f2 * a.c[e - 2] * a.c[e - 2]);
  }
  return r;
}

// Pullback of c = f(a):  abar += (d c / d a)^T cbar   (needs f''' for the second-order rows).
template <typename T, int S>
__device__ __forceinline__ void jfunc_pull_acc(Jet<T, S>& abar, const Jet<T, S>& cbar,
                                               const Jet<T, S>& a, T f1, T f2, T f3) {
  T s0 = cbar.c[0] * f1;
  if constexpr (S == 6) {
#pragma unroll
    for (int i = 1; i <= 3; ++i) s0 = fma(cbar.c[i] * f2, a.c[i], s0);
#pragma unroll
    for (int e = 4; e <= 5; ++e)
      s0 = fma(cbar.c[e], fma(f2, a.c[e], f3 * a.c[e - 2] * a.c[e - 2]), s0);
    abar.c[1] = fma(cbar.c[1], f1, abar.c[1]);
#pragma unroll
    for (int i = 2; i <= 3; ++i)
      abar.c[i] = fma(cbar.c[i], f1, fma(T(2) * cbar.c[i + 2] * f2, a.c[i], abar.c[i]));
#pragma unroll
    for (int e = 4; e <= 5; ++e) abar.c[e] = fma(cbar.c[e], f1, abar.c[e]);
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
