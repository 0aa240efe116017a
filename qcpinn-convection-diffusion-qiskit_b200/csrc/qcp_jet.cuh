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
  static __device__ __forceinline__ float tanh_tab(float x, const float*) { return tanhf(x); }
  static __device__ __forceinline__ void sincos_(float x, float* s, float* c) { sincosf(x, s, c); }
};

// Branch-free double-precision tanh:  tanh|x| = 1 - 2 / (exp(2|x|) + 1).
// libm's tanh() takes one of two code paths per lane (|x| < 0.55: odd polynomial, else expm1-style),
// so a warp with mixed arguments executes both (~35 FP64-pipe + ~45 other instructions); this one
// is 24 FP64-pipe instructions and no branch.  exp: k = rint(y log2 e) by the magic-number trick,
// r = y - k ln2 in two pieces, Taylor polynomial of degree 13 on |r| <= 0.347 (truncation 5e-18),
// scaled by an exponent add (y <= 40, so k <= 58); reciprocal: MUFU seed + two Newton steps.
// Absolute error <= 2e-16 (it is a difference from 1, so tiny |x| lose RELATIVE accuracy; the
// activations enter every result additively next to O(1) terms, far inside the 1e-10 parity bar).
#ifndef QCP_FAST_TANH
#define QCP_FAST_TANH 1
#endif
__device__ __forceinline__ double tanh_branchfree(double x) {
  const double y = fmin(fabs(x) * 2.0, 40.0);
  const double magic = 6755399441055744.0;                    // 1.5 * 2^52
  const double kf = fma(y, 1.4426950408889634074, magic);
  const int k = __double2loint(kf);
  const double kd = kf - magic;
  double r = fma(kd, -6.93147180369123816490e-01, y);
  r = fma(kd, -1.90821492927058770002e-10, r);
  double p = 1.6059043836821614599e-10;                       // 1/13!
  p = fma(p, r, 2.0876756987868098979e-09);                   // 1/12!
  p = fma(p, r, 2.5052108385441718775e-08);                   // 1/11!
  p = fma(p, r, 2.7557319223985890653e-07);                   // 1/10!
  p = fma(p, r, 2.7557319223985892511e-06);                   // 1/9!
  p = fma(p, r, 2.4801587301587301566e-05);                   // 1/8!
  p = fma(p, r, 1.9841269841269841253e-04);                   // 1/7!
  p = fma(p, r, 1.3888888888888889419e-03);                   // 1/6!
  p = fma(p, r, 8.3333333333333332177e-03);                   // 1/5!
  p = fma(p, r, 4.1666666666666664354e-02);                   // 1/4!
  p = fma(p, r, 1.6666666666666665741e-01);                   // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double e = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
  const double d = e + 1.0;
  double inv;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(inv) : "d"(d));
  double t = fma(-d, inv, 1.0);
  inv = fma(inv, t, inv);
  t = fma(-d, inv, 1.0);
  inv = fma(inv, t, inv);
  const double th = copysign(fma(-2.0, inv, 1.0), x);
  return x != x ? x : th;                                     // NaN in, NaN out (like libm)
}

// The same with a 64-entry table of 2^(j/64) (shared memory): exp(y) = 2^m * tab[j] * exp(r) with
// 64 m + j = rint(64 y / ln 2) and |r| <= ln2/128, where a degree-5 Taylor polynomial is exact to
// 4e-17.  16 FP64-pipe instructions + one LDS instead of 24.
constexpr int kExpTabSize = 64;
__device__ __forceinline__ double tanh_table(double x, const double* __restrict__ tab) {
  const double y = fmin(fabs(x) * 2.0, 40.0);
  const double magic = 6755399441055744.0;                    // 1.5 * 2^52
  const double kf = fma(y, 92.33248261689366, magic);         // 64 / ln 2
  const int k = __double2loint(kf);
  const double kd = kf - magic;
  double r = fma(kd, -0.01083042469326756, y);                // ln2/64, high part (low 21 bits zero)
  r = fma(kd, -2.9815858269852933e-12, r);                   // ln2/64, low part
  double p = 8.3333333333333332177e-03;                       // 1/5!
  p = fma(p, r, 4.1666666666666664354e-02);
  p = fma(p, r, 1.6666666666666665741e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  p *= tab[k & (kExpTabSize - 1)];
  const double e = __hiloint2double(__double2hiint(p) + ((k >> 6) << 20), __double2loint(p));
  const double d = e + 1.0;
  double inv;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(inv) : "d"(d));
  double t = fma(-d, inv, 1.0);
  inv = fma(inv, t, inv);
  t = fma(-d, inv, 1.0);
  inv = fma(inv, t, inv);
  const double th = copysign(fma(-2.0, inv, 1.0), x);
  return x != x ? x : th;
}

template <>
struct Math<double> {
  static __device__ __forceinline__ double tanh_tab(double x, const double* tab) {
#if QCP_FAST_TANH
    return tanh_table(x, tab);
#else
    return tanh(x);
#endif
  }
  static __device__ __forceinline__ double tanh_(double x) {
#if QCP_FAST_TANH
    return tanh_branchfree(x);
#else
    return tanh(x);
#endif
  }
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
