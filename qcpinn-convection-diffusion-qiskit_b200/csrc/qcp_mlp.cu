// qcpinn_b200 -- generic-n (5..16 qubits) pre / post MLP stages of engine L.
//
// Same maths as the fused n <= 4 kernels (reference nn/DVPDESolver.py:28-51 in Taylor mode) but
// with the qubit count a run-time value: jets live in per-thread local arrays, weights in shared
// memory, and the stages exchange jets with the statevector kernels through the saved-jet
// workspace ws[slot][j*S + c][B].  These stages are < 1 % of the work at n >= 10 (the 2^n
// statevector dominates), so they are written for clarity, not for the last FMA.
#include "qcp_point.cuh"
#include "qcp_state.cuh"

namespace qcp {

template <typename T>
struct GenWeights {
  Vec4<T>* w1b;   // [H] (w1[k,0..2], b1[k])
  T* w2;          // [H][n]  w2[j,k] stored k-major
  T* w3;          // [H][n]  w3[k,i]
  T* b3w4;        // [H][2]
  T* b2;          // [n]
  T* b4;          // [4]
};

__host__ __device__ inline size_t gen_weight_elems(int n, int H) {
  return (size_t)4 * H + 2 * (size_t)H * n + 2 * H + ((n + 3) & ~3) + 4;
}

template <typename T>
__device__ GenWeights<T> gen_carve(unsigned char* base, int n, int H) {
  GenWeights<T> g;
  T* p = reinterpret_cast<T*>(base);
  g.w1b = reinterpret_cast<Vec4<T>*>(p); p += 4 * H;
  g.w2 = p; p += (size_t)H * n;
  g.w3 = p; p += (size_t)H * n;
  g.b3w4 = p; p += 2 * H;
  g.b2 = p; p += (n + 3) & ~3;
  g.b4 = p;
  return g;
}

template <typename T>
__device__ void gen_load(const GenWeights<T>& g, const SolverArgs& a, int n) {
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
    v.v[0] = w1[k * 3]; v.v[1] = w1[k * 3 + 1]; v.v[2] = w1[k * 3 + 2]; v.v[3] = b1[k];
    g.w1b[k] = v;
    g.b3w4[2 * k] = b3[k];
    g.b3w4[2 * k + 1] = w4[k];
  }
  for (int e = threadIdx.x; e < H * n; e += blockDim.x) {
    const int k = e / n, j = e % n;
    g.w2[e] = w2[j * H + k];
    g.w3[e] = w3[e];
  }
  for (int j = threadIdx.x; j < n; j += blockDim.x) g.b2[j] = b2[j];
  if (threadIdx.x == 0) g.b4[0] = b4[0];
}

template <typename T, int S>
__device__ __forceinline__ void gws_store(T* ws, long long B, int n, int slot, long long p,
                                          const Jet<T, S>* v) {
  T* base = ws + (size_t)slot * n * S * B + p;
  for (int j = 0; j < n; ++j)
#pragma unroll
    for (int c = 0; c < S; ++c) base[(size_t)(j * S + c) * B] = v[j].c[c];
}

template <typename T, int S>
__device__ __forceinline__ void gws_load(const T* ws, long long B, int n, int slot, long long p,
                                         Jet<T, S>* v) {
  const T* base = ws + (size_t)slot * n * S * B + p;
  for (int j = 0; j < n; ++j)
#pragma unroll
    for (int c = 0; c < S; ++c) v[j].c[c] = base[(size_t)(j * S + c) * B];
}

template <typename T, int S>
__global__ void __launch_bounds__(kThreads)
gen_pre_forward_kernel(const SolverArgs a, int n) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  const GenWeights<T> g = gen_carve<T>(smem_raw, n, a.H);
  gen_load<T>(g, a, n);
  __syncthreads();
  const T* Xg = static_cast<const T*>(a.X);
  T* wsg = static_cast<T*>(a.ws);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < a.B; p += stride) {
    const T X[3] = {Xg[3 * p], Xg[3 * p + 1], Xg[3 * p + 2]};
    Jet<T, S> z[kMaxQubitsSv];
    for (int j = 0; j < n; ++j) { jzero(z[j]); z[j].c[0] = g.b2[j]; }
    for (int k = 0; k < a.H; ++k) {
      const Jet<T, S> av = pre_activation<T, S>(g.w1b[k], X);
      const T h0 = Math<T>::tanh_(av.c[0]);
      const T f1 = fma(-h0, h0, T(1)), f2 = T(-2) * h0 * f1;
      const Jet<T, S> h = jfunc(av, h0, f1, f2);
      for (int j = 0; j < n; ++j) jaxpy(z[j], g.w2[k * n + j], h);
    }
    gws_store<T, S>(wsg, a.B, n, 0, p, z);
  }
}

template <typename T, int S>
__global__ void __launch_bounds__(kThreads)
gen_post_forward_kernel(const SolverArgs a, int n) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  const GenWeights<T> g = gen_carve<T>(smem_raw, n, a.H);
  gen_load<T>(g, a, n);
  __syncthreads();
  const T* wsg = static_cast<const T*>(a.ws);
  T* ug = static_cast<T*>(a.u);
  T* rg = static_cast<T*>(a.r);
  T* sg = static_cast<T*>(a.streams);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < a.B; p += stride) {
    Jet<T, S> q[kMaxQubitsSv], u;
    gws_load<T, S>(wsg, a.B, n, 1, p, q);
    jzero(u);
    u.c[0] = g.b4[0];
    for (int k = 0; k < a.H; ++k) {
      Jet<T, S> pj;
      jzero(pj);
      pj.c[0] = g.b3w4[2 * k];
      for (int i = 0; i < n; ++i) jaxpy(pj, g.w3[k * n + i], q[i]);
      const T g0 = Math<T>::tanh_(pj.c[0]);
      const T f1 = fma(-g0, g0, T(1)), f2 = T(-2) * g0 * f1;
      jaxpy(u, g.b3w4[2 * k + 1], jfunc(pj, g0, f1, f2));
    }
    ug[p] = u.c[0];
    if constexpr (S == 6) {
      if (rg)
        rg[p] = T(a.pde.ct) * u.c[1] + T(a.pde.cx) * u.c[2] + T(a.pde.cy) * u.c[3] +
                T(a.pde.cxx) * u.c[4] + T(a.pde.cyy) * u.c[5];
      if (sg)
        for (int c = 0; c < 6; ++c) sg[6 * p + c] = u.c[c];
    }
  }
}

// accumulator order: [b4 | k: w3[k,0..n-1], b3[k], w4[k]]
template <typename T, int S>
__global__ void __launch_bounds__(kThreads)
gen_post_backward_kernel(const SolverArgs a, int n) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  const int H = a.H, nacc = nacc_post(n, H);
  const GenWeights<T> g = gen_carve<T>(smem_raw, n, H);
  T* acc_all = reinterpret_cast<T*>(smem_raw + ((sizeof(T) * gen_weight_elems(n, H) + 31) & ~size_t(31)));
  T* tile_all = acc_all + (size_t)kWarpsPerBlock * nacc;
  gen_load<T>(g, a, n);
  for (int i = threadIdx.x; i < kWarpsPerBlock * nacc; i += blockDim.x) acc_all[i] = T(0);
  __syncthreads();
  Stager<T> st = make_stager<T>(acc_all, tile_all, nacc);
  T* wsg = static_cast<T*>(a.ws);
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long Bpad = (a.B + 31) & ~31LL;
  for (long long p0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; p0 < Bpad; p0 += stride) {
    const bool valid = p0 < a.B;
    const long long p = valid ? p0 : a.B - 1;
    const Jet<T, S> ub = seed_cotangent<T, S>(a, p, valid);
    Jet<T, S> q[kMaxQubitsSv], qb[kMaxQubitsSv];
    gws_load<T, S>(wsg, a.B, n, 1, p, q);
    for (int i = 0; i < n; ++i) jzero(qb[i]);
    st.begin();
    st.put(ub.c[0]);
    for (int k = 0; k < H; ++k) {
      Jet<T, S> pj;
      jzero(pj);
      pj.c[0] = g.b3w4[2 * k];
      for (int i = 0; i < n; ++i) jaxpy(pj, g.w3[k * n + i], q[i]);
      T g0, f1, f2, f3;
      tanh_derivs(pj.c[0], g0, f1, f2, f3);
      const Jet<T, S> gj = jfunc(pj, g0, f1, f2);
      const T w4 = g.b3w4[2 * k + 1];
      Jet<T, S> gb, pb;
#pragma unroll
      for (int c = 0; c < S; ++c) gb.c[c] = w4 * ub.c[c];
      jzero(pb);
      jfunc_pull_acc(pb, gb, pj, f1, f2, f3);
      st.reserve(n + 2);
      for (int i = 0; i < n; ++i) st.put(jdot(pb, q[i]));
      st.put(pb.c[0]);
      st.put(jdot(ub, gj));
      for (int i = 0; i < n; ++i) jaxpy(qb[i], g.w3[k * n + i], pb);
    }
    st.flush();
    if (valid) gws_store<T, S>(wsg, a.B, n, 1, p, qb);
  }
  write_partials<T>(acc_all, nacc, static_cast<T*>(a.partials));
}

// accumulator order: [b2[0..n-1] | k: w1[k,0..2], b1[k], w2[0..n-1,k]]
template <typename T, int S>
__global__ void __launch_bounds__(kThreads)
gen_pre_backward_kernel(const SolverArgs a, int n) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  const int H = a.H, nacc = nacc_pre(n, H);
  const GenWeights<T> g = gen_carve<T>(smem_raw, n, H);
  T* acc_all = reinterpret_cast<T*>(smem_raw + ((sizeof(T) * gen_weight_elems(n, H) + 31) & ~size_t(31)));
  T* tile_all = acc_all + (size_t)kWarpsPerBlock * nacc;
  gen_load<T>(g, a, n);
  for (int i = threadIdx.x; i < kWarpsPerBlock * nacc; i += blockDim.x) acc_all[i] = T(0);
  __syncthreads();
  Stager<T> st = make_stager<T>(acc_all, tile_all, nacc);
  const T* Xg = static_cast<const T*>(a.X);
  T* gXg = static_cast<T*>(a.gX);
  const T* wsg = static_cast<const T*>(a.ws);
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long Bpad = (a.B + 31) & ~31LL;
  for (long long p0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; p0 < Bpad; p0 += stride) {
    const bool valid = p0 < a.B;
    const long long p = valid ? p0 : a.B - 1;
    const T X[3] = {Xg[3 * p], Xg[3 * p + 1], Xg[3 * p + 2]};
    Jet<T, S> zb[kMaxQubitsSv];
    gws_load<T, S>(wsg, a.B, n, 0, p, zb);
    if (!valid)
      for (int j = 0; j < n; ++j) jzero(zb[j]);
    st.begin();
    st.reserve(n);
    for (int j = 0; j < n; ++j) st.put(zb[j].c[0]);
    T Xb[3] = {T(0), T(0), T(0)};
    for (int k = 0; k < H; ++k) {
      const Vec4<T> w = g.w1b[k];
      const Jet<T, S> av = pre_activation<T, S>(w, X);
      T h0, f1, f2, f3;
      tanh_derivs(av.c[0], h0, f1, f2, f3);
      const Jet<T, S> h = jfunc(av, h0, f1, f2);
      Jet<T, S> hb, ab;
      jzero(hb);
      for (int j = 0; j < n; ++j) jaxpy(hb, g.w2[k * n + j], zb[j]);
      jzero(ab);
      jfunc_pull_acc(ab, hb, av, f1, f2, f3);
      st.reserve(4 + n);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        T gd = ab.c[0] * X[d];
        if constexpr (S == 6) gd += ab.c[1 + d];
        st.put(gd);
        Xb[d] = fma(ab.c[0], w.v[d], Xb[d]);
      }
      st.put(ab.c[0]);
      for (int j = 0; j < n; ++j) st.put(jdot(zb[j], h));
    }
    st.flush();
    if (gXg && valid) { gXg[3 * p] = Xb[0]; gXg[3 * p + 1] = Xb[1]; gXg[3 * p + 2] = Xb[2]; }
  }
  write_partials<T>(acc_all, nacc, static_cast<T*>(a.partials));
}

// ---------------------------------------------------------------------------------------------
size_t mlp_smem_bytes(int dtype, int n, int H, int nacc) {
  const size_t es = dtype == QCP_F64 ? 8 : 4;
  size_t w = (es * gen_weight_elems(n, H) + 31) & ~size_t(31);
  if (nacc > 0) w += es * ((size_t)kWarpsPerBlock * nacc + (size_t)kWarpsPerBlock * kStageRows * kStagePitch);
  return w;
}

template <typename K>
static int gen_launch(K kernel, const MlpLaunch& L, size_t smem, const char* what, cudaStream_t s) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("%s: smem opt-in failed: %s", what, cudaGetErrorString(e)); return 1; }
  }
  kernel<<<L.grid, kThreads, smem, s>>>(L.args, L.n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("%s: launch failed: %s", what, cudaGetErrorString(e)); return 1; }
  return 0;
}

#define QCP_GEN_DISPATCH(KERNEL, NACC, WHAT)                                                    \
  const size_t smem = mlp_smem_bytes(dtype, L.n, L.H, NACC);                                    \
  if (dtype == QCP_F64)                                                                         \
    return S == 6 ? gen_launch(&KERNEL<double, 6>, L, smem, WHAT, s)                            \
                  : gen_launch(&KERNEL<double, 1>, L, smem, WHAT, s);                           \
  return S == 6 ? gen_launch(&KERNEL<float, 6>, L, smem, WHAT, s)                               \
                : gen_launch(&KERNEL<float, 1>, L, smem, WHAT, s);

int mlp_pre_forward(int dtype, int S, const MlpLaunch& L, cudaStream_t s) {
  QCP_GEN_DISPATCH(gen_pre_forward_kernel, 0, "gen_pre_forward")
}
int mlp_post_forward(int dtype, int S, const MlpLaunch& L, cudaStream_t s) {
  QCP_GEN_DISPATCH(gen_post_forward_kernel, 0, "gen_post_forward")
}
int mlp_post_backward(int dtype, int S, const MlpLaunch& L, cudaStream_t s) {
  QCP_GEN_DISPATCH(gen_post_backward_kernel, nacc_post(L.n, L.H), "gen_post_backward")
}
int mlp_pre_backward(int dtype, int S, const MlpLaunch& L, cudaStream_t s) {
  QCP_GEN_DISPATCH(gen_pre_backward_kernel, nacc_pre(L.n, L.H), "gen_pre_backward")
}

int mlp_backward_grid(int dtype, int S, int n, int H, int num_sms) {
  (void)dtype; (void)S; (void)n; (void)H;
  return num_sms * 2;
}

}  // namespace qcp
