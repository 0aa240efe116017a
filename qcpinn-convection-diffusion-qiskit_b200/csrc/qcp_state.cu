// qcpinn_b200 -- "engine L": per-sample statevector path for 5 <= n <= 16 qubits.
//
// For n <= 4 the observables are pre-multiplied into a 3^n feature matrix (qcp_point.cuh); beyond
// that 3^n outgrows 2^n x gates, so each collocation point carries its own 2^n statevector
// (reference nn/DVQuantumLayer.py:176-214 executed by default.qubit), here with S Taylor streams
// (psi, psi_t, psi_x, psi_y, psi_xx, psi_yy) pushed through the SAME batch-shared gates (they are
// linear), and an adjoint-method backward:
//
//   forward   z jets --RX(z_j) jet gates--> psi streams --gate program--> q_i streams
//             q = <psi|Z|psi>, q_d = 2Re<psi_d|Z|psi>, q_dd = 2Re<psi_dd|Z|psi> + 2<psi_d|Z|psi_d>
//   backward  lambda streams from the q cotangents, then the gate program in reverse: for every
//             parametrised gate dL/dtheta = 1/2 sum_streams Im<lambda|H|psi> (both after the gate),
//             un-apply the gate to psi and lambda; finally the encoding gates in reverse give the
//             cotangent jets of z.
//
// One CTA owns one point at a time (persistent grid-stride loop); its 2*S state vectors live in
// shared memory when they fit (n <= 10 in complex64 residual mode) and otherwise in a per-CTA slab
// of a plan-owned global workspace that stays L2 resident.  The MLPs around the circuit run as
// separate generic-n kernels that exchange Taylor jets through the saved-jet workspace
// (same layout as the n <= 4 split backward).  This is the correctness-first version of the path:
// gates are applied one by one (no diagonal / Kronecker fusion, no tcgen05 panel GEMM yet).
#include <cstdio>

#include "qcp_common.cuh"
#include "qcp_jet.cuh"
#include "qcp_state.cuh"

namespace qcp {

constexpr int kSvThreads = 256;
constexpr int kMaxOpsSv = 1024;

template <typename T>
struct Cx {
  T x, y;
};
template <typename T>
__device__ __forceinline__ Cx<T> cmul(Cx<T> a, Cx<T> b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T>
__device__ __forceinline__ Cx<T> cadd(Cx<T> a, Cx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T>
__device__ __forceinline__ Cx<T> csub(Cx<T> a, Cx<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T>
__device__ __forceinline__ Cx<T> cscale(T s, Cx<T> a) { return {s * a.x, s * a.y}; }
template <typename T>
__device__ __forceinline__ Cx<T> cconj(Cx<T> a) { return {a.x, -a.y}; }
template <typename T>
__device__ __forceinline__ T cre_conj_mul(Cx<T> a, Cx<T> b) { return a.x * b.x + a.y * b.y; }   // Re(conj(a) b)
template <typename T>
__device__ __forceinline__ T cim_conj_mul(Cx<T> a, Cx<T> b) { return a.x * b.y - a.y * b.x; }   // Im(conj(a) b)

template <typename T>
struct M2 {
  Cx<T> m[4];
};
template <typename T>
__device__ __forceinline__ M2<T> m2_dagger(const M2<T>& u) {
  return {{cconj(u.m[0]), cconj(u.m[2]), cconj(u.m[1]), cconj(u.m[3])}};
}
template <typename T>
__device__ __forceinline__ void m2_apply(const M2<T>& u, Cx<T>& a0, Cx<T>& a1) {
  const Cx<T> b0 = cadd(cmul(u.m[0], a0), cmul(u.m[1], a1));
  const Cx<T> b1 = cadd(cmul(u.m[2], a0), cmul(u.m[3], a1));
  a0 = b0;
  a1 = b1;
}

template <typename T>
__device__ M2<T> gate_matrix(int kind, double theta) {
  double s, c;
  sincos(0.5 * theta, &s, &c);
  const T C = (T)c, Sn = (T)s, Z = T(0);
  switch (kind) {
    case QCP_GATE_RX: case QCP_GATE_CRX: return {{{C, Z}, {Z, -Sn}, {Z, -Sn}, {C, Z}}};
    case QCP_GATE_RY: return {{{C, Z}, {-Sn, Z}, {Sn, Z}, {C, Z}}};
    case QCP_GATE_RZ: case QCP_GATE_CRZ: return {{{C, -Sn}, {Z, Z}, {Z, Z}, {C, Sn}}};
    case QCP_GATE_CNOT: return {{{Z, Z}, {T(1), Z}, {T(1), Z}, {Z, Z}}};
    default: {
      const T h = (T)0.70710678118654752440;
      return {{{h, Z}, {h, Z}, {h, Z}, {-h, Z}}};
    }
  }
}

__device__ __forceinline__ int ins0(int r, int pos) {
  return ((r >> pos) << (pos + 1)) | (r & ((1 << pos) - 1));
}

// block-wide sum (all threads get the result); scratch: >= blockDim/32 doubles
__device__ double block_reduce(double v, double* scratch) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += scratch[w];
  return s;
}

// ---------------------------------------------------------------------------------------------
// batch-shared gates on NS state vectors stored as st[s*M + k]
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ void apply_shared(Cx<T>* st, int NS, int n, const GateOp op, const M2<T>& mat,
                             const double2* consts, bool dag) {
  const int M = 1 << n;
  if (op.kind == QCP_GATE_U4) {
    const int pa = n - 1 - op.a, pb = n - 1 - op.b;
    const int plo = pa < pb ? pa : pb, phi = pa < pb ? pb : pa;
    const double2* U = consts + 16 * op.p;
    const int per = M >> 2, items = NS * per;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      Cx<T>* base = st + (size_t)(it / per) * M;
      const int k = ins0(ins0(it % per, plo), phi);
      const int idx[4] = {k, k | (1 << pb), k | (1 << pa), k | (1 << pa) | (1 << pb)};
      Cx<T> v[4], o[4];
      for (int j = 0; j < 4; ++j) v[j] = base[idx[j]];
      for (int i = 0; i < 4; ++i) {
        Cx<T> acc = {T(0), T(0)};
        for (int j = 0; j < 4; ++j) {
          const double2 e = dag ? U[j * 4 + i] : U[i * 4 + j];
          const Cx<T> uij = {(T)e.x, dag ? (T)(-e.y) : (T)e.y};
          acc = cadd(acc, cmul(uij, v[j]));
        }
        o[i] = acc;
      }
      for (int j = 0; j < 4; ++j) base[idx[j]] = o[j];
    }
  } else if (op.kind == QCP_GATE_CZ) {
    // diag(1, 1, 1, -1) on (a, b): its own inverse
    const int pa = n - 1 - op.a, pb = n - 1 - op.b;
    const int both = (1 << pa) | (1 << pb);
    for (int it = threadIdx.x; it < NS * M; it += blockDim.x) {
      if (((it & (M - 1)) & both) == both) {
        Cx<T>& v = st[it];
        v = {-v.x, -v.y};
      }
    }
  } else {
    const bool ctl = op.kind == QCP_GATE_CRX || op.kind == QCP_GATE_CRZ || op.kind == QCP_GATE_CNOT;
    const int pt = n - 1 - (ctl ? op.b : op.a);
    const int pc = ctl ? n - 1 - op.a : -1;
    const M2<T> u = dag ? m2_dagger(mat) : mat;
    const int per = M >> 1, items = NS * per;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      const int k0 = ins0(it % per, pt);
      if (ctl && !((k0 >> pc) & 1)) continue;
      Cx<T>* base = st + (size_t)(it / per) * M;
      Cx<T> a0 = base[k0], a1 = base[k0 | (1 << pt)];
      m2_apply(u, a0, a1);
      base[k0] = a0;
      base[k0 | (1 << pt)] = a1;
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// per-sample RX(z) "jet gate" of the angle encoding.  G = RX(z0); dG = dG/dz = (-i/2) X G;
// d2G = -G/4; d3G = -dG/4.  Streams: 0 value | 1..3 first order (t,x,y) | 4,5 second order (x,y).
// ---------------------------------------------------------------------------------------------
template <typename T>
struct RxJet {
  M2<T> G, dG;
  T zd[3], zdd[2];   // first / second derivatives of the angle along (t,x,y) / (x,y)
};

template <typename T, int S>
__device__ RxJet<T> make_rx_jet(const T* zj /* S comps */) {
  RxJet<T> r;
  T s, c;
  Math<T>::sincos_(T(0.5) * zj[0], &s, &c);
  r.G = {{{c, T(0)}, {T(0), -s}, {T(0), -s}, {c, T(0)}}};
  const T hs = T(0.5) * s, hc = T(0.5) * c;
  r.dG = {{{-hs, T(0)}, {T(0), -hc}, {T(0), -hc}, {-hs, T(0)}}};
  for (int d = 0; d < 3; ++d) r.zd[d] = S == 6 ? zj[1 + d] : T(0);
  for (int e = 0; e < 2; ++e) r.zdd[e] = S == 6 ? zj[4 + e] : T(0);
  return r;
}

// Per-sample Pauli rotation exp(-i a P / 2) anywhere in the program, a = scale * z (jets scaled
// alike).  Every Pauli rotation satisfies dG/da = (-i/2) P G, d2G = -G/4, d3G = -dG/4, which is all
// rx_jet_forward / rx_jet_backward use, so the same two routines serve RY and RZ; the cotangents
// they return are with respect to the jet of a and are scaled back by the caller.
template <typename T, int S>
__device__ RxJet<T> make_sample_jet(int kind, T scale, const T* zj /* S comps */) {
  RxJet<T> r;
  T s, c;
  Math<T>::sincos_(T(0.5) * scale * zj[0], &s, &c);
  const T hs = T(0.5) * s, hc = T(0.5) * c, Z = T(0);
  if (kind == QCP_GATE_RY_IN) {
    r.G = {{{c, Z}, {-s, Z}, {s, Z}, {c, Z}}};
    r.dG = {{{-hs, Z}, {-hc, Z}, {hc, Z}, {-hs, Z}}};
  } else {   // QCP_GATE_RZ_IN: diag(e^{-ia/2}, e^{+ia/2})
    r.G = {{{c, -s}, {Z, Z}, {Z, Z}, {c, s}}};
    r.dG = {{{-hs, -hc}, {Z, Z}, {Z, Z}, {-hs, hc}}};
  }
  for (int d = 0; d < 3; ++d) r.zd[d] = S == 6 ? scale * zj[1 + d] : T(0);
  for (int e = 0; e < 2; ++e) r.zdd[e] = S == 6 ? scale * zj[4 + e] : T(0);
  return r;
}

__device__ __forceinline__ bool is_sample_gate(int kind) {
  return kind == QCP_GATE_RY_IN || kind == QCP_GATE_RZ_IN;
}

template <typename T>
__device__ __forceinline__ void m2_mulvec(const M2<T>& u, Cx<T> a0, Cx<T> a1, Cx<T>& b0, Cx<T>& b1) {
  b0 = cadd(cmul(u.m[0], a0), cmul(u.m[1], a1));
  b1 = cadd(cmul(u.m[2], a0), cmul(u.m[3], a1));
}

// forward: new = J(G) applied to the stream jet, one thread per amplitude pair, all streams
template <typename T, int S>
__device__ void rx_jet_forward(Cx<T>* st, int n, int wire, const RxJet<T>& g) {
  const int M = 1 << n, pt = n - 1 - wire;
  for (int it = threadIdx.x; it < (M >> 1); it += blockDim.x) {
    const int k0 = ins0(it, pt), k1 = k0 | (1 << pt);
    Cx<T> a0[S], a1[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { a0[s] = st[(size_t)s * M + k0]; a1[s] = st[(size_t)s * M + k1]; }
    Cx<T> G0[S], G1[S];                         // G a^s
#pragma unroll
    for (int s = 0; s < S; ++s) m2_mulvec(g.G, a0[s], a1[s], G0[s], G1[s]);
    if constexpr (S == 6) {
      Cx<T> D0[4], D1[4];                       // dG a^s for s = 0 (value), 1..3 -> index s
#pragma unroll
      for (int s = 0; s < 4; ++s) m2_mulvec(g.dG, a0[s], a1[s], D0[s], D1[s]);
#pragma unroll
      for (int d = 0; d < 3; ++d) {             // first order: G a^d + zd dG a^0
        G0[1 + d] = cadd(G0[1 + d], cscale(g.zd[d], D0[0]));
        G1[1 + d] = cadd(G1[1 + d], cscale(g.zd[d], D1[0]));
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {             // second: G a^dd + 2 zd dG a^d + (zdd dG - zd^2/4 G) a^0
        const T zd = g.zd[1 + e], zdd = g.zdd[e];
        Cx<T> r0 = cadd(cscale(T(2) * zd, D0[2 + e]), cscale(zdd, D0[0]));
        Cx<T> r1 = cadd(cscale(T(2) * zd, D1[2 + e]), cscale(zdd, D1[0]));
        // G0[0], G1[0] still hold G a^0 (value stream is not modified above)
        r0 = cadd(r0, cscale(T(-0.25) * zd * zd, G0[0]));
        r1 = cadd(r1, cscale(T(-0.25) * zd * zd, G1[0]));
        G0[4 + e] = cadd(G0[4 + e], r0);
        G1[4 + e] = cadd(G1[4 + e], r1);
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) { st[(size_t)s * M + k0] = G0[s]; st[(size_t)s * M + k1] = G1[s]; }
  }
  __syncthreads();
}

// reverse: psi streams -> before the gate; lambda streams -> before the gate; returns this
// thread's partial cotangents of the angle jet in zb[S] (caller block-reduces).
template <typename T, int S>
__device__ void rx_jet_backward(Cx<T>* psi, Cx<T>* lam, int n, int wire, const RxJet<T>& g,
                                double (&zb)[S]) {
  const int M = 1 << n, pt = n - 1 - wire;
  const M2<T> Gd = m2_dagger(g.G), dGd = m2_dagger(g.dG);
#pragma unroll
  for (int s = 0; s < S; ++s) zb[s] = 0.0;
  for (int it = threadIdx.x; it < (M >> 1); it += blockDim.x) {
    const int k0 = ins0(it, pt), k1 = k0 | (1 << pt);
    Cx<T> a0[S], a1[S], l0[S], l1[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      a0[s] = psi[(size_t)s * M + k0]; a1[s] = psi[(size_t)s * M + k1];
      l0[s] = lam[(size_t)s * M + k0]; l1[s] = lam[(size_t)s * M + k1];
    }
    // ---- un-apply: B = psi before the gate --------------------------------------------------
    Cx<T> B0[S], B1[S];
    m2_mulvec(Gd, a0[0], a1[0], B0[0], B1[0]);
    Cx<T> dB0[S], dB1[S];                        // dG B^s (needed below), s = 0..S-1
    m2_mulvec(g.dG, B0[0], B1[0], dB0[0], dB1[0]);
    if constexpr (S == 6) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {             // B^d = G^dag (a^d - zd dG B^0)
        const Cx<T> t0 = csub(a0[1 + d], cscale(g.zd[d], dB0[0]));
        const Cx<T> t1 = csub(a1[1 + d], cscale(g.zd[d], dB1[0]));
        m2_mulvec(Gd, t0, t1, B0[1 + d], B1[1 + d]);
        m2_mulvec(g.dG, B0[1 + d], B1[1 + d], dB0[1 + d], dB1[1 + d]);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {             // B^dd = G^dag (a^dd - 2 zd dG B^d - (zdd dG - zd^2/4 G) B^0)
        const T zd = g.zd[1 + e], zdd = g.zdd[e];
        Cx<T> GB0, GB1;
        m2_mulvec(g.G, B0[0], B1[0], GB0, GB1);
        Cx<T> t0 = csub(a0[4 + e], cadd(cscale(T(2) * zd, dB0[2 + e]), cscale(zdd, dB0[0])));
        Cx<T> t1 = csub(a1[4 + e], cadd(cscale(T(2) * zd, dB1[2 + e]), cscale(zdd, dB1[0])));
        t0 = cadd(t0, cscale(T(0.25) * zd * zd, GB0));
        t1 = cadd(t1, cscale(T(0.25) * zd * zd, GB1));
        m2_mulvec(Gd, t0, t1, B0[4 + e], B1[4 + e]);
        m2_mulvec(g.dG, B0[4 + e], B1[4 + e], dB0[4 + e], dB1[4 + e]);
      }
    }
    // ---- cotangents of the angle jet ---------------------------------------------------------
    // <L, V> := Re(conj(L0) V0 + conj(L1) V1)
    auto dotp = [](Cx<T> L0, Cx<T> L1, Cx<T> V0, Cx<T> V1) -> T {
      return cre_conj_mul(L0, V0) + cre_conj_mul(L1, V1);
    };
    T acc0 = dotp(l0[0], l1[0], dB0[0], dB1[0]);
    if constexpr (S == 6) {
      Cx<T> GB0, GB1;                            // G B^0 (d2G = -G/4, d3G = -dG/4)
      m2_mulvec(g.G, B0[0], B1[0], GB0, GB1);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const T zd = g.zd[d];
        // value: <L^d, dG B^d + zd d2G B^0>
        acc0 += dotp(l0[1 + d], l1[1 + d], dB0[1 + d], dB1[1 + d]) +
                T(-0.25) * zd * dotp(l0[1 + d], l1[1 + d], GB0, GB1);
        // first order: <L^d, dG B^0>
        T accd = dotp(l0[1 + d], l1[1 + d], dB0[0], dB1[0]);
        if (d >= 1) {
          const int e = d - 1;
          // + <L^dd, 2 dG B^d + 2 zd d2G B^0>
          accd += T(2) * dotp(l0[4 + e], l1[4 + e], dB0[1 + d], dB1[1 + d]) +
                  T(-0.5) * zd * dotp(l0[4 + e], l1[4 + e], GB0, GB1);
        }
        zb[1 + d] += (double)accd;
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const T zd = g.zd[1 + e], zdd = g.zdd[e];
        // value: <L^dd, dG B^dd + 2 zd d2G B^d + (zdd d2G + zd^2 d3G) B^0>
        Cx<T> GBd0, GBd1;
        m2_mulvec(g.G, B0[2 + e], B1[2 + e], GBd0, GBd1);
        acc0 += dotp(l0[4 + e], l1[4 + e], dB0[4 + e], dB1[4 + e]) +
                T(-0.5) * zd * dotp(l0[4 + e], l1[4 + e], GBd0, GBd1) +
                T(-0.25) * zdd * dotp(l0[4 + e], l1[4 + e], GB0, GB1) +
                T(-0.25) * zd * zd * dotp(l0[4 + e], l1[4 + e], dB0[0], dB1[0]);
        // second order: <L^dd, dG B^0>
        zb[4 + e] += (double)dotp(l0[4 + e], l1[4 + e], dB0[0], dB1[0]);
      }
    }
    zb[0] += (double)acc0;
    // ---- lambda before the gate ---------------------------------------------------------------
    Cx<T> N0[S], N1[S];
#pragma unroll
    for (int s = 0; s < S; ++s) m2_mulvec(Gd, l0[s], l1[s], N0[s], N1[s]);
    if constexpr (S == 6) {
      Cx<T> E0[6], E1[6];                        // dG^dag L^s
#pragma unroll
      for (int s = 1; s < 6; ++s) m2_mulvec(dGd, l0[s], l1[s], E0[s], E1[s]);
#pragma unroll
      for (int d = 0; d < 3; ++d) {             // lambda^0 += zd dG^dag L^d
        N0[0] = cadd(N0[0], cscale(g.zd[d], E0[1 + d]));
        N1[0] = cadd(N1[0], cscale(g.zd[d], E1[1 + d]));
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const T zd = g.zd[1 + e], zdd = g.zdd[e];
        // lambda^0 += (zdd dG^dag - zd^2/4 G^dag) L^dd ;  lambda^d += 2 zd dG^dag L^dd
        N0[0] = cadd(N0[0], cadd(cscale(zdd, E0[4 + e]), cscale(T(-0.25) * zd * zd, N0[4 + e])));
        N1[0] = cadd(N1[0], cadd(cscale(zdd, E1[4 + e]), cscale(T(-0.25) * zd * zd, N1[4 + e])));
        N0[2 + e] = cadd(N0[2 + e], cscale(T(2) * zd, E0[4 + e]));
        N1[2 + e] = cadd(N1[2 + e], cscale(T(2) * zd, E1[4 + e]));
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      psi[(size_t)s * M + k0] = B0[s]; psi[(size_t)s * M + k1] = B1[s];
      lam[(size_t)s * M + k0] = N0[s]; lam[(size_t)s * M + k1] = N1[s];
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
template <typename T>
struct SvArgs {
  int n, enc, n_ops, n_theta;
  const GateOp* ops;
  const double2* consts;
  const T* theta;
  const T* ws;          // saved-jet workspace: slot 0 = z jets, slot 1 = q jets / their cotangents
  T* ws_out;            // same buffer, writable
  long long B;
  Cx<T>* slab;          // per-CTA state storage in global memory (or null => shared memory)
  size_t slab_stride;   // elements per CTA
  double* theta_partials;   // [grid][n_theta]
};

template <typename T>
__device__ void build_table(M2<T>* table, const SvArgs<T>& a) {
  for (int g = threadIdx.x; g < a.n_ops; g += blockDim.x) {
    const GateOp op = a.ops[g];
    if (op.kind != QCP_GATE_U4 && op.kind != QCP_GATE_CZ && !is_sample_gate(op.kind))
      table[g] = gate_matrix<T>(op.kind, op.p >= 0 ? (double)a.theta[op.p] : 0.0);
  }
  __syncthreads();
}

// encoding -> psi streams (angle: RX jet gates on |0..0>; amplitude: normalised padded features)
template <typename T, int S>
__device__ void encode_forward(Cx<T>* psi, const SvArgs<T>& a, const T* zj /* [n][S] in smem */,
                               T* amp_scratch /* >= 3*S + n*S */) {
  const int n = a.n, M = 1 << n;
  for (int i = threadIdx.x; i < S * M; i += blockDim.x) psi[i] = {T(0), T(0)};
  __syncthreads();
  if (a.enc == QCP_ENC_NONE) {
    if (threadIdx.x == 0) psi[0] = {T(1), T(0)};
    __syncthreads();
  } else if (a.enc == QCP_ENC_ANGLE) {
    if (threadIdx.x == 0) psi[0] = {T(1), T(0)};
    __syncthreads();
    for (int j = 0; j < n; ++j) {
      const RxJet<T> g = make_rx_jet<T, S>(zj + j * S);
      rx_jet_forward<T, S>(psi, n, j, g);
    }
  } else {
    // psi0[a] = f_a / |f| for a < n (jets via the shared jet algebra), zero elsewhere
    if (threadIdx.x == 0) {
      Jet<T, S> nrm;
      jzero(nrm);
      for (int j = 0; j < n; ++j) {
        Jet<T, S> f;
        for (int c = 0; c < S; ++c) f.c[c] = zj[j * S + c];
        jmul_acc(nrm, f, f);
      }
      const T r = T(1) / sqrt(nrm.c[0]);          // g(x) = x^-1/2
      const T g1 = T(-0.5) * r / nrm.c[0], g2 = T(0.75) * r / (nrm.c[0] * nrm.c[0]);
      const Jet<T, S> inv = jfunc(nrm, r, g1, g2);
      for (int j = 0; j < n; ++j) {
        Jet<T, S> f;
        for (int c = 0; c < S; ++c) f.c[c] = zj[j * S + c];
        const Jet<T, S> e = jmul(f, inv);
        for (int c = 0; c < S; ++c) psi[(size_t)c * M + j] = {e.c[c], T(0)};
      }
    }
    __syncthreads();
  }
}

// the gate program on the psi streams: batch-shared gates, and per-sample jet gates in place
template <typename T, int S>
__device__ void run_program_forward(Cx<T>* psi, const SvArgs<T>& a, const M2<T>* table, const T* zj) {
  for (int g = 0; g < a.n_ops; ++g) {
    const GateOp op = a.ops[g];
    if (is_sample_gate(op.kind)) {
      const RxJet<T> jg = make_sample_jet<T, S>(op.kind, T(0.25) * (T)op.p, zj + op.b * S);
      rx_jet_forward<T, S>(psi, a.n, op.a, jg);
    } else {
      apply_shared<T>(psi, S, a.n, op, table[g], a.consts, false);
    }
  }
}

// q_i streams from the final psi streams; `tmp` = S*M reals of scratch (the lambda storage)
template <typename T, int S>
__device__ void measure(const Cx<T>* psi, T* tmp, int n, T* q_out /* smem [n][S] */) {
  const int M = 1 << n;
  for (int k = threadIdx.x; k < M; k += blockDim.x) {
    const Cx<T> p0 = psi[k];
    tmp[k] = p0.x * p0.x + p0.y * p0.y;
    if constexpr (S == 6) {
      Cx<T> pd[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        pd[d] = psi[(size_t)(1 + d) * M + k];
        tmp[(size_t)(1 + d) * M + k] = T(2) * cre_conj_mul(pd[d], p0);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const Cx<T> pdd = psi[(size_t)(4 + e) * M + k];
        tmp[(size_t)(4 + e) * M + k] =
            T(2) * cre_conj_mul(pdd, p0) + T(2) * (pd[1 + e].x * pd[1 + e].x + pd[1 + e].y * pd[1 + e].y);
      }
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int item = warp; item < n * S; item += nw) {
    const int i = item / S, s = item % S, pos = n - 1 - i;
    double acc = 0.0;
    for (int k = lane; k < M; k += 32) {
      const T v = tmp[(size_t)s * M + k];
      acc += ((k >> pos) & 1) ? -(double)v : (double)v;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) q_out[i * S + s] = (T)acc;
  }
  __syncthreads();
}

template <typename T, int S>
__global__ void __launch_bounds__(kSvThreads)
sv_forward_kernel(const SvArgs<T> a) {
  extern __shared__ __align__(16) unsigned char sv_smem[];
  const int n = a.n, M = 1 << n;
  M2<T>* table = reinterpret_cast<M2<T>*>(sv_smem);
  T* zj = reinterpret_cast<T*>(table + a.n_ops);
  T* qj = zj + n * S;
  T* scratch = qj + n * S;
  Cx<T>* psi = a.slab ? a.slab + (size_t)blockIdx.x * a.slab_stride
                      : reinterpret_cast<Cx<T>*>(scratch + 64);
  T* tmp = reinterpret_cast<T*>(psi + (size_t)S * M);
  build_table<T>(table, a);
  for (long long p = blockIdx.x; p < a.B; p += gridDim.x) {
    for (int e = threadIdx.x; e < n * S; e += blockDim.x) zj[e] = a.ws[(size_t)e * a.B + p];
    __syncthreads();
    encode_forward<T, S>(psi, a, zj, scratch);
    run_program_forward<T, S>(psi, a, table, zj);
    measure<T, S>(psi, tmp, n, qj);
    for (int e = threadIdx.x; e < n * S; e += blockDim.x)
      a.ws_out[(size_t)(n * S + e) * a.B + p] = qj[e];
    __syncthreads();
  }
}

template <typename T, int S>
__global__ void __launch_bounds__(kSvThreads)
sv_backward_kernel(const SvArgs<T> a) {
  extern __shared__ __align__(16) unsigned char sv_smem[];
  __shared__ double red[kSvThreads / 32];
  const int n = a.n, M = 1 << n;
  M2<T>* table = reinterpret_cast<M2<T>*>(sv_smem);
  T* zj = reinterpret_cast<T*>(table + a.n_ops);
  T* qb = zj + n * S;
  T* scratch = qb + n * S;
  double* gth = reinterpret_cast<double*>(scratch + 64);           // [n_theta]
  double* zacc = gth + a.n_theta;                                  // [n*S] z cotangents of jet gates
  Cx<T>* base = a.slab ? a.slab + (size_t)blockIdx.x * a.slab_stride
                       : reinterpret_cast<Cx<T>*>(zacc + n * S);
  Cx<T>* psi = base;
  Cx<T>* lam = base + (size_t)S * M;
  build_table<T>(table, a);
  for (int p = threadIdx.x; p < a.n_theta; p += blockDim.x) gth[p] = 0.0;
  __syncthreads();

  for (long long p = blockIdx.x; p < a.B; p += gridDim.x) {
    for (int e = threadIdx.x; e < n * S; e += blockDim.x) {
      zj[e] = a.ws[(size_t)e * a.B + p];
      qb[e] = a.ws[(size_t)(n * S + e) * a.B + p];
      zacc[e] = 0.0;
    }
    __syncthreads();
    // ---- recompute the forward state ------------------------------------------------------
    encode_forward<T, S>(psi, a, zj, scratch);
    run_program_forward<T, S>(psi, a, table, zj);
    // ---- lambda streams from the q cotangents (lambda = 2 dL/d conj(psi)) --------------------
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
      T zb[S];
#pragma unroll
      for (int s = 0; s < S; ++s) zb[s] = T(0);
      for (int i = 0; i < n; ++i) {
        const T sg = ((k >> (n - 1 - i)) & 1) ? T(-1) : T(1);
#pragma unroll
        for (int s = 0; s < S; ++s) zb[s] = fma(sg, qb[i * S + s], zb[s]);
      }
      const Cx<T> p0 = psi[k];
      Cx<T> l0 = cscale(T(2) * zb[0], p0);
      if constexpr (S == 6) {
        Cx<T> pd[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          pd[d] = psi[(size_t)(1 + d) * M + k];
          l0 = cadd(l0, cscale(T(2) * zb[1 + d], pd[d]));
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const Cx<T> pdd = psi[(size_t)(4 + e) * M + k];
          l0 = cadd(l0, cscale(T(2) * zb[4 + e], pdd));
          lam[(size_t)(4 + e) * M + k] = cscale(T(2) * zb[4 + e], p0);
        }
        lam[(size_t)1 * M + k] = cscale(T(2) * zb[1], p0);
#pragma unroll
        for (int e = 0; e < 2; ++e)
          lam[(size_t)(2 + e) * M + k] =
              cadd(cscale(T(2) * zb[2 + e], p0), cscale(T(4) * zb[4 + e], pd[1 + e]));
      }
      lam[k] = l0;
    }
    __syncthreads();
    // ---- gate program in reverse ------------------------------------------------------------
    for (int g = a.n_ops - 1; g >= 0; --g) {
      const GateOp op = a.ops[g];
      if (is_sample_gate(op.kind)) {
        const T scale = T(0.25) * (T)op.p;
        const RxJet<T> jg = make_sample_jet<T, S>(op.kind, scale, zj + op.b * S);
        double zb[S];
        rx_jet_backward<T, S>(psi, lam, n, op.a, jg, zb);
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const double tot = block_reduce(zb[s], red);
          if (threadIdx.x == 0) zacc[op.b * S + s] += (double)scale * tot;
        }
        __syncthreads();
        continue;
      }
      if (op.p >= 0 && op.kind != QCP_GATE_U4) {
        const bool ctl = op.kind == QCP_GATE_CRX || op.kind == QCP_GATE_CRZ;
        const int pt = n - 1 - (ctl ? op.b : op.a);
        const int pc = ctl ? n - 1 - op.a : -1;
        double part = 0.0;
        for (int it = threadIdx.x; it < S * M; it += blockDim.x) {
          const int k = it % M;
          if (ctl && !((k >> pc) & 1)) continue;
          const size_t off = (size_t)(it / M) * M;
          const int bit = (k >> pt) & 1;
          Cx<T> h;
          if (op.kind == QCP_GATE_RX || op.kind == QCP_GATE_CRX) {
            h = psi[off + (k ^ (1 << pt))];
          } else if (op.kind == QCP_GATE_RY) {
            const Cx<T> o = psi[off + (k ^ (1 << pt))];
            h = bit ? Cx<T>{-o.y, o.x} : Cx<T>{o.y, -o.x};
          } else {
            const Cx<T> o = psi[off + k];
            h = bit ? Cx<T>{-o.x, -o.y} : o;
          }
          part += (double)cim_conj_mul(lam[off + k], h);
        }
        const double tot = block_reduce(part, red);
        if (threadIdx.x == 0) gth[op.p] += 0.5 * tot;
      }
      apply_shared<T>(psi, 2 * S, n, op, table[g], a.consts, true);   // psi and lambda are contiguous
    }
    // ---- encoding in reverse -> cotangent jets of z --------------------------------------------
    if (a.enc == QCP_ENC_NONE) {
      for (int e = threadIdx.x; e < n * S; e += blockDim.x) qb[e] = T(0);
      __syncthreads();
    } else if (a.enc == QCP_ENC_ANGLE) {
      for (int j = n - 1; j >= 0; --j) {
        const RxJet<T> g = make_rx_jet<T, S>(zj + j * S);
        double zb[S];
        rx_jet_backward<T, S>(psi, lam, n, j, g, zb);
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const double tot = block_reduce(zb[s], red);
          if (threadIdx.x == 0) qb[j * S + s] = (T)tot;            // reuse qb as the zb staging
        }
      }
      __syncthreads();
    } else {
      // psi0_a = f_a * inv(|f|^2): pull lambda (real parts of the first n amplitudes) back
      if (threadIdx.x == 0) {
        Jet<T, S> nrm, fj[kMaxQubitsSv], fb[kMaxQubitsSv], invb;
        jzero(nrm);
        jzero(invb);
        for (int j = 0; j < n; ++j) {
          for (int c = 0; c < S; ++c) fj[j].c[c] = zj[j * S + c];
          jzero(fb[j]);
          jmul_acc(nrm, fj[j], fj[j]);
        }
        const T r = T(1) / sqrt(nrm.c[0]);
        const T g1 = T(-0.5) * r / nrm.c[0], g2 = T(0.75) * r / (nrm.c[0] * nrm.c[0]);
        const T g3 = T(-1.875) * r / (nrm.c[0] * nrm.c[0] * nrm.c[0]);
        const Jet<T, S> inv = jfunc(nrm, r, g1, g2);
        for (int j = 0; j < n; ++j) {
          Jet<T, S> eb;
          for (int c = 0; c < S; ++c) eb.c[c] = lam[(size_t)c * M + j].x;
          jmul_pull_acc(fb[j], eb, inv);
          jmul_pull_acc(invb, eb, fj[j]);
        }
        Jet<T, S> nb;
        jzero(nb);
        jfunc_pull_acc(nb, invb, nrm, g1, g2, g3);
        for (int j = 0; j < n; ++j) {
          jmul_pull_acc(fb[j], nb, fj[j]);
          jmul_pull_acc(fb[j], nb, fj[j]);
          for (int c = 0; c < S; ++c) qb[j * S + c] = fb[j].c[c];
        }
      }
      __syncthreads();
    }
    for (int e = threadIdx.x; e < n * S; e += blockDim.x)
      a.ws_out[(size_t)e * a.B + p] = (T)((double)qb[e] + zacc[e]);
    __syncthreads();
  }
  double* out = a.theta_partials + (size_t)blockIdx.x * a.n_theta;
  for (int p = threadIdx.x; p < a.n_theta; p += blockDim.x) out[p] = gth[p];
}

template <typename T>
__global__ void sv_reduce_theta_kernel(const double* partials, int grid, int n_theta, T* gtheta) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_theta) return;
  double s = 0.0;
  for (int g = 0; g < grid; ++g) s += partials[(size_t)g * n_theta + p];
  gtheta[p] = (T)s;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename T>
static size_t sv_small_smem(int n, int S, int n_ops, int n_theta, bool backward) {
  size_t b = sizeof(M2<T>) * (size_t)n_ops + sizeof(T) * ((size_t)2 * n * S + 64);
  if (backward) b += sizeof(double) * ((size_t)n_theta + (size_t)n * S);
  return (b + 15) & ~size_t(15);
}

template <typename T>
size_t sv_state_bytes(int n, int S, bool backward) {
  // forward: S psi streams + S*M reals of measurement scratch (<= S more complex streams)
  return sizeof(Cx<T>) * (size_t)(2 * S) << n;
}

template <typename T, int S>
static int sv_launch(bool backward, const SvLaunch& L, cudaStream_t s) {
  SvArgs<T> a{};
  a.n = L.n; a.enc = L.enc; a.n_ops = L.n_ops; a.n_theta = L.n_theta;
  a.ops = L.ops; a.consts = L.consts; a.theta = static_cast<const T*>(L.theta);
  a.ws = static_cast<const T*>(L.ws); a.ws_out = static_cast<T*>(L.ws);
  a.B = L.B; a.theta_partials = L.theta_partials;
  size_t smem = sv_small_smem<T>(L.n, S, L.n_ops, L.n_theta, backward);
  const size_t state = sv_state_bytes<T>(L.n, S, backward);
  if (L.slab) {
    a.slab = static_cast<Cx<T>*>(L.slab);
    a.slab_stride = state / sizeof(Cx<T>);
  } else {
    smem += state;
  }
  auto kernel = backward ? &sv_backward_kernel<T, S> : &sv_forward_kernel<T, S>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("sv kernel: cannot opt in to %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
      return 1;
    }
  }
  kernel<<<L.grid, kSvThreads, smem, s>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("sv %s kernel launch failed: %s", backward ? "backward" : "forward", cudaGetErrorString(e));
    return 1;
  }
  if (backward) {
    sv_reduce_theta_kernel<T><<<(L.n_theta + 127) / 128, 128, 0, s>>>(
        L.theta_partials, L.grid, L.n_theta, static_cast<T*>(L.grad_theta));
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("sv reduce launch failed: %s", cudaGetErrorString(e)); return 1; }
  }
  return 0;
}

size_t sv_state_bytes_rt(int dtype, int n, int S) {
  return dtype == QCP_F64 ? sv_state_bytes<double>(n, S, true) : sv_state_bytes<float>(n, S, true);
}

size_t sv_fixed_smem_rt(int dtype, int n, int S, int n_ops, int n_theta) {
  return dtype == QCP_F64 ? sv_small_smem<double>(n, S, n_ops, n_theta, true)
                          : sv_small_smem<float>(n, S, n_ops, n_theta, true);
}

int sv_run(int dtype, int S, bool backward, const SvLaunch& L, cudaStream_t s) {
  if (L.n_ops > kMaxOpsSv) { set_error("gate program too long for engine L (%d ops)", L.n_ops); return 1; }
  if (dtype == QCP_F64)
    return S == 6 ? sv_launch<double, 6>(backward, L, s) : sv_launch<double, 1>(backward, L, s);
  return S == 6 ? sv_launch<float, 6>(backward, L, s) : sv_launch<float, 1>(backward, L, s);
}

}  // namespace qcp
