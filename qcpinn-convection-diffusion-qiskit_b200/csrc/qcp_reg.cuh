// qcpinn_b200 -- "engine R": register-resident per-sample statevector path (5 <= n <= 10 qubits).
//
// Same mathematics as engine L (qcp_state.cu; reference nn/DVQuantumLayer.py:176-214 executed by
// default.qubit, S Taylor streams pushed through the batch-shared gates, adjoint-method backward)
// but laid out for the SM instead of for simplicity:
//
//   * ONE WARP OWNS ONE STREAM VECTOR.  The 2^n amplitudes of psi_s (and, in the backward, of its
//     adjoint lambda_s) live in registers: the low LB index bits ("local" positions) select one of
//     2^LB amplitudes of a lane, the remaining n-LB bits ("lane" positions) select the lane.
//     Streams only couple in the encoding and in the measurement, so the six warps of a residual
//     point run the gate program independently, with no __syncthreads per gate.
//   * Dense gates act on local positions only (pure register FMAs).  The host compiler
//     (qcp_reg.cu) tracks where every logical qubit currently lives and inserts SWAP ops that
//     exchange a local with a lane position (half of the amplitudes through warp shuffles) just
//     before a lane-resident qubit is needed, evicting the qubit whose next use is farthest away.
//   * Every run of commuting diagonal gates (RZ, CRZ -- e.g. the n(n-1) CRZs of cross_mesh,
//     reference nn/DVQuantumLayer.py:348-371) is ONE multiplication with a phase table built once
//     per parameter update; its parameter gradients come from W_k = sum Im(conj(lambda_k) psi_k),
//     which is invariant under the whole diagonal block, accumulated per CTA and projected on the
//     gate generators by a tiny kernel afterwards.
//   * The angle encoding RX(z_j)|0> is a product state: psi0[k] = (-i)^popcount(k) prod_j r_j[k_j]
//     with REAL one-qubit jets r_j = (cos z_j/2, sin z_j/2), so every stream of psi0 is a product
//     of two small real jet tables (lane part x local part), and its pullback is real as well.
//   * The forward saves the final psi streams (2^n complex per stream); the backward un-applies
//     the gates from there instead of recomputing the forward (it can recompute when no state
//     workspace was given).
//
// For n < 2 LB... a vector needs only G = 2^(n-LB) lanes, so one warp carries 32/G points.
#pragma once

#include "qcp_common.cuh"
#include "qcp_jet.cuh"
#include "qcp_state.cuh"

namespace qcp {
namespace rg {

enum RKind { R_L1 = 0, R_CX = 1, R_SWAP = 2, R_DIAG = 3, R_U4 = 4, R_PERM = 5, R_PERMB = 6 };
enum RType { T_X = 0, T_R = 1, T_Z = 2 };   // RX-like (c, -is; -is, c) | real 2x2 (RY, H) | RZ-like diag(c - is, c + is)

// physical op: positions are bit positions of the amplitude index (0..LB-1 local, LB.. lane)
struct ROp {
  int32_t kind;   // RKind
  int32_t pt;     // L1/CX: target (local)   SWAP: local position
  int32_t pc;     // L1/CX: control position or -1   SWAP: lane position
  int32_t type;   // L1: RType
  int32_t g;      // L1: original gate index (coefficient table)  DIAG: block  U4: const index
  int32_t p;      // L1: theta index for the gradient or -1
  int32_t m;      // L1/CX with a LOCAL control: bit h set = pair h (target bit removed) is active
  int32_t pad;
  // PERM / PERMB: two consecutive slots; pc .. pad (6 words) hold the PermMasks of the permutation
  // (PERM, executed by the forward program) and of its inverse (PERMB, executed in reverse)
};

struct DiagGate {   // one diagonal gate of a block (table builder + gradient projection)
  int32_t blk, kind, a, b, p;
};

// shared-memory carve-up (byte offsets; computed on the host, read from the constant bank)
struct SmemLayout {
  int rops, cs, u4, zj, qj, rj, tab, exch, xred, atab, tabbar, rbar, gth, wacc, total;
};

struct RgArgs {
  SmemLayout lay;
  int n, enc, n_rops, n_gates, n_theta, n_blk, n_consts;
  int meas_pos[kMaxQubitsReg];   // final position of logical qubit q
  const ROp* rops;
  const GateOp* gates;
  const double2* consts;
  const void* theta;             // T[n_theta]
  const void* diag;              // C2<T>[n_blk << n], index (blk << n) + i * G + lane
  void* ws;                      // saved-jet workspace [2][n*S][B]
  long long B;
  void* state;                   // final psi streams [B][S][2^n] complex, or null
  double* theta_partials;        // [grid * warps][n_theta]     one row per WARP: plain read-modify-
  void* w_partials;              // T[grid][n_blk << n]  one row per CTA, one writing thread per entry
};

template <typename T>
struct C2 {
  T x, y;
};
template <typename T>
struct alignas(2 * sizeof(T)) C2A {
  T x, y;
};

template <typename T>
__device__ __forceinline__ T shx(T v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }

// WV = warps per stream vector: 1 (n - LB <= 5 lane bits) or 2 (six lane bits: float64 at n = 10; the
// sixth lane bit is the warp parity, relayouts go through a pair-wide shared-memory PERM)
__host__ __device__ constexpr int rg_warps(int S, int WV = 1) { return S == 6 ? 6 * WV : 4; }

// ---------------------------------------------------------------------------------------------
// shared-memory carve-up (identical on host and device)
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int rg_align(size_t v) { return (int)((v + 15) & ~size_t(15)); }

__host__ __device__ inline SmemLayout rg_layout(size_t es, int LB, int S, int n, int n_rops, int n_gates,
                                                int n_consts, int n_theta, int n_blk, bool backward) {
  const int NA = 1 << LB, G = 1 << (n - LB), WV = G > 32 ? 2 : 1, PP = G > 32 ? 1 : 32 / G;
  const int NW = rg_warps(S, WV);
  const int NPT = S == 6 ? PP : (NW / WV) * PP;
  SmemLayout L{};
  int o = 0;
  L.rops = o; o = rg_align(o + sizeof(ROp) * (size_t)n_rops);
  L.cs = o; o = rg_align(o + es * 4 * (size_t)n_gates);
  L.u4 = o; o = rg_align(o + es * 32 * (size_t)n_consts);
  L.zj = o; o = rg_align(o + es * (size_t)NPT * n * S);
  L.qj = o; o = rg_align(o + es * (size_t)NPT * n * S);
  L.rj = o; o = rg_align(o + es * (size_t)NPT * n * 2 * S);
  L.tab = o; o = rg_align(o + es * (size_t)NPT * (NA + G) * S);
  L.exch = o; o = rg_align(o + es * 2 * (size_t)NW * 32 * NA);
  L.xred = o; o = rg_align(o + es * (size_t)NW * kMaxQubitsReg);
  L.atab = o; L.tabbar = o; L.rbar = o; L.gth = o; L.wacc = o;
  if (backward) {
    L.atab = o; o = rg_align(o + es * (size_t)NPT * S * NA);
    L.tabbar = o; o = rg_align(o + es * (size_t)NPT * (NA + G) * S);
    L.rbar = o; o = rg_align(o + es * (size_t)NPT * n * 2 * S);
  }
  L.total = o;
  return L;
}

// ---------------------------------------------------------------------------------------------
// register-level gate primitives.  PT = local target position (compile time), NA = 2^LB.
// ---------------------------------------------------------------------------------------------
// resident CTAs per SM the register allocation is bounded for (S = 6: 192 threads, S = 1: 128)
#ifndef RG_MINB_F6
#define RG_MINB_F6 3
#endif
#ifndef RG_MINB_B6
#define RG_MINB_B6 2
#endif
#ifndef RG_MINB_F1
#define RG_MINB_F1 4
#endif
#ifndef RG_MINB_B1
#define RG_MINB_B1 3
#endif
__host__ __device__ constexpr int rg_min_blocks(int S, bool backward) {
  return S == 6 ? (backward ? RG_MINB_B6 : RG_MINB_F6) : (backward ? RG_MINB_B1 : RG_MINB_F1);
}

#define RG_PAIR(h, PT) \
  const int i0 = (((h) >> (PT)) << ((PT) + 1)) | ((h) & ((1 << (PT)) - 1)), i1 = i0 | (1 << (PT))

// Controlled gates with a LANE control are a per-lane predicate around the plain variant; with a
// LOCAL control the host precomputes cmask (bit h = pair h is active) and the MASKED variant
// selects identity coefficients for the inactive pairs -- no per-(target, control) code variants.

// dense one-qubit gate on a local position.  m = (c, s, -, -) for T_X, (m00, m01, m10, m11) for T_R
// (already daggered by the caller when un-applying).
template <typename T, int LB, int PT, bool MASKED>
__device__ __forceinline__ void l1_apply(T (&ax)[1 << LB], T (&ay)[1 << LB], int type, T m0, T m1,
                                         T m2, T m3, unsigned cmask) {
  constexpr int NA = 1 << LB;
  if (type == T_X) {
#pragma unroll
    for (int h = 0; h < NA / 2; ++h) {
      RG_PAIR(h, PT);
      T c = m0, s = m1;
      if constexpr (MASKED) {
        const bool act = (cmask >> h) & 1;
        c = act ? m0 : T(1);
        s = act ? m1 : T(0);
      }
      const T x0 = ax[i0], y0 = ay[i0], x1 = ax[i1], y1 = ay[i1];
      ax[i0] = fma(c, x0, s * y1);
      ay[i0] = fma(c, y0, -s * x1);
      ax[i1] = fma(c, x1, s * y0);
      ay[i1] = fma(c, y1, -s * x0);
    }
  } else if (type == T_Z) {
#pragma unroll
    for (int h = 0; h < NA / 2; ++h) {
      RG_PAIR(h, PT);
      T c = m0, s = m1;
      if constexpr (MASKED) {
        const bool act = (cmask >> h) & 1;
        c = act ? m0 : T(1);
        s = act ? m1 : T(0);
      }
      const T x0 = ax[i0], y0 = ay[i0], x1 = ax[i1], y1 = ay[i1];
      ax[i0] = fma(c, x0, s * y0);
      ay[i0] = fma(c, y0, -s * x0);
      ax[i1] = fma(c, x1, -s * y1);
      ay[i1] = fma(c, y1, s * x1);
    }
  } else if constexpr (!MASKED) {   // real 2x2 gates are never controlled
#pragma unroll
    for (int h = 0; h < NA / 2; ++h) {
      RG_PAIR(h, PT);
      const T x0 = ax[i0], y0 = ay[i0], x1 = ax[i1], y1 = ay[i1];
      ax[i0] = fma(m0, x0, m1 * x1);
      ay[i0] = fma(m0, y0, m1 * y1);
      ax[i1] = fma(m2, x0, m3 * x1);
      ay[i1] = fma(m2, y0, m3 * y1);
    }
  }
}

// generator expectation Im<lambda|H|psi> over this lane's pairs (both vectors AFTER the gate):
// H = X for T_X (RX, CRX), H = Y for T_R (RY), H = Z for T_Z (RZ, CRZ)
template <typename T, int LB, int PT, bool MASKED>
__device__ __forceinline__ T l1_grad(const T (&ax)[1 << LB], const T (&ay)[1 << LB],
                                     const T (&lx)[1 << LB], const T (&ly)[1 << LB], int type,
                                     unsigned cmask) {
  constexpr int NA = 1 << LB;
  T p0 = T(0), p1 = T(0);
  if (type == T_X) {
#pragma unroll
    for (int h = 0; h < NA / 2; ++h) {
      RG_PAIR(h, PT);
      if constexpr (MASKED) {
        const bool act = (cmask >> h) & 1;
        const T t0 = fma(lx[i0], ay[i1], lx[i1] * ay[i0]);
        const T t1 = fma(ly[i0], ax[i1], ly[i1] * ax[i0]);
        p0 += act ? t0 : T(0);
        p1 += act ? t1 : T(0);
      } else {
        p0 = fma(lx[i0], ay[i1], p0);
        p1 = fma(ly[i0], ax[i1], p1);
        p0 = fma(lx[i1], ay[i0], p0);
        p1 = fma(ly[i1], ax[i0], p1);
      }
    }
  } else if (type == T_Z) {
#pragma unroll
    for (int h = 0; h < NA / 2; ++h) {
      RG_PAIR(h, PT);
      // Im(conj(l0) a0) - Im(conj(l1) a1)
      const T t0 = fma(lx[i0], ay[i0], ly[i1] * ax[i1]);
      const T t1 = fma(ly[i0], ax[i0], lx[i1] * ay[i1]);
      if constexpr (MASKED) {
        const bool act = (cmask >> h) & 1;
        p0 += act ? t0 : T(0);
        p1 += act ? t1 : T(0);
      } else {
        p0 += t0;
        p1 += t1;
      }
    }
  } else if constexpr (!MASKED) {
#pragma unroll
    for (int h = 0; h < NA / 2; ++h) {
      RG_PAIR(h, PT);
      p0 = fma(lx[i1], ax[i0], p0);
      p1 = fma(lx[i0], ax[i1], p1);
      p0 = fma(ly[i1], ay[i0], p0);
      p1 = fma(ly[i0], ay[i1], p1);
    }
  }
  return p0 - p1;
}

// CNOT with a local target: conditional exchange of the two amplitudes of every active pair
template <typename T, int LB, int PT>
__device__ __forceinline__ void cx_apply(T (&ax)[1 << LB], T (&ay)[1 << LB], unsigned cmask) {
  constexpr int NA = 1 << LB;
#pragma unroll
  for (int h = 0; h < NA / 2; ++h) {
    RG_PAIR(h, PT);
    const bool act = (cmask >> h) & 1;
    const T x0 = ax[i0], y0 = ay[i0], x1 = ax[i1], y1 = ay[i1];
    ax[i0] = act ? x1 : x0; ay[i0] = act ? y1 : y0;
    ax[i1] = act ? x0 : x1; ay[i1] = act ? y0 : y1;
  }
}

// exchange the roles of local position PT and the lane position whose xor mask is lmask
// (branch free: every lane sends one amplitude of each pair and overwrites the slot it sent)
template <typename T, int LB, int PT>
__device__ __forceinline__ void swap_ll(T (&ax)[1 << LB], T (&ay)[1 << LB], int lmask, bool mybit) {
  constexpr int NA = 1 << LB;
#pragma unroll
  for (int h = 0; h < NA / 2; ++h) {
    RG_PAIR(h, PT);
    const T rx = shx(mybit ? ax[i0] : ax[i1], lmask);
    const T ry = shx(mybit ? ay[i0] : ay[i1], lmask);
    ax[i0] = mybit ? rx : ax[i0];
    ax[i1] = mybit ? ax[i1] : rx;
    ay[i0] = mybit ? ry : ay[i0];
    ay[i1] = mybit ? ay[i1] : ry;
  }
}

// Arbitrary permutation of the (local | lane) index bits of a register tile through a warp-private
// shared-memory buffer of 32 * 2^LB complex values: ONE round trip instead of one shuffle swap per
// exchanged pair.  Element (i, lane) is written at slot (i << 5 | lane) ^ W(i) and read back from
// slot R(lane) ^ R(i): every destination bit contributes the slot mask of its source position, and
// a per-op bank swizzle (a local source bit that becomes lane bit y toggles COLUMN bit y) makes the
// lane -> column map of the permuted read a bijection, so the write and the read are both bank-
// conflict free.  All masks are precomputed by the host (LayoutTracker::emit_perm):
//   w[0..1]  write masks of the LB local bits (row bit | swizzle), 10 bits each, three per word
//   w[2..5]  read masks of the 10 destination positions, 10 bits each, three per word
struct PermMasks {
  uint32_t w[6];
};

__device__ __forceinline__ PermMasks perm_masks_of(const ROp& op) {
  return {{(uint32_t)op.pc, (uint32_t)op.type, (uint32_t)op.g, (uint32_t)op.p, (uint32_t)op.m, (uint32_t)op.pad}};
}

__device__ __forceinline__ unsigned perm_field(const PermMasks& pm, int base, int j) {
  return (pm.w[base + j / 3] >> (10 * (j % 3))) & 1023u;
}

// the WV warps of one stream vector meet at a named barrier (ids 1..15; id 0 is __syncthreads)
template <int WV>
__device__ __forceinline__ void vec_sync(int vec) {
  if constexpr (WV == 1) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(1 + vec), "n"(32 * WV) : "memory");
}

template <typename T, int LB, int WV = 1>
__device__ __forceinline__ void perm_apply(T (&ax)[1 << LB], T (&ay)[1 << LB], C2<T>* buf,
                                           const PermMasks& pm, int lane, int vec = 0) {
  constexpr int NA = 1 << LB;
  unsigned wl[LB], rl[LB];
#pragma unroll
  for (int x = 0; x < LB; ++x) {
    wl[x] = perm_field(pm, 0, x);
    rl[x] = perm_field(pm, 2, x);
  }
  unsigned rn = 0;       // `lane` = lane index inside the vector (5 bits, or 6 with WV = 2)
#pragma unroll
  for (int y = 0; y < (WV == 2 ? 6 : 5); ++y) rn ^= ((lane >> y) & 1) ? perm_field(pm, 2, LB + y) : 0u;
  vec_sync<WV>(vec);
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    unsigned sl = lane;
#pragma unroll
    for (int x = 0; x < LB; ++x)
      if ((i >> x) & 1) sl ^= wl[x];
    buf[sl] = {ax[i], ay[i]};
  }
  vec_sync<WV>(vec);
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    unsigned sl = rn;
#pragma unroll
    for (int x = 0; x < LB; ++x)
      if ((i >> x) & 1) sl ^= rl[x];
    const C2<T> v = buf[sl];
    ax[i] = v.x; ay[i] = v.y;
  }
}

template <typename T, int LB>
__device__ __forceinline__ void diag_apply(T (&ax)[1 << LB], T (&ay)[1 << LB], const C2A<T>* tab,
                                           int G, bool dag) {
  constexpr int NA = 1 << LB;
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const C2A<T> d = tab[i * G];
    const T dy = dag ? -d.y : d.y;
    const T x = ax[i], y = ay[i];
    ax[i] = fma(x, d.x, -y * dy);
    ay[i] = fma(x, dy, y * d.x);
  }
}

// fixed 4x4 unitary on local positions (1, 0): row index 2 * bit1 + bit0
template <typename T, int LB>
__device__ __forceinline__ void u4_apply(T (&ax)[1 << LB], T (&ay)[1 << LB], const C2<T>* U, bool dag) {
  constexpr int NA = 1 << LB;
#pragma unroll
  for (int gidx = 0; gidx < NA / 4; ++gidx) {
    T vx[4], vy[4], ox[4], oy[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { vx[j] = ax[4 * gidx + j]; vy[j] = ay[4 * gidx + j]; }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      T sx = T(0), sy = T(0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const C2<T> e = dag ? U[j * 4 + r] : U[r * 4 + j];
        const T ey = dag ? -e.y : e.y;
        sx = fma(e.x, vx[j], fma(-ey, vy[j], sx));
        sy = fma(e.x, vy[j], fma(ey, vx[j], sy));
      }
      ox[r] = sx; oy[r] = sy;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { ax[4 * gidx + j] = ox[j]; ay[4 * gidx + j] = oy[j]; }
  }
}

#define RG_PT_SWITCH(PTVAR, ...)                                           \
  switch (PTVAR) {                                                         \
    case 0: { constexpr int PT = 0; __VA_ARGS__; } break;                  \
    case 1: { constexpr int PT = 1; __VA_ARGS__; } break;                  \
    case 2: { constexpr int PT = 2; __VA_ARGS__; } break;                  \
    case 3: { constexpr int PT = 3; __VA_ARGS__; } break;                  \
    default: if constexpr (LB > 4) { constexpr int PT = 4; __VA_ARGS__; } break; \
  }

// ---------------------------------------------------------------------------------------------
// per-CTA context (shared-memory views)
// ---------------------------------------------------------------------------------------------
extern __shared__ __align__(16) unsigned char rg_smem[];

// Pointers are rebuilt from the constant-bank offsets at every use: the hot gate loop keeps 128+
// registers of amplitudes live, so nothing else should be pinned in registers across it.
template <typename T, int S>
struct Ctx {
  const RgArgs& a;
  template <typename U>
  __device__ __forceinline__ U* at(int off) const { return reinterpret_cast<U*>(rg_smem + off); }
  __device__ __forceinline__ const ROp* rops() const { return at<const ROp>(a.lay.rops); }
  __device__ __forceinline__ const T* cs() const { return at<const T>(a.lay.cs); }              // [n_gates][4]
  __device__ __forceinline__ const C2<T>* u4() const { return at<const C2<T>>(a.lay.u4); }      // [n_consts][16]
  __device__ __forceinline__ T* zj() const { return at<T>(a.lay.zj); }                          // [NPT][n*S]
  __device__ __forceinline__ T* qj() const { return at<T>(a.lay.qj); }                          // [NPT][n*S]
  __device__ __forceinline__ Jet<T, S>* rj() const { return at<Jet<T, S>>(a.lay.rj); }          // [NPT][n][2]
  __device__ __forceinline__ Jet<T, S>* tab() const { return at<Jet<T, S>>(a.lay.tab); }        // [NPT][NA+G]
  __device__ __forceinline__ C2<T>* exch() const { return at<C2<T>>(a.lay.exch); }              // [NW][32*NA]
  __device__ __forceinline__ T* xred() const { return at<T>(a.lay.xred); }                      // [NW][kMaxQubitsReg]
  __device__ __forceinline__ T* atab() const { return at<T>(a.lay.atab); }                      // [NPT][S][NA]
  __device__ __forceinline__ Jet<T, S>* tabbar() const { return at<Jet<T, S>>(a.lay.tabbar); }  // [NPT][NA+G]
  __device__ __forceinline__ Jet<T, S>* rbar() const { return at<Jet<T, S>>(a.lay.rbar); }      // [NPT][n][2]
};

// rops, gate coefficients and the Haar constants -> shared memory (once per CTA)
template <typename T, int S>
__device__ void load_program(const RgArgs& a) {
  ROp* rops = reinterpret_cast<ROp*>(rg_smem + a.lay.rops);
  T* cs = reinterpret_cast<T*>(rg_smem + a.lay.cs);
  C2<T>* u4 = reinterpret_cast<C2<T>*>(rg_smem + a.lay.u4);
  for (int r = threadIdx.x; r < a.n_rops; r += blockDim.x) rops[r] = a.rops[r];
  const T* theta = static_cast<const T*>(a.theta);
  for (int g = threadIdx.x; g < a.n_gates; g += blockDim.x) {
    const GateOp op = a.gates[g];
    double s = 0.0, c = 1.0;
    if (op.p >= 0 && op.kind != QCP_GATE_U4) sincos(0.5 * (double)theta[op.p], &s, &c);
    T m0 = T(0), m1 = T(0), m2 = T(0), m3 = T(0);
    switch (op.kind) {
      case QCP_GATE_RX: case QCP_GATE_CRX: case QCP_GATE_RZ: case QCP_GATE_CRZ: m0 = (T)c; m1 = (T)s; break;
      case QCP_GATE_RY: m0 = (T)c; m1 = (T)(-s); m2 = (T)s; m3 = (T)c; break;
      case QCP_GATE_H: {
        const T h = (T)0.70710678118654752440;
        m0 = h; m1 = h; m2 = h; m3 = -h;
        break;
      }
      default: break;
    }
    cs[4 * g] = m0; cs[4 * g + 1] = m1; cs[4 * g + 2] = m2; cs[4 * g + 3] = m3;
  }
  for (int e = threadIdx.x; e < 16 * a.n_consts; e += blockDim.x) {
    const double2 v = a.consts[e];
    u4[e] = {(T)v.x, (T)v.y};
  }
}

// ---------------------------------------------------------------------------------------------
// gate program, forward direction, on one register vector
// ---------------------------------------------------------------------------------------------
template <typename T, int LB, int S, int WV>
__device__ __forceinline__ void run_forward(T (&ax)[1 << LB], T (&ay)[1 << LB], const Ctx<T, S>& c,
                                            const RgArgs& a, int lig, int G) {
  const C2A<T>* diag = static_cast<const C2A<T>*>(a.diag);
  for (int r = 0; r < a.n_rops; ++r) {
    // LOCKSTEP: the unrolled register code has no temporal reuse, so instruction fetch is the
    // bottleneck; warps of a CTA that execute the same op at the same time share the fetched lines
    // (measured +11 % on the forward, +2 % on the backward; bigger lockstep CTAs lose to barrier
    // stalls)
    __syncthreads();
    const ROp op = c.rops()[r];
    switch (op.kind) {
      case R_L1: {
        const T m0 = c.cs()[4 * op.g], m1 = c.cs()[4 * op.g + 1], m2 = c.cs()[4 * op.g + 2], m3 = c.cs()[4 * op.g + 3];
        if (op.pc >= 0 && op.pc < LB) {
          RG_PT_SWITCH(op.pt, l1_apply<T, LB, PT, true>(ax, ay, op.type, m0, m1, m2, m3, (unsigned)op.m))
        } else if (op.pc < 0 || ((lig >> (op.pc - LB)) & 1)) {
          RG_PT_SWITCH(op.pt, l1_apply<T, LB, PT, false>(ax, ay, op.type, m0, m1, m2, m3, 0u))
        }
        break;
      }
      case R_CX: {
        const unsigned cm = op.pc < LB ? (unsigned)op.m : (((lig >> (op.pc - LB)) & 1) ? 0xffffu : 0u);
        RG_PT_SWITCH(op.pt, cx_apply<T, LB, PT>(ax, ay, cm))
        break;
      }
      case R_SWAP: {
        const int sh = op.pc - LB;
        const bool mybit = (lig >> sh) & 1;
        RG_PT_SWITCH(op.pt, swap_ll<T, LB, PT>(ax, ay, 1 << sh, mybit))
        break;
      }
      case R_DIAG:
        diag_apply<T, LB>(ax, ay, diag + ((size_t)op.g << a.n) + lig, G, false);
        break;
      case R_PERM:
        if constexpr (WV == 1)
          perm_apply<T, LB>(ax, ay, c.exch() + (size_t)(threadIdx.x >> 5) * 32 * (1 << LB),
                            perm_masks_of(op), threadIdx.x & 31);
        else       // the two rows of the warp pair are one 64 x NA buffer
          perm_apply<T, LB, 2>(ax, ay, c.exch() + (size_t)(threadIdx.x >> 6) * 64 * (1 << LB),
                               perm_masks_of(op), lig, threadIdx.x >> 6);
        break;
      case R_PERMB:
        break;
      default:
        u4_apply<T, LB>(ax, ay, c.u4() + 16 * op.g, false);
        break;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// encoding: z jets -> one-qubit jets -> lane/local product tables -> psi0 stream in registers
// ---------------------------------------------------------------------------------------------
// step 1 (threads over (slot, j)): r_j = (cos z_j/2, sin z_j/2) as jets        [angle]
//                                  e_j = f_j / |f|                              [amplitude, thread per slot]
template <typename T, int S>
__device__ void encode_qubit_jets(const Ctx<T, S>& c, int n, int enc, int NPT) {
  if (enc == QCP_ENC_ANGLE) {
    for (int e = threadIdx.x; e < NPT * n; e += blockDim.x) {
      const int slot = e / n, j = e % n;
      Jet<T, S> z;
#pragma unroll
      for (int k = 0; k < S; ++k) z.c[k] = c.zj()[(size_t)slot * n * S + j * S + k];
      T sn, cs;
      Math<T>::sincos_(T(0.5) * z.c[0], &sn, &cs);
      Jet<T, S> h = z;     // jet of z/2
#pragma unroll
      for (int k = 0; k < S; ++k) h.c[k] = T(0.5) * z.c[k];
      c.rj()[((size_t)slot * n + j) * 2] = jfunc(h, cs, -sn, -cs);
      c.rj()[((size_t)slot * n + j) * 2 + 1] = jfunc(h, sn, cs, -sn);
    }
  } else {
    for (int slot = threadIdx.x; slot < NPT; slot += blockDim.x) {
      Jet<T, S> nrm;
      jzero(nrm);
      for (int j = 0; j < n; ++j) {
        Jet<T, S> f;
#pragma unroll
        for (int k = 0; k < S; ++k) f.c[k] = c.zj()[(size_t)slot * n * S + j * S + k];
        jmul_acc(nrm, f, f);
      }
      // padded lanes carry z = 0: keep them finite (their outputs are never stored)
      const T n0 = nrm.c[0] > T(0) ? nrm.c[0] : T(1);
      const T r = T(1) / sqrt(n0);
      const T g1 = T(-0.5) * r / n0, g2 = T(0.75) * r / (n0 * n0);
      const Jet<T, S> inv = jfunc(nrm, r, g1, g2);
      for (int j = 0; j < n; ++j) {
        Jet<T, S> f;
#pragma unroll
        for (int k = 0; k < S; ++k) f.c[k] = c.zj()[(size_t)slot * n * S + j * S + k];
        c.rj()[((size_t)slot * n + j) * 2] = jmul(f, inv);
      }
    }
  }
}

// step 2 (angle, threads over (slot, entry)): tab[slot][i] = prod_x r_{n-1-x}[bit_x(i)] (local part),
// tab[slot][NA + l] = prod_y r_{n-1-LB-y}[bit_y(l)] (lane part)
template <typename T, int LB, int S>
__device__ void encode_tables(const Ctx<T, S>& c, int n, int NPT) {
  constexpr int NA = 1 << LB;
  const int G = 1 << (n - LB), NE = NA + G;
  for (int e = threadIdx.x; e < NPT * NE; e += blockDim.x) {
    const int slot = e / NE, ent = e % NE;
    const bool lane_part = ent >= NA;
    const int idx = lane_part ? ent - NA : ent;
    const int K = lane_part ? n - LB : LB;
    const int q0 = lane_part ? n - 1 - LB : n - 1;
    Jet<T, S> acc = c.rj()[((size_t)slot * n + q0) * 2 + (idx & 1)];
    for (int k = 1; k < K; ++k)
      acc = jmul(acc, c.rj()[((size_t)slot * n + (q0 - k)) * 2 + ((idx >> k) & 1)]);
    c.tab()[(size_t)slot * NE + ent] = acc;
  }
}

// (vx, vy) * (-i)^m4; m4 is a constant after unrolling, so the switch folds away
template <typename T>
__device__ __forceinline__ void phase_rot(int m4, T vx, T vy, T& ox, T& oy) {
  switch (m4 & 3) {
    case 0: ox = vx; oy = vy; break;
    case 1: ox = vy; oy = -vx; break;
    case 2: ox = -vx; oy = -vy; break;
    default: ox = -vy; oy = vx; break;
  }
}

// popcount of a 5-bit constant, written so that it folds after unrolling
#define RG_POPC5(v) (((v) & 1) + (((v) >> 1) & 1) + (((v) >> 2) & 1) + (((v) >> 3) & 1) + (((v) >> 4) & 1))

template <typename T, int LB, int S>
__device__ __forceinline__ void encode_stream(T (&ax)[1 << LB], T (&ay)[1 << LB], const Ctx<T, S>& c,
                                              int n, int enc, int slot, int lig, int s) {
  constexpr int NA = 1 << LB;
  const int G = 1 << (n - LB), NE = NA + G;
  if (enc == QCP_ENC_ANGLE) {
    const Jet<T, S>* tab = c.tab() + (size_t)slot * NE;
    const Jet<T, S>& Lj = tab[NA + lig];
    const T L0 = Lj.c[0], Ls = Lj.c[s], Lp2 = (S == 6 && s >= 4) ? T(2) * Lj.c[s - 2] : T(0);
    const int pl = __popc(lig) & 3;
    const T bx = pl == 0 ? T(1) : (pl == 2 ? T(-1) : T(0));
    const T by = pl == 1 ? T(-1) : (pl == 3 ? T(1) : T(0));
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const Jet<T, S>& R = tab[i];
      T v = Ls * R.c[0];
      if constexpr (S == 6) {
        if (s > 0) v = fma(L0, R.c[s], v);
        if (s >= 4) v = fma(Lp2, R.c[s - 2], v);
      }
      phase_rot<T>(RG_POPC5(i), v * bx, v * by, ax[i], ay[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int k = (lig << LB) | i;
      ax[i] = k < n ? c.rj()[((size_t)slot * n + k) * 2].c[s] : T(0);
      ay[i] = T(0);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// measurement helpers
// ---------------------------------------------------------------------------------------------
// signed sums of a per-amplitude real density over the group -> q[qubit] (valid in every lane)
template <typename T, int LB, int WV = 1>
__device__ __forceinline__ void signed_sums(const T (&w)[1 << LB], const RgArgs& a, int lig, int G,
                                            T* out /* [n] registers via static loop bound kMaxQubitsReg */,
                                            T* xred = nullptr /* [NW][kMaxQubitsReg] when WV == 2 */) {
  constexpr int NA = 1 << LB;
  T t = T(0), sx[LB];
#pragma unroll
  for (int x = 0; x < LB; ++x) sx[x] = T(0);
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    t += w[i];
#pragma unroll
    for (int x = 0; x < LB; ++x) sx[x] += ((i >> x) & 1) ? -w[i] : w[i];
  }
#pragma unroll
  for (int q = 0; q < kMaxQubitsReg; ++q) {
    if (q < a.n) {
      const int pos = a.meas_pos[q];
      T v;
      if (pos >= LB) {
        v = ((lig >> (pos - LB)) & 1) ? -t : t;
      } else {
        v = sx[0];
#pragma unroll
        for (int x = 1; x < LB; ++x) v = pos == x ? sx[x] : v;
      }
      for (int m = 1; m < (G < 32 ? G : 32); m <<= 1) v += shx(v, m);
      out[q] = v;
    }
  }
  if constexpr (WV == 2) {       // the other half of the vector lives in the partner warp
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int q = 0; q < kMaxQubitsReg; ++q)
        if (q < a.n) xred[warp * kMaxQubitsReg + q] = out[q];
    }
    vec_sync<2>(warp >> 1);
#pragma unroll
    for (int q = 0; q < kMaxQubitsReg; ++q)
      if (q < a.n) out[q] += xred[(warp ^ 1) * kMaxQubitsReg + q];
    vec_sync<2>(warp >> 1);
  }
}

// ---------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------
template <typename T, int LB, int S, int WV>
__global__ void __launch_bounds__(rg_warps(S, WV) * 32, WV == 2 ? 1 : rg_min_blocks(S, false))
rg_forward_kernel(const __grid_constant__ RgArgs a) {
  constexpr int NA = 1 << LB, NW = rg_warps(S, WV);
  const int n = a.n, G = 1 << (n - LB), PP = WV == 2 ? 1 : 32 / G, NPT = S == 6 ? PP : (NW / WV) * PP;
  const Ctx<T, S> c{a};
  load_program<T, S>(a);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wv = WV == 2 ? warp & 1 : 0, vec = warp / WV;          // half of / index of this warp's vector
  const int lig = WV == 2 ? (wv << 5) | lane : lane & (G - 1), sub = WV == 2 ? 0 : lane / G;
  const int slot = S == 6 ? sub : vec * PP + sub;
  const int s = S == 6 ? vec : 0;
  const T* ws = static_cast<const T*>(a.ws);
  T* ws_out = static_cast<T*>(a.ws);
  const int nS = n * S;
  __syncthreads();

  for (long long base = (long long)blockIdx.x * NPT; base < a.B; base += (long long)gridDim.x * NPT) {
    for (int e = threadIdx.x; e < NPT * nS; e += blockDim.x) {
      const int sl = e % NPT, rem = e / NPT;
      const long long p = base + sl;
      c.zj()[(size_t)sl * nS + rem] = p < a.B ? ws[(size_t)rem * a.B + p] : T(0);
    }
    __syncthreads();
    encode_qubit_jets<T, S>(c, n, a.enc, NPT);
    __syncthreads();
    if (a.enc == QCP_ENC_ANGLE) encode_tables<T, LB, S>(c, n, NPT);
    __syncthreads();

    T ax[NA], ay[NA];
    encode_stream<T, LB, S>(ax, ay, c, n, a.enc, slot, lig, s);
    run_forward<T, LB, S, WV>(ax, ay, c, a, lig, G);

    const long long p = base + slot;
    const bool valid = p < a.B;
    if (a.state && valid) {
      C2A<T>* st = static_cast<C2A<T>*>(a.state) + (((size_t)p * S + s) << n) + lig;
#pragma unroll
      for (int i = 0; i < NA; ++i) st[i * G] = {ax[i], ay[i]};
    }
    C2<T>* ex = c.exch();          // row of (stream k, half wv) = k * WV + wv
    if constexpr (S == 6) {
      if constexpr (WV == 2) __syncthreads();   // the rows were PERM staging buffers until now
      if (s < 4) {
#pragma unroll
        for (int i = 0; i < NA; ++i) ex[(size_t)(s * WV + wv) * 32 * NA + i * 32 + lane] = {ax[i], ay[i]};
      }
      __syncthreads();
    }
    T w[NA];
    if (s == 0) {
#pragma unroll
      for (int i = 0; i < NA; ++i) w[i] = fma(ax[i], ax[i], ay[i] * ay[i]);
    } else {
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        const C2<T> p0 = ex[(size_t)wv * 32 * NA + i * 32 + lane];
        T v = T(2) * fma(ax[i], p0.x, ay[i] * p0.y);
        if (s >= 4) {
          const C2<T> pd = ex[(size_t)((s - 2) * WV + wv) * 32 * NA + i * 32 + lane];
          v = fma(T(2), fma(pd.x, pd.x, pd.y * pd.y), v);
        }
        w[i] = v;
      }
    }
    T q[kMaxQubitsReg];
    signed_sums<T, LB, WV>(w, a, lig, G, q, c.xred());
    if (lig == 0 && valid) {
#pragma unroll
      for (int j = 0; j < kMaxQubitsReg; ++j)
        if (j < n) ws_out[(size_t)(nS + j * S + s) * a.B + p] = q[j];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// backward kernel
// ---------------------------------------------------------------------------------------------
template <typename T, int LB, int S, int WV>
__global__ void __launch_bounds__(rg_warps(S, WV) * 32, WV == 2 ? 1 : rg_min_blocks(S, true))
rg_backward_kernel(const __grid_constant__ RgArgs a) {
  constexpr int NA = 1 << LB, NW = rg_warps(S, WV);
  const int n = a.n, G = 1 << (n - LB), PP = WV == 2 ? 1 : 32 / G, NPT = S == 6 ? PP : (NW / WV) * PP, NE = NA + G;
  const Ctx<T, S> c{a};
  load_program<T, S>(a);
  // per-WARP accumulator rows in global memory (zeroed by the host).  Every address has exactly one
  // writing THREAD, whose fire-and-forget RED.ADDs to it are applied in program order, so the sums
  // are bit-reproducible (round 1 shared one row per CTA between its warps: the order of their
  // atomics varied); the rows are summed in row order by the reduction kernels.  (A plain
  // load-add-store instead of the RED costs 26 % of the step: its latency is exposed.)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t wrow = (size_t)blockIdx.x * NW + warp;
  double* const gth = a.theta_partials + wrow * (a.n_theta > 0 ? a.n_theta : 1);
  T* const wacc = static_cast<T*>(a.w_partials) + ((size_t)blockIdx.x * a.n_blk << n);
  const int wv = WV == 2 ? warp & 1 : 0, vec = warp / WV;          // half of / index of this warp's vector
  const int lig = WV == 2 ? (wv << 5) | lane : lane & (G - 1), sub = WV == 2 ? 0 : lane / G;
  const int slot = S == 6 ? sub : vec * PP + sub;
  const int s = S == 6 ? vec : 0;
  const int row = warp;                            // exchange-buffer row of this warp (= s * WV + wv)
  const T* ws = static_cast<const T*>(a.ws);
  T* ws_out = static_cast<T*>(a.ws);
  const C2A<T>* diag = static_cast<const C2A<T>*>(a.diag);
  const int nS = n * S;
  __syncthreads();

  for (long long base = (long long)blockIdx.x * NPT; base < a.B; base += (long long)gridDim.x * NPT) {
    // ---- B1: z jets and q cotangents ----------------------------------------------------------
    for (int e = threadIdx.x; e < NPT * nS; e += blockDim.x) {
      const int sl = e % NPT, rem = e / NPT;
      const long long p = base + sl;
      const bool ok = p < a.B;
      c.zj()[(size_t)sl * nS + rem] = ok ? ws[(size_t)rem * a.B + p] : T(0);
      c.qj()[(size_t)sl * nS + rem] = ok ? ws[(size_t)(nS + rem) * a.B + p] : T(0);
    }
    __syncthreads();
    // ---- B2/B3: encoding tables, local part of the sign sums -----------------------------------
    encode_qubit_jets<T, S>(c, n, a.enc, NPT);
    for (int e = threadIdx.x; e < NPT * S * NA; e += blockDim.x) {
      const int sl = e / (S * NA), k = (e / NA) % S, i = e % NA;
      T acc = T(0);
      for (int q = 0; q < n; ++q) {
        const int pos = a.meas_pos[q];
        if (pos < LB) {
          const T v = c.qj()[(size_t)sl * nS + q * S + k];
          acc += ((i >> pos) & 1) ? -v : v;
        }
      }
      c.atab()[e] = acc;
    }
    __syncthreads();
    if (a.enc == QCP_ENC_ANGLE) encode_tables<T, LB, S>(c, n, NPT);
    __syncthreads();

    // ---- B4: final psi of this warp's stream ------------------------------------------------------
    const long long p = base + slot;
    const bool valid = p < a.B;
    T ax[NA], ay[NA];
    if (a.state) {
      const C2A<T>* st = static_cast<const C2A<T>*>(a.state) + (((size_t)(valid ? p : 0) * S + s) << n) + lig;
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        const C2A<T> v = st[i * G];
        ax[i] = valid ? v.x : T(0);
        ay[i] = valid ? v.y : T(0);
      }
    } else {
      encode_stream<T, LB, S>(ax, ay, c, n, a.enc, slot, lig, s);
      run_forward<T, LB, S, WV>(ax, ay, c, a, lig, G);
    }
    C2<T>* ex = c.exch();          // row of (stream k, half wv) = k * WV + wv
    if constexpr (S == 6) {
      if constexpr (WV == 2) __syncthreads();   // (recompute path) rows were PERM staging buffers
#pragma unroll
      for (int i = 0; i < NA; ++i) ex[(size_t)row * 32 * NA + i * 32 + lane] = {ax[i], ay[i]};
      __syncthreads();
    }
    // ---- B5: lambda streams from the q cotangents -----------------------------------------------
    T lx[NA], ly[NA];
    {
      T bl[S];   // lane part of the sign sums
#pragma unroll
      for (int k = 0; k < S; ++k) bl[k] = T(0);
      for (int q = 0; q < n; ++q) {
        const int pos = a.meas_pos[q];
        if (pos >= LB) {
          const bool neg = (lig >> (pos - LB)) & 1;
#pragma unroll
          for (int k = 0; k < S; ++k) {
            const T v = c.qj()[(size_t)slot * nS + q * S + k];
            bl[k] += neg ? -v : v;
          }
        }
      }
      const T* at = c.atab() + (size_t)slot * S * NA;
      if (s == 0) {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          const T z0 = T(2) * (at[i] + bl[0]);
          T vx = z0 * ax[i], vy = z0 * ay[i];
          if constexpr (S == 6) {
#pragma unroll
            for (int k = 1; k < 6; ++k) {
              const T zk = T(2) * (at[k * NA + i] + bl[k]);
              const C2<T> pk = ex[(size_t)(k * WV + wv) * 32 * NA + i * 32 + lane];
              vx = fma(zk, pk.x, vx);
              vy = fma(zk, pk.y, vy);
            }
          }
          lx[i] = vx; ly[i] = vy;
        }
      } else if constexpr (S == 6) {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          const C2<T> p0 = ex[(size_t)wv * 32 * NA + i * 32 + lane];
          const T zs = T(2) * (at[s * NA + i] + bl[s]);
          T vx = zs * p0.x, vy = zs * p0.y;
          if (s == 2 || s == 3) {
            const T ze = T(4) * (at[(s + 2) * NA + i] + bl[s + 2]);
            vx = fma(ze, ax[i], vx);
            vy = fma(ze, ay[i], vy);
          }
          lx[i] = vx; ly[i] = vy;
        }
      }
    }
    // ---- B6: gate program in reverse --------------------------------------------------------------
    __syncthreads();   // the exchange rows become the warps' private PERM buffers
    for (int r = a.n_rops - 1; r >= 0; --r) {
      __syncthreads();             // lockstep, see run_forward
      const ROp op = c.rops()[r];
      switch (op.kind) {
        case R_L1: {
          const T m0 = c.cs()[4 * op.g], m1 = c.cs()[4 * op.g + 1], m2 = c.cs()[4 * op.g + 2], m3 = c.cs()[4 * op.g + 3];
          // dagger: T_X -> s = -s ; T_R -> transpose
          const T d1 = op.type == T_R ? m2 : -m1, d2 = op.type == T_R ? m1 : m2;
          T part = T(0);
          if (op.pc >= 0 && op.pc < LB) {
            RG_PT_SWITCH(op.pt, {
              part = l1_grad<T, LB, PT, true>(ax, ay, lx, ly, op.type, (unsigned)op.m);
              l1_apply<T, LB, PT, true>(ax, ay, op.type, m0, d1, d2, m3, (unsigned)op.m);
              l1_apply<T, LB, PT, true>(lx, ly, op.type, m0, d1, d2, m3, (unsigned)op.m);
            })
          } else if (op.pc < 0 || ((lig >> (op.pc - LB)) & 1)) {
            RG_PT_SWITCH(op.pt, {
              if (op.p >= 0) part = l1_grad<T, LB, PT, false>(ax, ay, lx, ly, op.type, 0u);
              l1_apply<T, LB, PT, false>(ax, ay, op.type, m0, d1, d2, m3, 0u);
              l1_apply<T, LB, PT, false>(lx, ly, op.type, m0, d1, d2, m3, 0u);
            })
          }
          if (op.p >= 0) {
            for (int m = 16; m > 0; m >>= 1) part += shx(part, m);
            if (lane == 0) atomicAdd(gth + op.p, 0.5 * (double)part);
          }
          break;
        }
        case R_CX: {
          const unsigned cm = op.pc < LB ? (unsigned)op.m : (((lig >> (op.pc - LB)) & 1) ? 0xffffu : 0u);
          RG_PT_SWITCH(op.pt, {
            cx_apply<T, LB, PT>(ax, ay, cm);
            cx_apply<T, LB, PT>(lx, ly, cm);
          })
          break;
        }
        case R_SWAP: {
          const int sh = op.pc - LB;
          const bool mybit = (lig >> sh) & 1;
          RG_PT_SWITCH(op.pt, {
            swap_ll<T, LB, PT>(ax, ay, 1 << sh, mybit);
            swap_ll<T, LB, PT>(lx, ly, 1 << sh, mybit);
          })
          break;
        }
        case R_DIAG: {
          // W[k] += Im(conj(lambda_k) psi_k), summed over the CTA's vectors in a FIXED order through
          // the exchange buffer (the ops run in lockstep, nobody else uses it now), then one RED per
          // table entry from the thread that always owns that entry: bit-reproducible, and 1/NV of
          // the atomics of a per-vector accumulation
          constexpr int NV = NW / WV;
          const int M = 1 << n;
          T* wb = reinterpret_cast<T*>(c.exch());                 // [NV][M] reals
#pragma unroll
          for (int i = 0; i < NA; ++i) {
            T w = fma(lx[i], ay[i], -ly[i] * ax[i]);
            // a warp carries 32 / G points when a vector needs fewer than 32 lanes: fold them
            for (int m = G; m < 32; m <<= 1) w += shx(w, m);
            if (sub == 0) wb[(size_t)vec * M + i * G + lig] = w;
          }
          __syncthreads();
          T* wa = wacc + ((size_t)op.g << n);
          for (int e = threadIdx.x; e < M; e += blockDim.x) {
            T sum = wb[e];
#pragma unroll
            for (int v = 1; v < NV; ++v) sum += wb[(size_t)v * M + e];
            atomicAdd(wa + e, sum);
          }
          __syncthreads();
          const C2A<T>* tb = diag + ((size_t)op.g << n) + lig;
          diag_apply<T, LB>(ax, ay, tb, G, true);
          diag_apply<T, LB>(lx, ly, tb, G, true);
          break;
        }
        case R_PERMB: {
          const PermMasks pm = perm_masks_of(op);
          if constexpr (WV == 1) {
            C2<T>* pb = c.exch() + (size_t)warp * 32 * NA;
            perm_apply<T, LB>(ax, ay, pb, pm, lane);
            perm_apply<T, LB>(lx, ly, pb, pm, lane);
          } else {
            C2<T>* pb = c.exch() + (size_t)vec * 64 * NA;
            perm_apply<T, LB, 2>(ax, ay, pb, pm, lig, vec);
            perm_apply<T, LB, 2>(lx, ly, pb, pm, lig, vec);
          }
          break;
        }
        case R_PERM:
          break;
        default:
          u4_apply<T, LB>(ax, ay, c.u4() + 16 * op.g, true);
          u4_apply<T, LB>(lx, ly, c.u4() + 16 * op.g, true);
          break;
      }
    }
    // ---- B7: pull lambda0 back through the encoding ---------------------------------------------
    __syncthreads();   // every warp is done reading the psi exchange rows
    T* rb = reinterpret_cast<T*>(c.exch());       // [NW][32 * NA] real cotangents
    if (a.enc == QCP_ENC_ANGLE) {
      const int pl = __popc(lig) & 3;
      const T bx = pl == 0 ? T(1) : (pl == 2 ? T(-1) : T(0));
      const T by = pl == 1 ? T(-1) : (pl == 3 ? T(1) : T(0));
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        T fx, fy;
        phase_rot<T>(RG_POPC5(i), bx, by, fx, fy);
        rb[(size_t)row * 32 * NA + i * 32 + lane] = fma(lx[i], fx, ly[i] * fy);
      }
      __syncthreads();
      // cotangent jets of the table entries.  Items (slot, entry, stream k) sit on 8 consecutive
      // lanes: stream k owns component k of the entry, component 0 is the sum over the streams and
      // the second-order streams (k = 4, 5) also feed components 2, 3 -- combined with shuffles, so
      // every entry is written once with plain stores (no shared-memory atomics).
      {
        const int total = NPT * NE * 8;
        for (int e0 = 0; e0 < total; e0 += blockDim.x) {
          const int e = e0 + threadIdx.x;
          const bool live = e < total;
          const int k = e & 7, ent = live ? (e >> 3) % NE : 0, sl = live ? (e >> 3) / NE : 0;
          T a0 = T(0), as = T(0), ap = T(0);
          if (live && k < S) {
            // row of (stream / vector, half) and the lane inside it that holds rho(i, l)
            const int vrow = (S == 6 ? k : sl / PP) * WV, sb = S == 6 ? sl : sl % PP;
            auto rho_at = [&](int i, int l) -> T {
              if constexpr (WV == 2) return rb[(size_t)(vrow + (l >> 5)) * 32 * NA + i * 32 + (l & 31)];
              else return rb[(size_t)vrow * 32 * NA + i * 32 + sb * G + l];
            };
            const Jet<T, S>* tab = c.tab() + (size_t)sl * NE;
            if (ent >= NA) {          // lane entry: sum over the local index
              const int l = ent - NA;
              for (int i = 0; i < NA; ++i) {
                const T rho = rho_at(i, l);
                const Jet<T, S>& R = tab[i];
                a0 = fma(rho, R.c[k], a0);
                if (k > 0) as = fma(rho, R.c[0], as);
                if (k >= 4) ap = fma(T(2) * rho, R.c[k - 2], ap);
              }
            } else {                  // local entry: sum over the lanes of the group
              for (int l = 0; l < G; ++l) {
                const T rho = rho_at(ent, l);
                const Jet<T, S>& Lj = tab[NA + l];
                a0 = fma(rho, Lj.c[k], a0);
                if (k > 0) as = fma(rho, Lj.c[0], as);
                if (k >= 4) ap = fma(T(2) * rho, Lj.c[k - 2], ap);
              }
            }
          }
          a0 += shx(a0, 1); a0 += shx(a0, 2); a0 += shx(a0, 4);
          const T ap2 = __shfl_down_sync(0xffffffffu, ap, 2);      // stream k + 2 -> component k
          if (live && k < S) {
            Jet<T, S>& tb = c.tabbar()[(size_t)sl * NE + ent];
            if (k == 0) tb.c[0] = a0;
            else tb.c[k] = as + ((S == 6 && (k == 2 || k == 3)) ? ap2 : T(0));
          }
        }
      }
      __syncthreads();
      // table entries -> one-qubit jets: item (slot, qubit, bit, quarter) sums the leave-one-out
      // pullbacks of a quarter of the entries whose index has that bit; four lanes, two shuffles
      {
        const int total = NPT * n * 8;
        for (int e0 = 0; e0 < total; e0 += blockDim.x) {
          const int e = e0 + threadIdx.x;
          const bool live = e < total;
          const int part = e & 3, bit = (e >> 2) & 1, qq = live ? (e >> 3) % n : 0, sl = live ? (e >> 3) / n : 0;
          const bool lane_part = qq >= LB;
          const int kq = lane_part ? qq - LB : qq;
          const int K = lane_part ? n - LB : LB;
          const int q0 = lane_part ? n - 1 - LB : n - 1;
          const int half = (lane_part ? G : NA) >> 1;
          const int ebase = lane_part ? NA : 0;
          Jet<T, S> acc;
          jzero(acc);
          if (live) {
            for (int j = part; j < half; j += 4) {
              const int idx = (((j >> kq) << (kq + 1)) | (j & ((1 << kq) - 1))) | (bit << kq);
              Jet<T, S> other;
              jzero(other);
              other.c[0] = T(1);
              for (int k2 = 0; k2 < K; ++k2)
                if (k2 != kq) other = jmul(other, c.rj()[((size_t)sl * n + (q0 - k2)) * 2 + ((idx >> k2) & 1)]);
              jmul_pull_acc(acc, c.tabbar()[(size_t)sl * NE + ebase + idx], other);
            }
          }
#pragma unroll
          for (int m = 0; m < S; ++m) {
            acc.c[m] += shx(acc.c[m], 1);
            acc.c[m] += shx(acc.c[m], 2);
          }
          if (live && part == 0) c.rbar()[((size_t)sl * n + (n - 1 - qq)) * 2 + bit] = acc;
        }
      }
      __syncthreads();
      for (int e = threadIdx.x; e < NPT * n; e += blockDim.x) {
        const int sl = e / n, j = e % n;
        const long long pp = base + sl;
        if (pp >= a.B) continue;
        Jet<T, S> h;
#pragma unroll
        for (int k = 0; k < S; ++k) h.c[k] = T(0.5) * c.zj()[(size_t)sl * nS + j * S + k];
        T sn, cs;
        Math<T>::sincos_(h.c[0], &sn, &cs);
        Jet<T, S> hb;
        jzero(hb);
        jfunc_pull_acc(hb, c.rbar()[((size_t)sl * n + j) * 2], h, -sn, -cs, sn);
        jfunc_pull_acc(hb, c.rbar()[((size_t)sl * n + j) * 2 + 1], h, cs, -sn, -cs);
#pragma unroll
        for (int k = 0; k < S; ++k) ws_out[(size_t)(j * S + k) * a.B + pp] = T(0.5) * hb.c[k];
      }
    } else {
      // amplitude encoding: psi0_k = e_k (real) for k < n
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        const int k = (lig << LB) | i;
        if (k < n) c.rbar()[((size_t)slot * n + k) * 2].c[s] = lx[i];
      }
      __syncthreads();
      for (int sl = threadIdx.x; sl < NPT; sl += blockDim.x) {
        const long long pp = base + sl;
        if (pp >= a.B) continue;
        Jet<T, S> nrm, invb, nb;
        jzero(nrm); jzero(invb); jzero(nb);
        for (int j = 0; j < n; ++j) {
          Jet<T, S> f;
#pragma unroll
          for (int k = 0; k < S; ++k) f.c[k] = c.zj()[(size_t)sl * nS + j * S + k];
          jmul_acc(nrm, f, f);
        }
        const T r = T(1) / sqrt(nrm.c[0]);
        const T g1 = T(-0.5) * r / nrm.c[0], g2 = T(0.75) * r / (nrm.c[0] * nrm.c[0]);
        const T g3 = T(-1.875) * r / (nrm.c[0] * nrm.c[0] * nrm.c[0]);
        const Jet<T, S> inv = jfunc(nrm, r, g1, g2);
        for (int j = 0; j < n; ++j) {
          Jet<T, S> f;
#pragma unroll
          for (int k = 0; k < S; ++k) f.c[k] = c.zj()[(size_t)sl * nS + j * S + k];
          jmul_pull_acc(invb, c.rbar()[((size_t)sl * n + j) * 2], f);
        }
        jfunc_pull_acc(nb, invb, nrm, g1, g2, g3);
        for (int j = 0; j < n; ++j) {
          Jet<T, S> f, fb;
#pragma unroll
          for (int k = 0; k < S; ++k) f.c[k] = c.zj()[(size_t)sl * nS + j * S + k];
          jzero(fb);
          jmul_pull_acc(fb, c.rbar()[((size_t)sl * n + j) * 2], inv);
          jmul_pull_acc(fb, nb, f);
          jmul_pull_acc(fb, nb, f);
#pragma unroll
          for (int k = 0; k < S; ++k) ws_out[(size_t)(j * S + k) * a.B + pp] = fb.c[k];
        }
      }
    }
    __syncthreads();
  }
}

// per-dtype launchers (qcp_reg_f32.cu / qcp_reg_f64.cu)
template <typename T>
int rg_launch(int LB, int WV, int S, bool backward, const RgArgs& a, int grid, size_t smem, cudaStream_t s);
template <typename T>
int rg_occupancy(int LB, int WV, int S, bool backward, size_t smem, int* blocks_per_sm);

template <typename K>
inline int rg_launch_one(K kernel, const RgArgs& a, int grid, int threads, size_t smem, cudaStream_t s,
                         const char* what) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("%s: cannot opt in to %zu bytes of shared memory: %s", what, smem, cudaGetErrorString(e));
      return 1;
    }
  }
  kernel<<<grid, threads, smem, s>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("%s: launch failed: %s", what, cudaGetErrorString(e)); return 1; }
  return 0;
}

template <typename K>
inline int rg_occ_one(K kernel, int threads, size_t smem, int* out) {
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    *out = 0;
    return 1;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, kernel, threads, smem) != cudaSuccess) {
    cudaGetLastError();
    *out = 0;
    return 1;
  }
  return 0;
}

// dispatch over (LB, S, direction); RG_CALL(kernel-template, T, LB, S) is defined by the includer
#define RG_INSTANTIATE(T, LBV, WVV)                                                         \
  if (LB == LBV && WV == WVV) {                                                             \
    if (S == 6)                                                                             \
      return backward ? RG_CALL(rg_backward_kernel, T, LBV, 6, WVV) : RG_CALL(rg_forward_kernel, T, LBV, 6, WVV); \
    return backward ? RG_CALL(rg_backward_kernel, T, LBV, 1, WVV) : RG_CALL(rg_forward_kernel, T, LBV, 1, WVV);   \
  }

}  // namespace rg
}  // namespace qcp
