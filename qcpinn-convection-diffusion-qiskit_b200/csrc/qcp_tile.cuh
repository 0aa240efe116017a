// qcpinn_b200 -- "engine T": tiled per-sample statevector path for 11 <= n <= 16 qubits
// (both precisions).  BASELINE config 4 (sim_circ_15, 16 qubits) runs here.
//
// A 2^n statevector no longer fits a warp's registers, so each stream vector lives in an
// HBM/L2-resident slab owned by the CTA (one collocation point per CTA at a time) and is processed
// in SWEEPS.  A sweep picks TB = LB + 5 index bits; every (stream, tile) pair -- a tile = the 2^TB
// amplitudes that differ only in those bits -- is loaded by ONE WARP into registers with the engine
// R layout (LB local bits per lane, 5 lane bits), runs a whole run of gates there with the engine R
// primitives (qcp_reg.cuh: register FMAs, shuffle swaps, masked controls), and is stored back.  Only
// a gate's TARGET has to be inside the tile: a control that is a tile-index bit is a warp-uniform
// predicate.  The host planner (qcp_tile.cu) cuts the gate list into sweeps greedily, and lets the
// logical->memory bit map evolve: a tile is stored with its qubits permuted so that the two qubits
// the next sweep shares with this one sit on memory bits 0 and 1 (every warp-wide access then
// covers whole 32-byte sectors).
//
// Traffic per residual point and sweep is one read + one write of the live vectors (6 forward,
// 12 in the adjoint sweeps), which is what bounds this engine (HBM roofline).
//
//   forward : generate psi0 tiles (product-state encoding) -> sweeps -> measure pass
//   backward: forward again into the slab -> lambda-init pass -> sweeps in reverse with gradient
//             accumulation (global RED rows) -> cotangent pass (pull lambda0 back to the z jets)
#pragma once

#include <algorithm>

#include "qcp_reg.cuh"

namespace qcp {
namespace tl {

using rg::C2;
using rg::C2A;
using rg::ROp;
using rg::shx;

constexpr int kMaxSweeps = 64;
constexpr int kMaxOther = 8;     // tile-index bits (n - TB <= 7)

struct Sweep {
  int32_t r0, r1, n_other, pad;
  int32_t ld_loc[32], ld_lane[32];   // element offsets of (local index i) / (lane) at load
  int32_t st_loc[32], st_lane[32];   // ... at store (same set of memory bits, permuted)
  int32_t other[kMaxOther];          // memory bits enumerated by the tile index
};

// logical-index offsets of one DIAG op occurrence (the register layout at that point of the sweep):
// amplitude (tile t, local i, lane) has logical index  sum_k bit_k(t) other[k] + lane[lane] + loc[i]
struct DiagOff {
  int32_t loc[32], lane[32], other[kMaxOther];
};

struct TlLayout {
  int rops, cs, u4, zj, qj, rj, tab, qacc, scr, warp_bytes, tabbar, rbar, total;
};

struct TlArgs {
  TlLayout lay;
  int n, enc, n_rops, n_gates, n_theta, n_consts, n_sweeps;
  int final_bit[kMaxQubitsSv];   // memory bit of logical qubit q after the last sweep
  const ROp* rops;
  const Sweep* sweeps;
  const GateOp* gates;
  const double2* consts;
  const void* theta;
  void* ws;                      // saved-jet workspace [2][n*S][B]
  long long B;
  void* slab;                    // per-CTA state storage
  size_t slab_stride;            // complex elements per CTA
  int n_blk;                     // diagonal blocks (phase tables, LOGICAL index order)
  const DiagOff* doff;           // one per DIAG op occurrence (ROp.m)
  const void* diag;              // C2A<T>[n_blk << n]
  void* w_partials;              // T[grid][n_blk << n]: sum Im(conj(lambda) psi) per block (backward)
  void* state;                   // optional per-POINT psi storage [B][S][2^n]: the forward works in it
                                 // (and leaves the final psi there), the backward starts from it
                                 // instead of recomputing the forward
  double* theta_partials;        // [grid * warps][n_theta]: one row per warp (single writer)
  void* cot_scratch;             // per CTA: [warps][NE*S] table cotangents + [NE][kMaxOther][S]
                                 // leave-one-out contributions of the cotangent pass (backward)
  size_t cot_stride;             // elements of T per CTA
};

__host__ __device__ constexpr int tl_warps(bool backward) { return backward ? 12 : 16; }

__host__ __device__ inline TlLayout tl_layout(size_t es, int LB, int S, int n, int n_rops, int n_gates,
                                              int n_consts, bool backward) {
  const int NA = 1 << LB, TB = LB + 5, NT = 1 << (n - TB), NE = NA + 32 + NT;
  const int NW = tl_warps(backward);
  TlLayout L{};
  int o = 0;
  L.rops = o; o = rg::rg_align(o + sizeof(ROp) * (size_t)n_rops);
  L.cs = o; o = rg::rg_align(o + es * 4 * (size_t)n_gates);
  L.u4 = o; o = rg::rg_align(o + es * 32 * (size_t)n_consts);
  L.zj = o; o = rg::rg_align(o + es * (size_t)n * S);
  L.qj = o; o = rg::rg_align(o + es * (size_t)n * S);
  L.rj = o; o = rg::rg_align(o + es * (size_t)n * 2 * S);
  L.tab = o; o = rg::rg_align(o + es * (size_t)NE * S);
  L.qacc = o; o = rg::rg_align(o + sizeof(double) * (size_t)n * S * NW);   // one row per warp
  // per-warp scratch of the cotangent pass (rho tile + LT components).  The sweeps move qubits
  // with shuffle SWAPs only: the shared-memory PERM of engine R measured slower here (one CTA per
  // SM, issue-bound), so the planner never emits it for this engine.
  const size_t per_warp = es * (32 * 33 + 32 * 4);
  L.warp_bytes = (int)per_warp;
  L.scr = o; L.tabbar = o; L.rbar = o;
  if (backward) {
    L.scr = o; o = rg::rg_align(o + per_warp * (size_t)NW);
    L.tabbar = o; o = rg::rg_align(o + es * (size_t)NE * S);
    L.rbar = o; o = rg::rg_align(o + es * (size_t)n * 2 * S);
  }
  L.total = o;
  return L;
}

extern __shared__ __align__(16) unsigned char tl_smem[];

template <typename T, int S>
struct Ctx {
  const TlArgs& a;
  template <typename U>
  __device__ __forceinline__ U* at(int off) const { return reinterpret_cast<U*>(tl_smem + off); }
  __device__ __forceinline__ const ROp* rops() const { return at<const ROp>(a.lay.rops); }
  __device__ __forceinline__ const T* cs() const { return at<const T>(a.lay.cs); }
  __device__ __forceinline__ const C2<T>* u4() const { return at<const C2<T>>(a.lay.u4); }
  __device__ __forceinline__ T* zj() const { return at<T>(a.lay.zj); }                          // [n*S]
  __device__ __forceinline__ T* qj() const { return at<T>(a.lay.qj); }                          // [n*S]
  __device__ __forceinline__ Jet<T, S>* rj() const { return at<Jet<T, S>>(a.lay.rj); }          // [n][2]
  __device__ __forceinline__ Jet<T, S>* tab() const { return at<Jet<T, S>>(a.lay.tab); }        // [NA | 32 | NT]
  __device__ __forceinline__ double* qacc() const { return at<double>(a.lay.qacc); }            // [NW][n*S]
  // this warp's private cotangent-pass scratch
  __device__ __forceinline__ T* scr() const { return at<T>(a.lay.scr + (threadIdx.x >> 5) * a.lay.warp_bytes); }
  __device__ __forceinline__ Jet<T, S>* tabbar() const { return at<Jet<T, S>>(a.lay.tabbar); }
  __device__ __forceinline__ Jet<T, S>* rbar() const { return at<Jet<T, S>>(a.lay.rbar); }
};

template <typename T, int S>
__device__ void load_program(const TlArgs& a) {
  ROp* rops = reinterpret_cast<ROp*>(tl_smem + a.lay.rops);
  T* cs = reinterpret_cast<T*>(tl_smem + a.lay.cs);
  C2<T>* u4 = reinterpret_cast<C2<T>*>(tl_smem + a.lay.u4);
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

// control predicate of an op for this lane / tile: 0 = skip, 1 = plain, 2 = masked (local control)
template <int LB>
__device__ __forceinline__ int ctl_mode(const ROp& op, int lane, int t) {
  constexpr int TB = LB + 5;
  if (op.pc < 0) return 1;
  if (op.pc < LB) return 2;
  if (op.pc < TB) return (lane >> (op.pc - LB)) & 1;
  return (t >> (op.pc - TB)) & 1;
}

__device__ __forceinline__ int diag_tile_offset(const DiagOff& d, int n_other, int t) {
  int o = 0;
  for (int k = 0; k < n_other; ++k) o |= ((t >> k) & 1) ? d.other[k] : 0;
  return o;
}

// DG: the program has diagonal blocks.  The gate interpreter is sensitive to the size of its switch
// (instruction fetch), so programs without them (sim_circ_15: BASELINE config 4) run a variant that
// does not contain the table code at all.
template <typename T, int LB, int S, bool DG>
__device__ __forceinline__ void run_ops_forward(T (&ax)[1 << LB], T (&ay)[1 << LB], const Ctx<T, S>& c,
                                                int r0, int r1, int lane, int t, int n_other) {
  constexpr int NA = 1 << LB;
  for (int r = r0; r < r1; ++r) {
    const ROp op = c.rops()[r];
    switch (op.kind) {
      case rg::R_DIAG: if constexpr (DG) {
        const DiagOff& d = c.a.doff[op.m];
        const C2A<T>* tab = static_cast<const C2A<T>*>(c.a.diag) + ((size_t)op.g << c.a.n) +
                            diag_tile_offset(d, n_other, t) + d.lane[lane];
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          const C2A<T> e = tab[d.loc[i]];
          const T x = ax[i], y = ay[i];
          ax[i] = fma(x, e.x, -y * e.y);
          ay[i] = fma(x, e.y, y * e.x);
        }
      } break;
      case rg::R_L1: {
        const T m0 = c.cs()[4 * op.g], m1 = c.cs()[4 * op.g + 1], m2 = c.cs()[4 * op.g + 2], m3 = c.cs()[4 * op.g + 3];
        const int mode = ctl_mode<LB>(op, lane, t);
        if (mode == 2) {
          RG_PT_SWITCH(op.pt, rg::l1_apply<T, LB, PT, true>(ax, ay, op.type, m0, m1, m2, m3, (unsigned)op.m))
        } else if (mode == 1) {
          RG_PT_SWITCH(op.pt, rg::l1_apply<T, LB, PT, false>(ax, ay, op.type, m0, m1, m2, m3, 0u))
        }
        break;
      }
      case rg::R_CX: {
        const int mode = ctl_mode<LB>(op, lane, t);
        const unsigned cm = mode == 2 ? (unsigned)op.m : (mode ? 0xffffu : 0u);
        RG_PT_SWITCH(op.pt, rg::cx_apply<T, LB, PT>(ax, ay, cm))
        break;
      }
      case rg::R_SWAP: {
        const int sh = op.pc - LB;
        const bool mybit = (lane >> sh) & 1;
        RG_PT_SWITCH(op.pt, rg::swap_ll<T, LB, PT>(ax, ay, 1 << sh, mybit))
        break;
      }
      default:
        rg::u4_apply<T, LB>(ax, ay, c.u4() + 16 * op.g, false);
        break;
    }
  }
}

template <typename T, int LB, int S, bool DG>
__device__ __forceinline__ void run_ops_backward(T (&ax)[1 << LB], T (&ay)[1 << LB], T (&lx)[1 << LB],
                                                 T (&ly)[1 << LB], const Ctx<T, S>& c, int r0, int r1,
                                                 int lane, int t, double* gth, int n_other, T* wacc) {
  constexpr int NA = 1 << LB;
  for (int r = r1 - 1; r >= r0; --r) {
    const ROp op = c.rops()[r];
    switch (op.kind) {
      case rg::R_DIAG: if constexpr (DG) {
        const DiagOff& d = c.a.doff[op.m];
        const size_t off = ((size_t)op.g << c.a.n) + diag_tile_offset(d, n_other, t) + d.lane[lane];
        const C2A<T>* tab = static_cast<const C2A<T>*>(c.a.diag) + off;
        T* wa = wacc + off;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          const int li = d.loc[i];
          atomicAdd(wa + li, fma(lx[i], ay[i], -ly[i] * ax[i]));     // W += Im(conj(lambda) psi)
          const C2A<T> e = tab[li];
          const T x = ax[i], y = ay[i], u = lx[i], v = ly[i];
          ax[i] = fma(x, e.x, y * e.y);  ay[i] = fma(y, e.x, -x * e.y);   // times conj(D)
          lx[i] = fma(u, e.x, v * e.y);  ly[i] = fma(v, e.x, -u * e.y);
        }
      } break;
      case rg::R_L1: {
        const T m0 = c.cs()[4 * op.g], m1 = c.cs()[4 * op.g + 1], m2 = c.cs()[4 * op.g + 2], m3 = c.cs()[4 * op.g + 3];
        const T d1 = op.type == rg::T_R ? m2 : -m1, d2 = op.type == rg::T_R ? m1 : m2;
        const int mode = ctl_mode<LB>(op, lane, t);
        T part = T(0);
        if (mode == 2) {
          RG_PT_SWITCH(op.pt, {
            part = rg::l1_grad<T, LB, PT, true>(ax, ay, lx, ly, op.type, (unsigned)op.m);
            rg::l1_apply<T, LB, PT, true>(ax, ay, op.type, m0, d1, d2, m3, (unsigned)op.m);
            rg::l1_apply<T, LB, PT, true>(lx, ly, op.type, m0, d1, d2, m3, (unsigned)op.m);
          })
        } else if (mode == 1) {
          RG_PT_SWITCH(op.pt, {
            if (op.p >= 0) part = rg::l1_grad<T, LB, PT, false>(ax, ay, lx, ly, op.type, 0u);
            rg::l1_apply<T, LB, PT, false>(ax, ay, op.type, m0, d1, d2, m3, 0u);
            rg::l1_apply<T, LB, PT, false>(lx, ly, op.type, m0, d1, d2, m3, 0u);
          })
        }
        if (op.p >= 0) {
          for (int m = 16; m > 0; m >>= 1) part += shx(part, m);
          if (lane == 0) atomicAdd(gth + op.p, 0.5 * (double)part);
        }
        break;
      }
      case rg::R_CX: {
        const int mode = ctl_mode<LB>(op, lane, t);
        const unsigned cm = mode == 2 ? (unsigned)op.m : (mode ? 0xffffu : 0u);
        RG_PT_SWITCH(op.pt, {
          rg::cx_apply<T, LB, PT>(ax, ay, cm);
          rg::cx_apply<T, LB, PT>(lx, ly, cm);
        })
        break;
      }
      case rg::R_SWAP: {
        const int sh = op.pc - LB;
        const bool mybit = (lane >> sh) & 1;
        RG_PT_SWITCH(op.pt, {
          rg::swap_ll<T, LB, PT>(ax, ay, 1 << sh, mybit);
          rg::swap_ll<T, LB, PT>(lx, ly, 1 << sh, mybit);
        })
        break;
      }
      default:
        rg::u4_apply<T, LB>(ax, ay, c.u4() + 16 * op.g, true);
        rg::u4_apply<T, LB>(lx, ly, c.u4() + 16 * op.g, true);
        break;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// tile addressing
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int tile_base(const Sweep& sw, int t) {
  int b = 0;
  for (int k = 0; k < sw.n_other; ++k) b |= ((t >> k) & 1) << sw.other[k];
  return b;
}

template <typename T, int LB>
__device__ __forceinline__ void tile_load(T (&ax)[1 << LB], T (&ay)[1 << LB], const C2A<T>* vec,
                                          const int32_t* loc, int lane_off) {
  constexpr int NA = 1 << LB;
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const C2A<T> v = vec[loc[i] + lane_off];
    ax[i] = v.x; ay[i] = v.y;
  }
}

template <typename T, int LB>
__device__ __forceinline__ void tile_store(const T (&ax)[1 << LB], const T (&ay)[1 << LB], C2A<T>* vec,
                                           const int32_t* loc, int lane_off) {
  constexpr int NA = 1 << LB;
#pragma unroll
  for (int i = 0; i < NA; ++i) vec[loc[i] + lane_off] = {ax[i], ay[i]};
}

// canonical tile mapping of the generate / measure / lambda-init / cotangent passes:
// memory index = (t << TB) | (i << 5) | lane   (lanes on the lowest bits: fully coalesced)
template <typename T, int LB>
__device__ __forceinline__ void canon_load(T (&ax)[1 << LB], T (&ay)[1 << LB], const C2A<T>* vec, int t, int lane) {
  constexpr int NA = 1 << LB, TB = LB + 5;
  const C2A<T>* p = vec + ((size_t)t << TB) + lane;
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const C2A<T> v = p[i << 5];
    ax[i] = v.x; ay[i] = v.y;
  }
}

template <typename T, int LB>
__device__ __forceinline__ void canon_store(const T (&ax)[1 << LB], const T (&ay)[1 << LB], C2A<T>* vec, int t, int lane) {
  constexpr int NA = 1 << LB, TB = LB + 5;
  C2A<T>* p = vec + ((size_t)t << TB) + lane;
#pragma unroll
  for (int i = 0; i < NA; ++i) p[i << 5] = {ax[i], ay[i]};
}

// ---------------------------------------------------------------------------------------------
// encoding tables.  Canonical initial layout: memory bit b holds qubit n-1-b, so
//   lane bits 0..4      <-> qubits n-1 .. n-5         table L  [32]
//   local bits 5..TB-1  <-> qubits n-6 .. n-TB        table R  [NA]
//   tile bits TB..n-1   <-> qubits n-TB-1 .. 0        table Tt [NT]
// tab = [R (NA) | L (32) | Tt (NT)]
// ---------------------------------------------------------------------------------------------
template <typename T, int S>
__device__ void encode_qubit_jets(const Ctx<T, S>& c, int n, int enc) {
  if (enc == QCP_ENC_ANGLE) {
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      Jet<T, S> h;
#pragma unroll
      for (int k = 0; k < S; ++k) h.c[k] = T(0.5) * c.zj()[j * S + k];
      T sn, cs;
      Math<T>::sincos_(h.c[0], &sn, &cs);
      c.rj()[j * 2] = jfunc(h, cs, -sn, -cs);
      c.rj()[j * 2 + 1] = jfunc(h, sn, cs, -sn);
    }
  } else if (threadIdx.x == 0) {
    Jet<T, S> nrm;
    jzero(nrm);
    for (int j = 0; j < n; ++j) {
      Jet<T, S> f;
#pragma unroll
      for (int k = 0; k < S; ++k) f.c[k] = c.zj()[j * S + k];
      jmul_acc(nrm, f, f);
    }
    const T n0 = nrm.c[0] > T(0) ? nrm.c[0] : T(1);
    const T r = T(1) / sqrt(n0);
    const T g1 = T(-0.5) * r / n0, g2 = T(0.75) * r / (n0 * n0);
    const Jet<T, S> inv = jfunc(nrm, r, g1, g2);
    for (int j = 0; j < n; ++j) {
      Jet<T, S> f;
#pragma unroll
      for (int k = 0; k < S; ++k) f.c[k] = c.zj()[j * S + k];
      c.rj()[j * 2] = jmul(f, inv);
    }
  }
}

// group of table entry `ent`: returns K = qubits in the group, q0 = qubit of the group's bit 0, idx
template <int LB>
__device__ __forceinline__ void table_group(int ent, int n, int& K, int& q0, int& idx) {
  constexpr int NA = 1 << LB, TB = LB + 5;
  if (ent < NA) { K = LB; q0 = n - 6; idx = ent; }
  else if (ent < NA + 32) { K = 5; q0 = n - 1; idx = ent - NA; }
  else { K = n - TB; q0 = n - TB - 1; idx = ent - NA - 32; }
}

template <typename T, int LB, int S>
__device__ void encode_tables(const Ctx<T, S>& c, int n) {
  constexpr int NA = 1 << LB, TB = LB + 5;
  const int NE = NA + 32 + (1 << (n - TB));
  for (int ent = threadIdx.x; ent < NE; ent += blockDim.x) {
    int K, q0, idx;
    table_group<LB>(ent, n, K, q0, idx);
    Jet<T, S> acc = c.rj()[q0 * 2 + (idx & 1)];
    for (int k = 1; k < K; ++k) acc = jmul(acc, c.rj()[(q0 - k) * 2 + ((idx >> k) & 1)]);
    c.tab()[ent] = acc;
  }
}

// psi0 tile of stream s in the canonical mapping
template <typename T, int LB, int S>
__device__ __forceinline__ void encode_tile(T (&ax)[1 << LB], T (&ay)[1 << LB], const Ctx<T, S>& c, int n,
                                            int enc, int t, int lane, int s) {
  constexpr int NA = 1 << LB;
  if (enc == QCP_ENC_ANGLE) {
    const Jet<T, S>* tab = c.tab();
    const Jet<T, S> LT = jmul(tab[NA + lane], tab[NA + 32 + t]);
    const T L0 = LT.c[0], Ls = LT.c[s], Lp2 = (S == 6 && s >= 4) ? T(2) * LT.c[s - 2] : T(0);
    const int pl = (__popc(lane) + __popc(t)) & 3;
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
      rg::phase_rot<T>(RG_POPC5(i), v * bx, v * by, ax[i], ay[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      ax[i] = (t == 0 && i == 0 && lane < n) ? c.rj()[lane * 2].c[s] : T(0);
      ay[i] = T(0);
    }
  }
}

// sign of qubit q (final memory bit fb) for amplitude (t, i, lane) in the canonical mapping
template <int LB>
__device__ __forceinline__ bool canon_bit(int fb, int t, int i, int lane) {
  constexpr int TB = LB + 5;
  return fb < 5 ? (lane >> fb) & 1 : (fb < TB ? (i >> (fb - 5)) & 1 : (t >> (fb - TB)) & 1);
}

// ---------------------------------------------------------------------------------------------
// forward of one point into the slab (shared by both kernels): generate pass + sweeps
// ---------------------------------------------------------------------------------------------
template <typename T, int LB, int S, bool DG>
__device__ void forward_point(const Ctx<T, S>& c, const TlArgs& a, C2A<T>* slab, long long p) {
  constexpr int NA = 1 << LB, TB = LB + 5;
  const int n = a.n, NT = 1 << (n - TB), nS = n * S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
  const T* ws = static_cast<const T*>(a.ws);
  for (int e = threadIdx.x; e < nS; e += blockDim.x) c.zj()[e] = ws[(size_t)e * a.B + p];
  __syncthreads();
  encode_qubit_jets<T, S>(c, n, a.enc);
  __syncthreads();
  if (a.enc == QCP_ENC_ANGLE) encode_tables<T, LB, S>(c, n);
  __syncthreads();
  const size_t M = (size_t)1 << n;
  for (int it = warp; it < S * NT; it += NW) {
    const int s = it / NT, t = it % NT;
    T ax[NA], ay[NA];
    encode_tile<T, LB, S>(ax, ay, c, n, a.enc, t, lane, s);
    canon_store<T, LB>(ax, ay, slab + (size_t)s * M, t, lane);
  }
  __syncthreads();
  for (int k = 0; k < a.n_sweeps; ++k) {
    const Sweep& sw = a.sweeps[k];
    const int ld_lane = sw.ld_lane[lane], st_lane = sw.st_lane[lane];
    for (int it = warp; it < S * NT; it += NW) {
      const int s = it / NT, t = it % NT;
      C2A<T>* vec = slab + (size_t)s * M + tile_base(sw, t);
      T ax[NA], ay[NA];
      tile_load<T, LB>(ax, ay, vec, sw.ld_loc, ld_lane);
      run_ops_forward<T, LB, S, DG>(ax, ay, c, sw.r0, sw.r1, lane, t, sw.n_other);
      tile_store<T, LB>(ax, ay, vec, sw.st_loc, st_lane);
    }
    __syncthreads();
  }
}

template <typename T, int LB, int S, bool DG>
__global__ void __launch_bounds__(tl_warps(false) * 32, 1)
tl_forward_kernel(const __grid_constant__ TlArgs a) {
  constexpr int NA = 1 << LB, TB = LB + 5;
  const Ctx<T, S> c{a};
  const int n = a.n, NT = 1 << (n - TB), nS = n * S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
  load_program<T, S>(a);
  C2A<T>* cta_slab = static_cast<C2A<T>*>(a.slab) + (size_t)blockIdx.x * a.slab_stride;
  T* ws_out = static_cast<T*>(a.ws);
  const size_t M = (size_t)1 << n;
  __syncthreads();
  for (long long p = blockIdx.x; p < a.B; p += gridDim.x) {
    // measurement sums: one accumulator row per warp (its items arrive in program order), rows
    // added in warp order afterwards -> bit-reproducible expectation values
    for (int e = threadIdx.x; e < nS * NW; e += blockDim.x) c.qacc()[e] = 0.0;
    C2A<T>* slab = a.state ? static_cast<C2A<T>*>(a.state) + (size_t)p * S * M : cta_slab;
    forward_point<T, LB, S, DG>(c, a, slab, p);
    // ---- measure pass (canonical mapping over the final layout) --------------------------------
    for (int it = warp; it < S * NT; it += NW) {
      const int s = it / NT, t = it % NT;
      T ax[NA], ay[NA], w[NA];
      canon_load<T, LB>(ax, ay, slab + (size_t)s * M, t, lane);
      if (s == 0) {
#pragma unroll
        for (int i = 0; i < NA; ++i) w[i] = fma(ax[i], ax[i], ay[i] * ay[i]);
      } else {
        const C2A<T>* p0 = slab + ((size_t)t << TB) + lane;
        const C2A<T>* pd = slab + (size_t)(s >= 4 ? s - 2 : 0) * M + ((size_t)t << TB) + lane;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          const C2A<T> v0 = p0[i << 5];
          T v = T(2) * fma(ax[i], v0.x, ay[i] * v0.y);
          if (s >= 4) {
            const C2A<T> vd = pd[i << 5];
            v = fma(T(2), fma(vd.x, vd.x, vd.y * vd.y), v);
          }
          w[i] = v;
        }
      }
      T tot = T(0), sx[LB];
#pragma unroll
      for (int x = 0; x < LB; ++x) sx[x] = T(0);
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        tot += w[i];
#pragma unroll
        for (int x = 0; x < LB; ++x) sx[x] += ((i >> x) & 1) ? -w[i] : w[i];
      }
      for (int q = 0; q < n; ++q) {
        const int fb = a.final_bit[q];
        T v;
        if (fb < 5) v = ((lane >> fb) & 1) ? -tot : tot;
        else if (fb >= TB) v = ((t >> (fb - TB)) & 1) ? -tot : tot;
        else {
          v = sx[0];
#pragma unroll
          for (int x = 1; x < LB; ++x) v = (fb - 5) == x ? sx[x] : v;
        }
        for (int m = 16; m > 0; m >>= 1) v += shx(v, m);
        if (lane == 0) c.qacc()[(size_t)warp * nS + q * S + s] += (double)v;
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nS; e += blockDim.x) {
      double tot = 0.0;
      for (int w = 0; w < NW; ++w) tot += c.qacc()[(size_t)w * nS + e];
      ws_out[(size_t)(nS + e) * a.B + p] = (T)tot;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// backward kernel
// ---------------------------------------------------------------------------------------------
template <typename T, int LB, int S, bool DG>
__global__ void __launch_bounds__(tl_warps(true) * 32, 1)
tl_backward_kernel(const __grid_constant__ TlArgs a) {
  constexpr int NA = 1 << LB, TB = LB + 5;
  const Ctx<T, S> c{a};
  const int n = a.n, NT = 1 << (n - TB), nS = n * S, NE = NA + 32 + NT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
  load_program<T, S>(a);
  C2A<T>* cta_slab = static_cast<C2A<T>*>(a.slab) + (size_t)blockIdx.x * a.slab_stride;
  const T* ws = static_cast<const T*>(a.ws);
  T* ws_out = static_cast<T*>(a.ws);
  // one dL/dtheta row per warp: every address has a single writing thread (lane 0 of that warp),
  // whose RED.ADDs apply in program order; rows are summed in order by tl_reduce_theta_kernel
  double* const gth = a.theta_partials +
                      ((size_t)blockIdx.x * NW + warp) * (a.n_theta > 0 ? a.n_theta : 1);
  T* const wacc = static_cast<T*>(a.w_partials) + ((size_t)blockIdx.x * a.n_blk << n);
  const size_t M = (size_t)1 << n;
  C2A<T>* lam = cta_slab + (size_t)S * M;
  // cotangent-pass scratch of this CTA (global: per-warp copies do not fit shared memory in float64)
  T* const tbw = static_cast<T*>(a.cot_scratch) + (size_t)blockIdx.x * a.cot_stride;   // [NW][NE*S]
  T* const contrib = tbw + (size_t)NW * NE * S;                                        // [NE][kMaxOther][S]
  __syncthreads();

  for (long long p = blockIdx.x; p < a.B; p += gridDim.x) {
    C2A<T>* slab = cta_slab;
    if (a.state) {
      // final psi saved by the forward (consumed: un-applied in place); the encoding tables are
      // still needed by the cotangent pass
      slab = static_cast<C2A<T>*>(a.state) + (size_t)p * S * M;
      for (int e = threadIdx.x; e < nS; e += blockDim.x) c.zj()[e] = ws[(size_t)e * a.B + p];
      __syncthreads();
      encode_qubit_jets<T, S>(c, n, a.enc);
      __syncthreads();
      if (a.enc == QCP_ENC_ANGLE) encode_tables<T, LB, S>(c, n);
      __syncthreads();
    } else {
      forward_point<T, LB, S, DG>(c, a, slab, p);
    }
    for (int e = threadIdx.x; e < nS; e += blockDim.x) c.qj()[e] = ws[(size_t)(nS + e) * a.B + p];
    for (int e = threadIdx.x; e < NW * NE * S; e += blockDim.x) tbw[e] = T(0);
    for (int e = threadIdx.x; e < n * 2 * S; e += blockDim.x) reinterpret_cast<T*>(c.rbar())[e] = T(0);
    __syncthreads();
    // ---- lambda-init pass: lambda streams from the q cotangents ---------------------------------
    for (int it = warp; it < S * NT; it += NW) {
      const int s = it / NT, t = it % NT;
      // sign sums: zb_k(t, i, lane) = sum_q (+-) qb[q][k]
      T zl[S];                      // lane + tile part
      T zi[S][LB];                  // per local bit: qb of the qubit sitting there (0 if none)
#pragma unroll
      for (int k = 0; k < S; ++k) {
        zl[k] = T(0);
#pragma unroll
        for (int x = 0; x < LB; ++x) zi[k][x] = T(0);
      }
      for (int q = 0; q < n; ++q) {
        const int fb = a.final_bit[q];
        if (fb >= 5 && fb < TB) {
#pragma unroll
          for (int x = 0; x < LB; ++x)
            if (fb - 5 == x) {
#pragma unroll
              for (int k = 0; k < S; ++k) zi[k][x] = c.qj()[q * S + k];
            }
        } else {
          const bool neg = fb < 5 ? (lane >> fb) & 1 : (t >> (fb - TB)) & 1;
#pragma unroll
          for (int k = 0; k < S; ++k) {
            const T v = c.qj()[q * S + k];
            zl[k] += neg ? -v : v;
          }
        }
      }
      const C2A<T>* base = slab + ((size_t)t << TB) + lane;
      T lx[NA], ly[NA];
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        T zb[S];
#pragma unroll
        for (int k = 0; k < S; ++k) {
          T v = zl[k];
#pragma unroll
          for (int x = 0; x < LB; ++x) v += ((i >> x) & 1) ? -zi[k][x] : zi[k][x];
          zb[k] = v;
        }
        const C2A<T> p0 = base[i << 5];
        T vx, vy;
        if (s == 0) {
          vx = T(2) * zb[0] * p0.x; vy = T(2) * zb[0] * p0.y;
          if constexpr (S == 6) {
#pragma unroll
            for (int k = 1; k < 6; ++k) {
              const C2A<T> pk = base[(size_t)k * M + (i << 5)];
              vx = fma(T(2) * zb[k], pk.x, vx);
              vy = fma(T(2) * zb[k], pk.y, vy);
            }
          }
        } else {
          T zs = T(0), ze = T(0);     // zb[s], zb[s + 2] without dynamic register indexing
#pragma unroll
          for (int k = 1; k < S; ++k) {
            zs = k == s ? zb[k] : zs;
            ze = k == s + 2 ? zb[k] : ze;
          }
          vx = T(2) * zs * p0.x; vy = T(2) * zs * p0.y;
          if (S == 6 && (s == 2 || s == 3)) {
            const C2A<T> ps = base[(size_t)s * M + (i << 5)];
            vx = fma(T(4) * ze, ps.x, vx);
            vy = fma(T(4) * ze, ps.y, vy);
          }
        }
        lx[i] = vx; ly[i] = vy;
      }
      canon_store<T, LB>(lx, ly, lam + (size_t)s * M, t, lane);
    }
    __syncthreads();
    // ---- sweeps in reverse -------------------------------------------------------------------------
    for (int k = a.n_sweeps - 1; k >= 0; --k) {
      const Sweep& sw = a.sweeps[k];
      const int ld_lane = sw.ld_lane[lane], st_lane = sw.st_lane[lane];
      // With diagonal blocks the W sums of a tile are accumulated from all S streams: the streams
      // of one tile then run on the SAME warp one after the other (single writer per address, program
      // order) instead of being dealt round-robin over the warps.
      const int n_items = DG ? NT * S : S * NT;
      for (int it0 = DG ? warp * S : warp; it0 < n_items; it0 += DG ? NW * S : NW)
      for (int sub = 0; sub < (DG ? S : 1); ++sub) {
        const int it = it0 + sub;
        const int s = DG ? it % S : it / NT, t = DG ? it / S : it % NT;
        const int tb = tile_base(sw, t);
        C2A<T>* vp = slab + (size_t)s * M + tb;
        C2A<T>* vl = lam + (size_t)s * M + tb;
        T ax[NA], ay[NA], lx[NA], ly[NA];
        tile_load<T, LB>(ax, ay, vp, sw.st_loc, st_lane);
        tile_load<T, LB>(lx, ly, vl, sw.st_loc, st_lane);
        run_ops_backward<T, LB, S, DG>(ax, ay, lx, ly, c, sw.r0, sw.r1, lane, t, gth, sw.n_other, wacc);
        tile_store<T, LB>(ax, ay, vp, sw.ld_loc, ld_lane);
        tile_store<T, LB>(lx, ly, vl, sw.ld_loc, ld_lane);
      }
      __syncthreads();
    }
    // ---- cotangent pass: lambda0 -> table cotangents -------------------------------------------------
    if (a.enc == QCP_ENC_ANGLE) {
      T* rho = c.scr();                                     // [32][33] rho, then LT components [32][4]
      T* ltc = rho + 32 * 33;
      const Jet<T, S>* tab = c.tab();
      // this warp's private copy of the table cotangents (every entry below has one writing lane per
      // warp): plain read-modify-writes, copies summed in warp order after the pass
      Jet<T, S>* tbar = reinterpret_cast<Jet<T, S>*>(tbw + (size_t)warp * NE * S);
      for (int it = warp; it < S * NT; it += NW) {
        const int s = it / NT, t = it % NT;
        T lx[NA], ly[NA];
        canon_load<T, LB>(lx, ly, lam + (size_t)s * M, t, lane);
        const int pl = (__popc(lane) + __popc(t)) & 3;
        const T bx = pl == 0 ? T(1) : (pl == 2 ? T(-1) : T(0));
        const T by = pl == 1 ? T(-1) : (pl == 3 ? T(1) : T(0));
        const Jet<T, S> Lj = tab[NA + lane], Tj = tab[NA + 32 + t];
        const Jet<T, S> LT = jmul(Lj, Tj);
        // rho_bar[i] = Re(conj(lambda) phase); lane-entry sums over the local index
        T a0 = T(0), as = T(0), ap = T(0);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          T fx, fy;
          rg::phase_rot<T>(RG_POPC5(i), bx, by, fx, fy);
          const T r = fma(lx[i], fx, ly[i] * fy);
          rho[i * 33 + lane] = r;
          const Jet<T, S>& R = tab[i];
          a0 = fma(r, R.c[s], a0);
          if (S == 6 && s > 0) as = fma(r, R.c[0], as);
          if (S == 6 && s >= 4) ap = fma(T(2) * r, R.c[s - 2], ap);
        }
        ltc[lane * 4] = LT.c[s];
        ltc[lane * 4 + 1] = LT.c[0];
        ltc[lane * 4 + 2] = (S == 6 && s >= 4) ? T(2) * LT.c[s - 2] : T(0);
        // cotangent jet of LT[lane] -> L[lane] and Tt[t]
        Jet<T, S> ltb;
        jzero(ltb);
        ltb.c[0] = a0;
        if (S == 6 && s > 0) ltb.c[s] += as;
        if (S == 6 && s >= 4) ltb.c[s - 2] += ap;
        Jet<T, S> lb, tb;
        jzero(lb); jzero(tb);
        jmul_pull_acc(lb, ltb, Tj);
        jmul_pull_acc(tb, ltb, Lj);
#pragma unroll
        for (int k = 0; k < S; ++k) {
          tbar[NA + lane].c[k] += lb.c[k];
          T v = tb.c[k];
          for (int m = 16; m > 0; m >>= 1) v += shx(v, m);
          if (lane == 0) tbar[NA + 32 + t].c[k] += v;
        }
        __syncwarp();
        // local-entry sums over the lanes: this lane handles local index i = lane (NA <= 32)
        if (lane < NA) {
          T r0 = T(0), rs = T(0), rp = T(0);
          for (int l = 0; l < 32; ++l) {
            const T r = rho[lane * 33 + l];
            r0 = fma(r, ltc[l * 4], r0);
            rs = fma(r, ltc[l * 4 + 1], rs);
            rp = fma(r, ltc[l * 4 + 2], rp);
          }
          Jet<T, S>& dst = tbar[lane];
          dst.c[0] += r0;
          if (S == 6 && s > 0) dst.c[s] += rs;
          if (S == 6 && s >= 4) dst.c[s - 2] += rp;
        }
      }
      __syncthreads();
      for (int e = threadIdx.x; e < NE * S; e += blockDim.x) {
        T tot = tbw[e];
        for (int w = 1; w < NW; ++w) tot += tbw[(size_t)w * NE * S + e];
        reinterpret_cast<T*>(c.tabbar())[e] = tot;
      }
      __syncthreads();
      // table entries -> one-qubit jets (leave-one-out products)
      for (int ent = threadIdx.x; ent < NE; ent += blockDim.x) {
        int K, q0, idx;
        table_group<LB>(ent, n, K, q0, idx);
        const Jet<T, S> eb = c.tabbar()[ent];
        Jet<T, S> suf[kMaxOther + 1];
        jzero(suf[K]);
        suf[K].c[0] = T(1);
        for (int k = K - 1; k >= 0; --k) suf[k] = jmul(c.rj()[(q0 - k) * 2 + ((idx >> k) & 1)], suf[k + 1]);
        Jet<T, S> pre;
        jzero(pre);
        pre.c[0] = T(1);
        for (int k = 0; k < K; ++k) {
          const int jq = q0 - k, bit = (idx >> k) & 1;
          const Jet<T, S> other = jmul(pre, suf[k + 1]);
          Jet<T, S> fb;
          jzero(fb);
          jmul_pull_acc(fb, eb, other);
          // contribution of this table entry to the jet cotangent of (qubit jq, bit): parked, then
          // summed per destination in entry order (no atomics: reproducible)
          T* slot = contrib + ((size_t)ent * kMaxOther + k) * S;
#pragma unroll
          for (int m = 0; m < S; ++m) slot[m] = fb.c[m];
          (void)bit;
          pre = jmul(pre, c.rj()[jq * 2 + bit]);
        }
      }
      __syncthreads();
      for (int d = threadIdx.x; d < n * 2; d += blockDim.x) {
        const int jq = d >> 1, bit = d & 1;
        // the table group that holds qubit jq, and jq's position k inside it (see table_group)
        int e0, cnt, K, q0;
        if (jq >= n - 5) { e0 = NA; cnt = 32; K = 5; q0 = n - 1; }
        else if (jq >= n - TB) { e0 = 0; cnt = NA; K = LB; q0 = n - 6; }
        else { e0 = NA + 32; cnt = NT; K = n - TB; q0 = n - TB - 1; }
        const int k = q0 - jq;
        (void)K;
        Jet<T, S> acc;
        jzero(acc);
        for (int idx = 0; idx < cnt; ++idx) {
          if (((idx >> k) & 1) != bit) continue;
          const T* slot = contrib + ((size_t)(e0 + idx) * kMaxOther + k) * S;
#pragma unroll
          for (int m = 0; m < S; ++m) acc.c[m] += slot[m];
        }
        c.rbar()[d] = acc;
      }
      __syncthreads();
      for (int j = threadIdx.x; j < n; j += blockDim.x) {
        Jet<T, S> h;
#pragma unroll
        for (int k = 0; k < S; ++k) h.c[k] = T(0.5) * c.zj()[j * S + k];
        T sn, cs;
        Math<T>::sincos_(h.c[0], &sn, &cs);
        Jet<T, S> hb;
        jzero(hb);
        jfunc_pull_acc(hb, c.rbar()[j * 2], h, -sn, -cs, sn);
        jfunc_pull_acc(hb, c.rbar()[j * 2 + 1], h, cs, -sn, -cs);
#pragma unroll
        for (int k = 0; k < S; ++k) ws_out[(size_t)(j * S + k) * a.B + p] = T(0.5) * hb.c[k];
      }
    } else {
      // amplitude encoding: psi0_k = e_k (real) for k < n  (t = 0, i = 0, lane = k)
      if (warp < S && lane < n) {
        const C2A<T> v = lam[(size_t)warp * M + lane];
        c.rbar()[lane * 2].c[warp] = v.x;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        Jet<T, S> nrm, invb, nb;
        jzero(nrm); jzero(invb); jzero(nb);
        for (int j = 0; j < n; ++j) {
          Jet<T, S> f;
#pragma unroll
          for (int k = 0; k < S; ++k) f.c[k] = c.zj()[j * S + k];
          jmul_acc(nrm, f, f);
        }
        const T r = T(1) / sqrt(nrm.c[0]);
        const T g1 = T(-0.5) * r / nrm.c[0], g2 = T(0.75) * r / (nrm.c[0] * nrm.c[0]);
        const T g3 = T(-1.875) * r / (nrm.c[0] * nrm.c[0] * nrm.c[0]);
        const Jet<T, S> inv = jfunc(nrm, r, g1, g2);
        for (int j = 0; j < n; ++j) {
          Jet<T, S> f;
#pragma unroll
          for (int k = 0; k < S; ++k) f.c[k] = c.zj()[j * S + k];
          jmul_pull_acc(invb, c.rbar()[j * 2], f);
        }
        jfunc_pull_acc(nb, invb, nrm, g1, g2, g3);
        for (int j = 0; j < n; ++j) {
          Jet<T, S> f, fb;
#pragma unroll
          for (int k = 0; k < S; ++k) f.c[k] = c.zj()[j * S + k];
          jzero(fb);
          jmul_pull_acc(fb, c.rbar()[j * 2], inv);
          jmul_pull_acc(fb, nb, f);
          jmul_pull_acc(fb, nb, f);
#pragma unroll
          for (int k = 0; k < S; ++k) ws_out[(size_t)(j * S + k) * a.B + p] = fb.c[k];
        }
      }
    }
    __syncthreads();
  }
}

// per-dtype launchers (qcp_tile_f32.cu / qcp_tile_f64.cu)
template <typename T>
int tl_launch(int LB, int S, bool backward, const TlArgs& a, int grid, size_t smem, cudaStream_t s);

template <typename K>
inline int tl_launch_one(K kernel, const TlArgs& a, int grid, int threads, size_t smem, cudaStream_t s,
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

#define TL_INSTANTIATE(T, LBV, DGV)                                                                    \
  if (LB == LBV && (a.n_blk > 0) == DGV) {                                                             \
    if (S == 6)                                                                                        \
      return backward ? TL_CALL(tl_backward_kernel, T, LBV, 6, DGV, true)                              \
                      : TL_CALL(tl_forward_kernel, T, LBV, 6, DGV, false);                             \
    return backward ? TL_CALL(tl_backward_kernel, T, LBV, 1, DGV, true)                                \
                    : TL_CALL(tl_forward_kernel, T, LBV, 1, DGV, false);                               \
  }

}  // namespace tl
}  // namespace qcp
