// qcpinn_b200 -- engine T host side: sweep planner and launch plumbing.  Device code: qcp_tile.cuh.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "qcp_layout.hpp"
#include "qcp_tile.cuh"

namespace qcp {

using namespace tl;
using rg::R_CX;
using rg::R_L1;
using rg::R_SWAP;
using rg::R_U4;
using rg::T_R;
using rg::T_X;
using rg::T_Z;

template <typename T>
__global__ void tl_reduce_theta_kernel(const double* __restrict__ partials, int grid, int n_theta,
                                       T* __restrict__ gtheta) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_theta) return;
  double s = 0.0;
  for (int g = 0; g < grid; ++g) s += partials[(size_t)g * n_theta + p];
  gtheta[p] = (T)s;
}

// phase tables of the diagonal blocks in LOGICAL index order: D[(blk << n) + k] = exp(i sum_g angle_g(k)),
// RZ(t) = diag(e^{-it/2}, e^{+it/2}) on bit n-1-wire of k, CRZ likewise on the control = 1 half
template <typename T>
__global__ void tl_diag_build_kernel(int n, int n_blk, const rg::DiagGate* __restrict__ dg, int n_dg,
                                     const T* __restrict__ theta, C2A<T>* __restrict__ out) {
  const long long total = (long long)n_blk << n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int blk = (int)(idx >> n), k = (int)(idx & (((long long)1 << n) - 1));
    double ang = 0.0;
    for (int g = 0; g < n_dg; ++g) {
      const rg::DiagGate d = dg[g];
      if (d.blk != blk) continue;
      const double half = 0.5 * (double)theta[d.p];
      if (d.kind == QCP_GATE_RZ) ang += ((k >> (n - 1 - d.a)) & 1) ? half : -half;
      else if ((k >> (n - 1 - d.a)) & 1) ang += ((k >> (n - 1 - d.b)) & 1) ? half : -half;
    }
    double s, c;
    sincos(ang, &s, &c);
    out[idx] = {(T)c, (T)s};
  }
}

template <typename T>
__global__ void tl_wsum_kernel(const T* __restrict__ wpart, int grid, long long total, double* __restrict__ wsum) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= total) return;
  double s = 0.0;
  for (int g = 0; g < grid; ++g) s += (double)wpart[(size_t)g * total + j];
  wsum[j] = s;
}

// dL/dtheta_g = 1/2 sum_k z_g(k) W_blk[k], z_g = eigenvalue of the generator (Z, or |1><1| (x) Z)
template <typename T>
__global__ void tl_diag_grad_kernel(int n, const rg::DiagGate* __restrict__ dg, const double* __restrict__ wsum,
                                    T* __restrict__ gtheta) {
  __shared__ double red[8];
  const rg::DiagGate d = dg[blockIdx.x];
  double acc = 0.0;
  for (int k = threadIdx.x; k < (1 << n); k += blockDim.x) {
    double z;
    if (d.kind == QCP_GATE_RZ) z = ((k >> (n - 1 - d.a)) & 1) ? -1.0 : 1.0;
    else z = ((k >> (n - 1 - d.a)) & 1) ? (((k >> (n - 1 - d.b)) & 1) ? -1.0 : 1.0) : 0.0;
    acc += z * wsum[((size_t)d.blk << n) + k];
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    gtheta[d.p] = (T)(0.5 * s);
  }
}

struct TilePlan {
  int n, enc, dtype, LB, TB, n_gates, n_theta, n_consts, n_rops, n_sweeps, num_sms;
  int kind_count[8];
  int final_bit[kMaxQubitsSv];
  ROp* d_rops;
  Sweep* d_sweeps;
  int n_blk, n_dg, n_doff;    // diagonal blocks, their gates, DIAG op occurrences
  rg::DiagGate* d_dg;
  DiagOff* d_doff;
  void* d_diag;               // C2A<T>[n_blk << n]
  void* d_wpart;
  size_t wpart_bytes;
  double* d_wsum;
  const GateOp* d_gates;      // borrowed from the owning plan
  const double2* d_consts;    // borrowed
  void* d_slab;
  size_t slab_bytes;
  double* d_tpart;
  void* d_cot = nullptr;       // cotangent-pass scratch (per-warp table cotangents, contributions)
  size_t cot_bytes = 0;
  size_t tpart_bytes;
  const void* theta;
};

static size_t es_of(int dtype) { return dtype == QCP_F64 ? 8 : 4; }

int tile_supported(int n, int dtype) {
  const char* env = std::getenv("QCP_ENGINE");
  if (env && (env[0] == 'L' || env[0] == 'l')) return 0;
  const int tb = dtype == QCP_F64 ? 9 : 10;
  return n > tb && n <= kMaxQubitsSv && n - tb < kMaxOther;
}

// qubits a gate needs INSIDE the tile (its dense / diagonal target; both wires of a Haar block)
static void gate_targets(const GateOp& g, int* t, int* nt) {
  *nt = 0;
  switch (g.kind) {
    case kDiagMarker: break;          // a table lookup by the full logical index: no residency needed
    case QCP_GATE_RX: case QCP_GATE_RY: case QCP_GATE_RZ: case QCP_GATE_H: t[(*nt)++] = g.a; break;
    case QCP_GATE_CRX: case QCP_GATE_CRZ: case QCP_GATE_CNOT: t[(*nt)++] = g.b; break;
    default: t[(*nt)++] = g.a; t[(*nt)++] = g.b; break;
  }
}

// Greedy prefix: the longest run of gates from g0 whose targets fit into `cap` tile qubits, starting
// from the qubits in `seed`.  Returns the end index; `Q` receives the tile qubits in order of first use.
static int grow_tile(const GateOp* ops, int n_ops, int g0, int cap, const std::vector<int>& seed,
                     std::vector<int>& Q) {
  Q = seed;
  int g = g0;
  for (; g < n_ops; ++g) {
    int t[2], nt;
    gate_targets(ops[g], t, &nt);
    int add = 0;
    for (int k = 0; k < nt; ++k)
      if (std::find(Q.begin(), Q.end(), t[k]) == Q.end() && (k == 0 || t[k] != t[0])) ++add;
    if ((int)Q.size() + add > cap) break;
    for (int k = 0; k < nt; ++k)
      if (std::find(Q.begin(), Q.end(), t[k]) == Q.end()) Q.push_back(t[k]);
  }
  return g;
}

static bool tile_diag_tables() {
  const char* env = std::getenv("QCP_TILE_DIAG");
  return !(env && env[0] == '0');
}

// reorder = false: sweeps are maximal runs of the gate list in program order.
// reorder = true: every sweep takes all gates the dependency DAG allows on its tile qubits, and new
// qubits join the tile in the order of the earliest blocked gate (a list scheduler); e.g. the
// RY / CNOT-chain / RY / CNOT-ring layers of sim_circ_15 run as a wavefront over the qubits
// instead of layer by layer.
static void plan_sweeps(const GateOp* ops_in, int n_ops_in, int n, int LB, bool reorder,
                        std::vector<ROp>& rops, std::vector<Sweep>& sweeps, int* final_bit,
                        std::vector<rg::DiagGate>& dgs, std::vector<DiagOff>& doffs, int* n_blk) {
  const int TB = LB + 5, NA = 1 << LB;
  std::vector<GateOp> vops;
  std::vector<int> orig;
  fold_diagonals(ops_in, n_ops_in, tile_diag_tables(), vops, orig, dgs, n_blk);
  const GateOp* ops = vops.data();
  const int n_ops = (int)vops.size();
  std::vector<int> mbit(n), qatm(n);               // memory bit of qubit q / qubit at memory bit b
  for (int q = 0; q < n; ++q) { mbit[q] = n - 1 - q; qatm[n - 1 - q] = q; }
  GateDag dag;
  std::vector<int> left;
  std::vector<char> done(n_ops, 0);
  if (reorder) {
    dag = build_dag(ops, n_ops, n);
    left = dag.npred;
  }
  auto in_set = [](const std::vector<int>& v, int q) { return std::find(v.begin(), v.end(), q) != v.end(); };
  // earliest unscheduled gate that needs qubit q inside a tile
  auto next_need = [&](int q) {
    for (int j = 0; j < n_ops; ++j) {
      if (done[j]) continue;
      int t[2], nt;
      gate_targets(ops[j], t, &nt);
      for (int k = 0; k < nt; ++k)
        if (t[k] == q) return j;
    }
    return 1 << 30;
  };
  int g = 0, n_done = 0;
  while (n_done < n_ops) {
    // ---- tile qubits of this sweep: the two on memory bits 0 / 1 are always inside ------------------
    std::vector<int> seed = {qatm[0], qatm[1]}, Q, run, next_seed;
    if (!reorder) {
      const int g_end = grow_tile(ops, n_ops, g, TB, seed, Q);
      for (int gg = g; gg < g_end; ++gg) run.push_back(gg);
      // pad to TB qubits with the ones used next (keeps them handy; the tile always has TB bits)
      for (int gg = g_end; gg < n_ops && (int)Q.size() < TB; ++gg) {
        int t[2], nt;
        gate_targets(ops[gg], t, &nt);
        for (int k = 0; k < nt && (int)Q.size() < TB; ++k)
          if (!in_set(Q, t[k])) Q.push_back(t[k]);
      }
      for (int q = 0; q < n && (int)Q.size() < TB; ++q)
        if (!in_set(Q, q)) Q.push_back(q);
      // ---- seeds of the NEXT sweep: two of this tile's qubits, preferably ones it needs anyway --------
      if (g_end < n_ops) {
        std::vector<int> Qn;
        grow_tile(ops, n_ops, g_end, TB - 2, {}, Qn);
        for (int q : Qn)
          if ((int)next_seed.size() < 2 && in_set(Q, q)) next_seed.push_back(q);
      }
      g = g_end;
    } else {
      Q = seed;
      auto schedule = [&](int j) {
        done[j] = 1;
        run.push_back(j);
        for (int s2 : dag.succ[j]) --left[s2];
      };
      auto missing = [&](int j) {
        int t[2], nt, add = 0;
        gate_targets(ops[j], t, &nt);
        for (int k = 0; k < nt; ++k)
          if (!in_set(Q, t[k]) && (k == 0 || t[k] != t[0])) ++add;
        return add;
      };
      for (;;) {
        bool any = false;
        for (int j = 0; j < n_ops; ++j)            // everything that is ready on the current tile
          if (!done[j] && left[j] == 0 && missing(j) == 0) { schedule(j); any = true; }
        if (any) continue;
        int best = -1;
        for (int j = 0; j < n_ops; ++j) {
          if (done[j] || left[j] != 0 || (int)Q.size() + missing(j) > TB) continue;
          if (best < 0 || dag.prio[j] < dag.prio[best]) best = j;
        }
        if (best < 0) break;
        int t[2], nt;
        gate_targets(ops[best], t, &nt);
        for (int k = 0; k < nt; ++k)
          if (!in_set(Q, t[k])) Q.push_back(t[k]);
        schedule(best);
      }
      // pad the tile and pick the next seeds by the earliest pending need
      std::vector<std::pair<int, int>> need;
      for (int q = 0; q < n; ++q) need.push_back({next_need(q), q});
      std::sort(need.begin(), need.end());
      for (const auto& nq : need)
        if ((int)Q.size() < TB && !in_set(Q, nq.second)) Q.push_back(nq.second);
      for (const auto& nq : need)
        if ((int)next_seed.size() < 2 && nq.first < (1 << 30) && in_set(Q, nq.second)) next_seed.push_back(nq.second);
    }
    n_done += (int)run.size();
    const int g_end = (int)run.size();              // gate loops below index `run`
    for (int k = (int)Q.size() - 1; k >= 0 && (int)next_seed.size() < 2; --k)
      if (!in_set(next_seed, Q[k])) next_seed.push_back(Q[k]);

    // ---- load mapping: tile position j <-> memory bit ---------------------------------------------------
    LayoutTracker lt;
    lt.LB = LB;
    // measured on cfg4 (16-qubit sim_circ_15): the tile kernels are issue-bound with one CTA per SM,
    // and the shared-memory round trip of a PERM stalls longer than the shuffle swaps it replaces
    lt.perm_min = 1 << 30;
    lt.pos.assign(n, -1);
    lt.qat.assign(16, -1);
    lt.rops = &rops;
    std::vector<int>& pos = lt.pos;
    std::vector<int>& qat = lt.qat;
    auto place = [&](int q, int j) { pos[q] = j; qat[j] = q; };
    place(qatm[0], LB);
    place(qatm[1], LB + 1);
    {
      int next_local = 0, next_lane = LB + 2;
      for (int q : Q) {
        if (pos[q] >= 0) continue;
        if (next_local < LB) place(q, next_local++);
        else place(q, next_lane++);
      }
    }
    Sweep sw{};
    std::vector<int> ldbit(TB), other;
    for (int j = 0; j < TB; ++j) ldbit[j] = mbit[qat[j]];
    for (int b = 0; b < n; ++b)
      if (std::find(ldbit.begin(), ldbit.end(), b) == ldbit.end()) other.push_back(b);
    sw.n_other = (int)other.size();
    for (int k = 0; k < sw.n_other; ++k) sw.other[k] = other[k];
    auto ctl_pos = [&](int q) {
      if (pos[q] >= 0) return pos[q];
      const int b = mbit[q];
      return TB + (int)(std::find(other.begin(), other.end(), b) - other.begin());
    };

    // ---- gates of the sweep in tile positions (engine R style: dense targets on local positions) ------
    sw.r0 = (int)rops.size();
    auto targets = [&](int gi, int* t, int* nt) { gate_targets(ops[run[gi]], t, nt); };
    auto make_local = [&](int q, int g_cur) { lt.make_local(q, g_cur, g_end, targets); };
    auto move_to = [&](int q, int X) { lt.move_to(q, X); };
    auto pair_mask = [&](int pt, int pc) {
      if (pc >= LB) return 0;
      int m = 0;
      for (int h = 0; h < NA / 2; ++h) {
        const int i0 = ((h >> pt) << (pt + 1)) | (h & ((1 << pt) - 1));
        if ((i0 >> pc) & 1) m |= 1 << h;
      }
      return m;
    };
    for (int gi = 0; gi < g_end; ++gi) {
      const int gg = run[gi];
      const GateOp op = ops[gg];
      switch (op.kind) {
        case kDiagMarker: {
          // logical-index offsets of the current register layout (logical bit of qubit q = n-1-q)
          DiagOff d{};
          for (int i = 0; i < 32; ++i) {
            int lo = 0, la = 0;
            for (int x = 0; x < LB; ++x)
              if ((i >> x) & 1) lo |= 1 << (n - 1 - qat[x]);
            for (int y = 0; y < 5; ++y)
              if ((i >> y) & 1) la |= 1 << (n - 1 - qat[LB + y]);
            d.loc[i] = lo; d.lane[i] = la;
          }
          for (int k = 0; k < (int)other.size(); ++k) d.other[k] = 1 << (n - 1 - qatm[other[k]]);
          rops.push_back({rg::R_DIAG, 0, -1, 0, op.a, -1, (int)doffs.size(), 0});
          doffs.push_back(d);
          break;
        }
        case QCP_GATE_RX: case QCP_GATE_RY: case QCP_GATE_RZ: case QCP_GATE_H: {
          make_local(op.a, gi);
          const int type = op.kind == QCP_GATE_RX ? T_X : (op.kind == QCP_GATE_RZ ? T_Z : T_R);
          rops.push_back({R_L1, pos[op.a], -1, type, orig[gg], op.kind == QCP_GATE_H ? -1 : op.p, 0, 0});
          break;
        }
        case QCP_GATE_CRX: case QCP_GATE_CRZ: {
          make_local(op.b, gi);
          const int pc = ctl_pos(op.a);
          rops.push_back({R_L1, pos[op.b], pc, op.kind == QCP_GATE_CRX ? T_X : T_Z, orig[gg], op.p,
                          pair_mask(pos[op.b], pc), 0});
          break;
        }
        case QCP_GATE_CNOT: {
          make_local(op.b, gi);
          const int pc = ctl_pos(op.a);
          rops.push_back({R_CX, pos[op.b], pc, 0, orig[gg], -1, pair_mask(pos[op.b], pc), 0});
          break;
        }
        default:
          move_to(op.a, 1);
          move_to(op.b, 0);
          rops.push_back({R_U4, 0, -1, 0, op.p, -1, 0, 0});
          break;
      }
    }
    // ---- the next sweep's seeds go to lane positions LB / LB + 1 (memory bits 0 / 1 at the store) ------
    for (int k = 0; k < 2; ++k) {
      const int P = LB + k, q = next_seed[k];
      if (pos[q] == P) continue;
      if (pos[q] < LB) { lt.emit_swap(pos[q], P); continue; }
      // lane -> lane: hop through a local position that does not hold the other seed
      int x = 0;
      while (qat[x] == next_seed[1 - k]) ++x;
      lt.emit_swap(x, pos[q]);
      lt.emit_swap(x, P);
    }
    sw.r1 = (int)rops.size();
    // ---- store mapping: same memory bits, positions LB / LB + 1 on bits 0 / 1 -----------------------------
    std::vector<int> rest;
    for (int j = 0; j < TB; ++j)
      if (ldbit[j] != 0 && ldbit[j] != 1) rest.push_back(ldbit[j]);
    std::sort(rest.begin(), rest.end());
    std::vector<int> stbit(TB);
    {
      int r = 0;
      for (int j = 0; j < TB; ++j) stbit[j] = j == LB ? 0 : (j == LB + 1 ? 1 : rest[r++]);
    }
    for (int i = 0; i < 32; ++i) {
      int lo = 0, so = 0, ll = 0, sl = 0;
      for (int x = 0; x < LB; ++x)
        if ((i >> x) & 1) { lo |= 1 << ldbit[x]; so |= 1 << stbit[x]; }
      for (int y = 0; y < 5; ++y)
        if ((i >> y) & 1) { ll |= 1 << ldbit[LB + y]; sl |= 1 << stbit[LB + y]; }
      sw.ld_loc[i] = lo; sw.st_loc[i] = so; sw.ld_lane[i] = ll; sw.st_lane[i] = sl;
    }
    for (int j = 0; j < TB; ++j) { mbit[qat[j]] = stbit[j]; qatm[stbit[j]] = qat[j]; }
    sweeps.push_back(sw);
  }
  for (int q = 0; q < n; ++q) final_bit[q] = mbit[q];
}

// both gate orders are planned; the one with fewer sweeps (HBM round trips) wins, ties keep the
// program order.  QCP_TILE_ORDER=program / dag forces one.
static void plan_best(const GateOp* ops, int n_ops, int n, int LB, std::vector<ROp>& rops,
                      std::vector<Sweep>& sweeps, int* final_bit, std::vector<rg::DiagGate>& dgs,
                      std::vector<DiagOff>& doffs, int* n_blk) {
  const char* env = std::getenv("QCP_TILE_ORDER");
  const bool force_program = env && env[0] == 'p', force_dag = env && env[0] == 'd';
  std::vector<ROp> r2;
  std::vector<Sweep> s2;
  std::vector<rg::DiagGate> g2;
  std::vector<DiagOff> d2;
  int fb2[kMaxQubitsSv], nb2 = 0;
  if (!force_dag) plan_sweeps(ops, n_ops, n, LB, false, rops, sweeps, final_bit, dgs, doffs, n_blk);
  if (force_program) return;
  plan_sweeps(ops, n_ops, n, LB, true, r2, s2, fb2, g2, d2, &nb2);
  if (force_dag || s2.size() < sweeps.size()) {
    rops.swap(r2); sweeps.swap(s2); dgs.swap(g2); doffs.swap(d2);
    for (int q = 0; q < n; ++q) final_bit[q] = fb2[q];
    *n_blk = nb2;
  }
}

// host-only entry for qcp_plancheck.cu
void tile_plan_host(const GateOp* ops, int n_ops, int n, int LB, std::vector<ROp>& rops,
                    std::vector<Sweep>& sweeps, int* final_bit, std::vector<rg::DiagGate>& dgs,
                    std::vector<DiagOff>& doffs) {
  int n_blk = 0;
  plan_best(ops, n_ops, n, LB, rops, sweeps, final_bit, dgs, doffs, &n_blk);
}

TilePlan* tile_create(int n, int enc, int dtype, const GateOp* host_ops, int n_ops, int n_theta,
                      int n_consts, const GateOp* d_ops, const double2* d_consts, int num_sms) {
  if (!tile_supported(n, dtype)) return nullptr;
  TilePlan* r = new TilePlan();
  memset(r, 0, sizeof(*r));
  r->n = n; r->enc = enc; r->dtype = dtype; r->n_gates = n_ops; r->n_theta = n_theta;
  r->n_consts = n_consts; r->num_sms = num_sms;
  r->LB = dtype == QCP_F64 ? 4 : 5;
  r->TB = r->LB + 5;
  r->d_gates = d_ops; r->d_consts = d_consts;
  std::vector<ROp> rops;
  std::vector<Sweep> sweeps;
  std::vector<rg::DiagGate> dgs;
  std::vector<DiagOff> doffs;
  plan_best(host_ops, n_ops, n, r->LB, rops, sweeps, r->final_bit, dgs, doffs, &r->n_blk);
  r->n_rops = (int)rops.size(); r->n_sweeps = (int)sweeps.size();
  r->n_dg = (int)dgs.size(); r->n_doff = (int)doffs.size();
  for (const ROp& o : rops) r->kind_count[o.kind & 7]++;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** ptr, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(ptr, bytes ? bytes : 16); };
  alloc((void**)&r->d_rops, sizeof(ROp) * rops.size());
  alloc((void**)&r->d_sweeps, sizeof(Sweep) * sweeps.size());
  alloc((void**)&r->d_dg, sizeof(rg::DiagGate) * dgs.size());
  alloc((void**)&r->d_doff, sizeof(DiagOff) * doffs.size());
  alloc(&r->d_diag, 2 * es_of(dtype) * ((size_t)r->n_blk << n));
  alloc((void**)&r->d_wsum, sizeof(double) * ((size_t)r->n_blk << n));
  if (e == cudaSuccess && !dgs.empty())
    e = cudaMemcpy(r->d_dg, dgs.data(), sizeof(rg::DiagGate) * dgs.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !doffs.empty())
    e = cudaMemcpy(r->d_doff, doffs.data(), sizeof(DiagOff) * doffs.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !rops.empty())
    e = cudaMemcpy(r->d_rops, rops.data(), sizeof(ROp) * rops.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !sweeps.empty())
    e = cudaMemcpy(r->d_sweeps, sweeps.data(), sizeof(Sweep) * sweeps.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("engine T: CUDA allocation/copy failed: %s", cudaGetErrorString(e));
    tile_destroy(r);
    return nullptr;
  }
  return r;
}

void tile_destroy(TilePlan* r) {
  if (!r) return;
  cudaFree(r->d_rops); cudaFree(r->d_sweeps); cudaFree(r->d_slab); cudaFree(r->d_tpart);
  cudaFree(r->d_dg); cudaFree(r->d_doff); cudaFree(r->d_diag); cudaFree(r->d_wpart); cudaFree(r->d_wsum); cudaFree(r->d_cot);
  delete r;
}

int tile_prepare(TilePlan* r, const void* d_theta, cudaStream_t s) {
  r->theta = d_theta;
  if (r->n_blk == 0) return 0;
  const long long total = (long long)r->n_blk << r->n;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  if (r->dtype == QCP_F64)
    tl_diag_build_kernel<double><<<blocks, 256, 0, s>>>(r->n, r->n_blk, r->d_dg, r->n_dg,
        static_cast<const double*>(d_theta), static_cast<C2A<double>*>(r->d_diag));
  else
    tl_diag_build_kernel<float><<<blocks, 256, 0, s>>>(r->n, r->n_blk, r->d_dg, r->n_dg,
        static_cast<const float*>(d_theta), static_cast<C2A<float>*>(r->d_diag));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("engine T: table build launch failed: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

int tile_num_sweeps(const TilePlan* r) { return r ? r->n_sweeps : 0; }

int tile_describe(const TilePlan* r, char* buf, int len) {
  return snprintf(buf, len, "engine=tiled n=%d LB=%d tile_bits=%d sweeps=%d ops=%d (dense=%d cnot=%d swap=%d "
                  "diag_blocks=%d covering %d diagonal gates, haar=%d)",
                  r->n, r->LB, r->TB, r->n_sweeps, r->n_rops, r->kind_count[R_L1], r->kind_count[R_CX],
                  r->kind_count[R_SWAP], r->n_blk, r->n_dg, r->kind_count[R_U4]);
}

static int grow(void** ptr, size_t* have, size_t want) {
  if (want <= *have) return 0;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr; *have = 0;
  cudaError_t e = cudaMalloc(ptr, want);
  if (e != cudaSuccess) { set_error("engine T: cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); return 1; }
  *have = want;
  return 0;
}

long long tile_state_elems(const TilePlan* r, long long B, int S) {
  return 2LL * S * B << r->n;
}

int tile_run(TilePlan* r, int S, bool backward, void* ws, long long B, void* state, void* grad_theta,
             cudaStream_t s) {
  if (S != 1 && S != 6) { set_error("engine T: bad stream count %d", S); return 1; }
  if (!r->theta && r->n_theta > 0) { set_error("engine T: qcp_prepare() has not run"); return 1; }
  const size_t es = es_of(r->dtype);
  const TlLayout L = tl_layout(es, r->LB, S, r->n, r->n_rops, r->n_gates, r->n_consts, backward);
  if (L.total > 227 * 1024) { set_error("engine T: gate program needs %d bytes of shared memory", L.total); return 1; }
  int grid = r->num_sms;
  if ((long long)grid > B) grid = (int)B;
  if (grid < 1) grid = 1;
  const size_t stride = (size_t)(backward ? 2 : 1) * S << r->n;          // complex elements per CTA
  if (grow(&r->d_slab, &r->slab_bytes, 2 * es * stride * (size_t)grid)) return 1;

  TlArgs a{};
  a.lay = L;
  a.n = r->n; a.enc = r->enc; a.n_rops = r->n_rops; a.n_gates = r->n_gates; a.n_theta = r->n_theta;
  a.n_consts = r->n_consts; a.n_sweeps = r->n_sweeps;
  for (int q = 0; q < r->n; ++q) a.final_bit[q] = r->final_bit[q];
  a.rops = r->d_rops; a.sweeps = r->d_sweeps; a.gates = r->d_gates; a.consts = r->d_consts;
  a.theta = r->theta; a.ws = ws; a.B = B; a.slab = r->d_slab; a.slab_stride = stride;
  a.state = state;
  a.n_blk = r->n_blk; a.doff = r->d_doff; a.diag = r->d_diag;
  if (backward) {
    const int nt = r->n_theta > 0 ? r->n_theta : 1;
    const int rows = grid * tl_warps(true);           // theta partials: one row per warp
    if (grow((void**)&r->d_tpart, &r->tpart_bytes, sizeof(double) * (size_t)rows * nt)) return 1;
    a.theta_partials = r->d_tpart;
    {
      const int NA = 1 << r->LB, NE = NA + 32 + (1 << (r->n - (r->LB + 5)));
      a.cot_stride = (size_t)tl_warps(true) * NE * S + (size_t)NE * kMaxOther * S;
      if (grow(&r->d_cot, &r->cot_bytes, es * a.cot_stride * (size_t)grid)) return 1;
      a.cot_scratch = r->d_cot;
    }
    const size_t wbytes = es * ((size_t)grid * (r->n_blk > 0 ? r->n_blk : 1) << r->n);
    if (grow(&r->d_wpart, &r->wpart_bytes, wbytes)) return 1;
    a.w_partials = r->d_wpart;
    if (r->n_blk > 0 && cudaMemsetAsync(r->d_wpart, 0, wbytes, s) != cudaSuccess) {
      set_error("engine T: cudaMemsetAsync failed"); return 1;
    }
    cudaError_t ez = cudaMemsetAsync(r->d_tpart, 0, sizeof(double) * (size_t)rows * nt, s);
    if (ez != cudaSuccess) { set_error("engine T: cudaMemsetAsync failed: %s", cudaGetErrorString(ez)); return 1; }
  }
  const int rc = r->dtype == QCP_F64 ? tl_launch<double>(r->LB, S, backward, a, grid, L.total, s)
                                     : tl_launch<float>(r->LB, S, backward, a, grid, L.total, s);
  if (rc) return rc;
  if (!backward) return 0;
  const int tb = (r->n_theta + 127) / 128;
  if (tb) {
    if (r->dtype == QCP_F64)
      tl_reduce_theta_kernel<double><<<tb, 128, 0, s>>>(r->d_tpart, grid * tl_warps(true), r->n_theta, static_cast<double*>(grad_theta));
    else
      tl_reduce_theta_kernel<float><<<tb, 128, 0, s>>>(r->d_tpart, grid * tl_warps(true), r->n_theta, static_cast<float*>(grad_theta));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("engine T: reduction launch failed: %s", cudaGetErrorString(e)); return 1; }
  }
  if (r->n_blk > 0) {
    const long long total = (long long)r->n_blk << r->n;
    const int wb = (int)((total + 127) / 128);
    if (r->dtype == QCP_F64) {
      tl_wsum_kernel<double><<<wb, 128, 0, s>>>(static_cast<const double*>(r->d_wpart), grid, total, r->d_wsum);
      tl_diag_grad_kernel<double><<<r->n_dg, 256, 0, s>>>(r->n, r->d_dg, r->d_wsum, static_cast<double*>(grad_theta));
    } else {
      tl_wsum_kernel<float><<<wb, 128, 0, s>>>(static_cast<const float*>(r->d_wpart), grid, total, r->d_wsum);
      tl_diag_grad_kernel<float><<<r->n_dg, 256, 0, s>>>(r->n, r->d_dg, r->d_wsum, static_cast<float*>(grad_theta));
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("engine T: diagonal gradient launch failed: %s", cudaGetErrorString(e)); return 1; }
  }
  return 0;
}

}  // namespace qcp
