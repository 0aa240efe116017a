// qcpinn_b200 -- host-only self check of the engine R / engine T planners (no CUDA device needed).
//
// qcp_debug_check_plan() compiles a gate list with the same planners the plans use
// (reg_compile_host / tile_plan_host), then runs BOTH the logical circuit and the planned physical
// program on the CPU over a random statevector, following exactly the layout semantics the device
// kernels implement (local / lane positions, SWAP, PERM masks, phase tables, sweeps with their
// load / store bit maps, controls on tile-index bits), and reports the largest amplitude
// difference.  The CPU test-suite uses it to cover the planners for every ansatz and qubit count.
#include <cmath>
#include <complex>
#include <cstdint>
#include <vector>

#include "qcp_layout.hpp"
#include "qcp_tile.cuh"

namespace qcp {

using cd = std::complex<double>;
using rg::ROp;

void reg_compile_host(const GateOp* ops, int n_ops, int n, int LB, std::vector<ROp>& rops,
                      std::vector<std::vector<int>>& blk_pos, std::vector<rg::DiagGate>& dgs, int* meas_pos);
void tile_plan_host(const GateOp* ops, int n_ops, int n, int LB, std::vector<ROp>& rops,
                    std::vector<tl::Sweep>& sweeps, int* final_bit, std::vector<rg::DiagGate>& dgs,
                    std::vector<tl::DiagOff>& doffs);

namespace {

struct Mat2h { cd m[4]; };

Mat2h gate_matrix(int kind, double theta) {
  const double c = std::cos(0.5 * theta), s = std::sin(0.5 * theta);
  const cd I(0, 1);
  switch (kind) {
    case QCP_GATE_RX: case QCP_GATE_CRX: return {{c, -I * s, -I * s, c}};
    case QCP_GATE_RY: return {{c, -s, s, c}};
    case QCP_GATE_RZ: case QCP_GATE_CRZ: return {{cd(c, -s), 0, 0, cd(c, s)}};
    case QCP_GATE_CNOT: return {{0, 1, 1, 0}};
    default: { const double h = std::sqrt(0.5); return {{h, h, h, -h}}; }
  }
}

// gate on index bits (target bit pt, control bit pc or -1) of a plain array
void apply_1q(std::vector<cd>& v, int pt, int pc, const Mat2h& u) {
  const size_t M = v.size();
  for (size_t k = 0; k < M; ++k) {
    if ((k >> pt) & 1) continue;
    if (pc >= 0 && !((k >> pc) & 1)) continue;
    const size_t k1 = k | ((size_t)1 << pt);
    const cd a0 = v[k], a1 = v[k1];
    v[k] = u.m[0] * a0 + u.m[1] * a1;
    v[k1] = u.m[2] * a0 + u.m[3] * a1;
  }
}

void apply_u4(std::vector<cd>& v, int pa, int pb, const double* U /* 16 complex, row major */) {
  const size_t M = v.size();
  for (size_t k = 0; k < M; ++k) {
    if (((k >> pa) & 1) || ((k >> pb) & 1)) continue;
    const size_t idx[4] = {k, k | ((size_t)1 << pb), k | ((size_t)1 << pa), k | ((size_t)1 << pa) | ((size_t)1 << pb)};
    cd in[4], out[4];
    for (int j = 0; j < 4; ++j) in[j] = v[idx[j]];
    for (int r = 0; r < 4; ++r) {
      out[r] = 0;
      for (int j = 0; j < 4; ++j) out[r] += cd(U[2 * (r * 4 + j)], U[2 * (r * 4 + j) + 1]) * in[j];
    }
    for (int j = 0; j < 4; ++j) v[idx[j]] = out[j];
  }
}

void logical_circuit(std::vector<cd>& v, const GateOp* ops, int n_ops, int n, const double* theta,
                     const double* consts) {
  for (int g = 0; g < n_ops; ++g) {
    const GateOp& o = ops[g];
    if (o.kind == QCP_GATE_U4) { apply_u4(v, n - 1 - o.a, n - 1 - o.b, consts + 32 * o.p); continue; }
    const bool ctl = o.kind == QCP_GATE_CRX || o.kind == QCP_GATE_CRZ || o.kind == QCP_GATE_CNOT;
    const Mat2h u = gate_matrix(o.kind, o.p >= 0 ? theta[o.p] : 0.0);
    apply_1q(v, n - 1 - (ctl ? o.b : o.a), ctl ? n - 1 - o.a : -1, u);
  }
}

// source position of destination position j, decoded from the PermMasks read masks
int perm_source(const ROp& op, int LB, int lane_bits, int j) {
  const uint32_t w[6] = {(uint32_t)op.pc, (uint32_t)op.type, (uint32_t)op.g, (uint32_t)op.p, (uint32_t)op.m, (uint32_t)op.pad};
  const uint32_t mask = (w[2 + j / 3] >> (10 * (j % 3))) & 1023u;
  if (mask >> lane_bits) { int sp = 0; while (!((mask >> (lane_bits + sp)) & 1)) ++sp; return sp; }
  int y = 0; while (!((mask >> y) & 1)) ++y;
  return LB + y;
}

// one physical op on a tile whose index bits are the tile positions; `tile_idx` = bits of the
// tile-index (controls at positions >= tile_bits)
void apply_rop(std::vector<cd>& v, const ROp& op, int LB, int tile_bits, int tile_idx, const GateOp* ops,
               const double* theta, const double* consts) {
  switch (op.kind) {
    case rg::R_L1: case rg::R_CX: {
      int pc = op.pc;
      if (pc >= tile_bits) {
        if (!((tile_idx >> (pc - tile_bits)) & 1)) return;
        pc = -1;
      }
      if (op.kind == rg::R_CX) { apply_1q(v, op.pt, pc, gate_matrix(QCP_GATE_CNOT, 0.0)); return; }
      const GateOp& g = ops[op.g];
      apply_1q(v, op.pt, pc, gate_matrix(g.kind, g.p >= 0 ? theta[g.p] : 0.0));
      return;
    }
    case rg::R_SWAP: {
      const size_t M = v.size();
      for (size_t k = 0; k < M; ++k) {
        const int ba = (k >> op.pt) & 1, bb = (k >> op.pc) & 1;
        if (ba == 0 && bb == 1) {
          const size_t k2 = (k | ((size_t)1 << op.pt)) & ~((size_t)1 << op.pc);
          std::swap(v[k], v[k2]);
        }
      }
      return;
    }
    case rg::R_PERM: {
      std::vector<cd> out(v.size());
      for (size_t d = 0; d < v.size(); ++d) {
        size_t src = 0;
        const int lane_bits = tile_bits - LB > 5 ? 6 : 5;
        for (int j = 0; j < LB + lane_bits; ++j)
          if ((d >> j) & 1) src |= (size_t)1 << perm_source(op, LB, lane_bits, j);
        out[d] = src < v.size() ? v[src] : cd(0);
      }
      v.swap(out);
      return;
    }
    case rg::R_U4: apply_u4(v, 1, 0, consts + 32 * op.g); return;
    default: return;   // R_PERMB (reverse direction only); R_DIAG is handled by the caller
  }
}

}  // namespace
}  // namespace qcp

extern "C" int qcp_debug_check_plan(int n_qubits, int dtype, const int32_t* ops_in, int n_ops,
                                    const double* consts, int n_consts, const double* theta, int n_theta,
                                    double* max_err, int* n_phys_ops, int* n_sweeps) {
  using namespace qcp;
  (void)n_consts; (void)n_theta;
  if (!ops_in || !max_err || n_qubits < 5 || n_qubits > kMaxQubitsSv) {
    set_error("qcp_debug_check_plan: bad argument");
    return 1;
  }
  const GateOp* ops = reinterpret_cast<const GateOp*>(ops_in);
  const int n = n_qubits;
  const size_t M = (size_t)1 << n;
  std::vector<cd> ref(M), phys(M);
  uint64_t lcg = 0x9e3779b97f4a7c15ull;
  double nrm = 0.0;
  for (size_t k = 0; k < M; ++k) {
    lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
    const double re = (double)((lcg >> 11) & 0xfffff) / 1048576.0 - 0.5;
    lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
    const double im = (double)((lcg >> 11) & 0xfffff) / 1048576.0 - 0.5;
    ref[k] = cd(re, im);
    nrm += re * re + im * im;
  }
  for (size_t k = 0; k < M; ++k) ref[k] /= std::sqrt(nrm);
  phys = ref;
  logical_circuit(ref, ops, n_ops, n, theta, consts);

  std::vector<rg::ROp> rops;
  std::vector<int> where(n);              // final index bit of logical qubit q in `phys`
  int sweeps_out = 0;
  if (reg_supported(n, dtype)) {
    const int lbmax = dtype == QCP_F64 ? 4 : 5;
    const int LB = n - 1 < lbmax ? n - 1 : lbmax;
    std::vector<std::vector<int>> blk_pos;
    std::vector<rg::DiagGate> dgs;
    reg_compile_host(ops, n_ops, n, LB, rops, blk_pos, dgs, where.data());
    for (const rg::ROp& op : rops) {
      if (op.kind == rg::R_DIAG) {
        const std::vector<int>& bp = blk_pos[op.g];
        for (size_t k = 0; k < M; ++k) {
          double ang = 0.0;
          for (const rg::DiagGate& d : dgs) {
            if (d.blk != op.g) continue;
            const double half = 0.5 * theta[d.p];
            if (d.kind == QCP_GATE_RZ) ang += ((k >> bp[d.a]) & 1) ? half : -half;
            else if ((k >> bp[d.a]) & 1) ang += ((k >> bp[d.b]) & 1) ? half : -half;
          }
          phys[k] *= cd(std::cos(ang), std::sin(ang));
        }
      } else {
        apply_rop(phys, op, LB, n, 0, ops, theta, consts);
      }
    }
  } else if (tile_supported(n, dtype)) {
    const int LB = dtype == QCP_F64 ? 4 : 5, TB = LB + 5;
    std::vector<tl::Sweep> sweeps;
    std::vector<rg::DiagGate> dgs;
    std::vector<tl::DiagOff> doffs;
    tile_plan_host(ops, n_ops, n, LB, rops, sweeps, where.data(), dgs, doffs);
    sweeps_out = (int)sweeps.size();
    std::vector<cd> tile((size_t)1 << TB);
    for (const tl::Sweep& sw : sweeps) {
      for (int t = 0; t < (1 << sw.n_other); ++t) {
        size_t base = 0;
        for (int k = 0; k < sw.n_other; ++k) base |= (size_t)((t >> k) & 1) << sw.other[k];
        for (int lane = 0; lane < 32; ++lane)
          for (int i = 0; i < (1 << LB); ++i)
            tile[(size_t)i | ((size_t)lane << LB)] = phys[base + sw.ld_loc[i] + sw.ld_lane[lane]];
        for (int r = sw.r0; r < sw.r1; ++r) {
          const rg::ROp& op = rops[r];
          if (op.kind != rg::R_DIAG) { apply_rop(tile, op, LB, TB, t, ops, theta, consts); continue; }
          // phase table lookup by the LOGICAL index of every tile amplitude
          const tl::DiagOff& d = doffs[op.m];
          int toff = 0;
          for (int k = 0; k < sw.n_other; ++k) toff |= ((t >> k) & 1) ? d.other[k] : 0;
          for (int lane = 0; lane < 32; ++lane)
            for (int i = 0; i < (1 << LB); ++i) {
              const int kl = toff + d.lane[lane] + d.loc[i];
              double ang = 0.0;
              for (const rg::DiagGate& dg : dgs) {
                if (dg.blk != op.g) continue;
                const double half = 0.5 * theta[dg.p];
                if (dg.kind == QCP_GATE_RZ) ang += ((kl >> (n - 1 - dg.a)) & 1) ? half : -half;
                else if ((kl >> (n - 1 - dg.a)) & 1) ang += ((kl >> (n - 1 - dg.b)) & 1) ? half : -half;
              }
              tile[(size_t)i | ((size_t)lane << LB)] *= cd(std::cos(ang), std::sin(ang));
            }
        }
        for (int lane = 0; lane < 32; ++lane)
          for (int i = 0; i < (1 << LB); ++i)
            phys[base + sw.st_loc[i] + sw.st_lane[lane]] = tile[(size_t)i | ((size_t)lane << LB)];
      }
    }
  } else {
    set_error("qcp_debug_check_plan: no planned engine for %d qubits", n);
    return 1;
  }
  double err = 0.0;
  for (size_t k = 0; k < M; ++k) {          // k = logical index; gather its physical location
    size_t p = 0;
    for (int q = 0; q < n; ++q)
      if ((k >> (n - 1 - q)) & 1) p |= (size_t)1 << where[q];
    const double e = std::abs(ref[k] - phys[p]);
    if (e > err) err = e;
  }
  *max_err = err;
  if (n_phys_ops) *n_phys_ops = (int)rops.size();
  if (n_sweeps) *n_sweeps = sweeps_out;
  return 0;
}
