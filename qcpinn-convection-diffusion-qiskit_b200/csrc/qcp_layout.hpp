// qcpinn_b200 -- host-side layout bookkeeping shared by the engine R / engine T planners.
//
// A register tile has LB "local" positions (bits of the per-lane amplitude index) and up to five
// "lane" positions.  Dense gates need their target on a local position; the planners move qubits
// with SWAP (one local <-> lane exchange through warp shuffles) or PERM (any permutation of the
// positions through the warp's shared-memory buffer, cheaper from two exchanges on).
#pragma once

#include <algorithm>
#include <utility>
#include <vector>

#include "qcp_reg.cuh"

namespace qcp {

constexpr int kDiagMarker = 100;     // virtual op: "multiply by the phase table of block a"

// Replace runs of commuting diagonal gates (RZ / CRZ) by ONE marker op per block.
//   * a diagonal gate JOINS the last emitted block E when no non-diagonal gate emitted since E
//     touches its qubits (it commutes back to E);
//   * otherwise it waits in a PENDING block P, which is only emitted when a non-diagonal gate
//     touches one of P's qubits (gates on other qubits commute past P and are emitted before it).
// So "RZ RX per wire" layers, "RX RZ per wire" layers and the RZ layers on both sides of a layer
// boundary all collapse into single blocks.  vops = virtual op list, orig[g] = index in the original
// list (-1 for markers), dgs = the diagonal gates tagged with their block.
inline void fold_diagonals(const GateOp* ops, int n_ops, bool enable, std::vector<GateOp>& vops,
                           std::vector<int>& orig, std::vector<rg::DiagGate>& dgs, int* n_blk) {
  int blocks = 0, E = -1;
  unsigned dirtyE = 0, qubitsP = 0;
  std::vector<rg::DiagGate> P;
  auto flush = [&]() {
    if (P.empty()) return;
    E = blocks++;
    for (rg::DiagGate d : P) { d.blk = E; dgs.push_back(d); }
    P.clear();
    qubitsP = 0;
    dirtyE = 0;
    vops.push_back({kDiagMarker, E, -1, -1});
    orig.push_back(-1);
  };
  for (int g = 0; g < n_ops; ++g) {
    const GateOp op = ops[g];
    unsigned qs = 1u << op.a;
    if (op.b >= 0) qs |= 1u << op.b;
    if (enable && (op.kind == QCP_GATE_RZ || op.kind == QCP_GATE_CRZ)) {
      if (E >= 0 && !(qs & dirtyE)) dgs.push_back({E, op.kind, op.a, op.b, op.p});
      else { P.push_back({-1, op.kind, op.a, op.b, op.p}); qubitsP |= qs; }
      continue;
    }
    if (qs & qubitsP) flush();
    dirtyE |= qs;
    vops.push_back(op);
    orig.push_back(g);
  }
  flush();
  *n_blk = blocks;
}

// Dependency DAG of a gate list.  Two gates must keep their order iff they share a qubit on which at
// least one of them is not diagonal: controls, RZ / CRZ targets and diagonal-block tables are
// "Z-type" uses of a qubit and commute with each other.  prio[g] = earliest dependent gate.
struct GateDag {
  std::vector<std::vector<int>> succ;
  std::vector<int> npred, prio;
};

inline GateDag build_dag(const GateOp* ops, int n_ops, int n) {
  GateDag d;
  d.succ.resize(n_ops);
  d.npred.assign(n_ops, 0);
  d.prio.assign(n_ops, 1 << 30);
  std::vector<int> last_x(n, -1);
  std::vector<std::vector<int>> z_since(n);
  for (int j = 0; j < n_ops; ++j) {
    const GateOp& g = ops[j];
    int q[16];
    bool x[16];
    int nu = 0;
    auto use = [&](int qq, bool xx) { q[nu] = qq; x[nu++] = xx; };
    switch (g.kind) {
      case kDiagMarker: for (int qq = 0; qq < n; ++qq) use(qq, false); break;
      case QCP_GATE_RX: case QCP_GATE_RY: case QCP_GATE_H: use(g.a, true); break;
      case QCP_GATE_RZ: use(g.a, false); break;
      case QCP_GATE_CRX: case QCP_GATE_CNOT: use(g.a, false); use(g.b, true); break;
      case QCP_GATE_CRZ: use(g.a, false); use(g.b, false); break;
      default: use(g.a, true); use(g.b, true); break;
    }
    std::vector<int> pred;
    for (int u = 0; u < nu; ++u) {
      if (x[u] && !z_since[q[u]].empty()) pred.insert(pred.end(), z_since[q[u]].begin(), z_since[q[u]].end());
      else if (last_x[q[u]] >= 0) pred.push_back(last_x[q[u]]);
      if (x[u]) { last_x[q[u]] = j; z_since[q[u]].clear(); }
      else z_since[q[u]].push_back(j);
    }
    std::sort(pred.begin(), pred.end());
    pred.erase(std::unique(pred.begin(), pred.end()), pred.end());
    for (int i : pred) {
      d.succ[i].push_back(j);
      d.npred[j]++;
      d.prio[i] = std::min(d.prio[i], j);
    }
  }
  return d;
}

// A dependency-respecting gate order that keeps the gates of a few qubits together: repeatedly take
// every ready gate whose dense targets lie in the working set Q (at most `cap` qubits), else admit
// the ready gate whose earliest dependent comes first; when nothing fits, start a new working set.
// targets(g, t, &nt) = qubits gate g needs "near" (nt = 0: none).
template <typename TargetsFn>
inline std::vector<int> dag_order(const GateOp* ops, int n_ops, int n, int cap, TargetsFn targets) {
  GateDag dag = build_dag(ops, n_ops, n);
  std::vector<int> left = dag.npred, order, Q;
  std::vector<char> done(n_ops, 0);
  auto in_q = [&](int q) { return std::find(Q.begin(), Q.end(), q) != Q.end(); };
  auto missing = [&](int j) {
    int t[2], nt, add = 0;
    targets(j, t, &nt);
    for (int k = 0; k < nt; ++k)
      if (!in_q(t[k]) && (k == 0 || t[k] != t[0])) ++add;
    return add;
  };
  auto schedule = [&](int j) {
    done[j] = 1;
    order.push_back(j);
    for (int s : dag.succ[j]) --left[s];
  };
  while ((int)order.size() < n_ops) {
    bool any = false;
    for (int j = 0; j < n_ops; ++j)
      if (!done[j] && left[j] == 0 && missing(j) == 0) { schedule(j); any = true; }
    if (any) continue;
    int best = -1;
    for (int j = 0; j < n_ops; ++j) {
      if (done[j] || left[j] != 0 || (int)Q.size() + missing(j) > cap) continue;
      if (best < 0 || dag.prio[j] < dag.prio[best]) best = j;
    }
    if (best < 0) { Q.clear(); continue; }          // working set full: start the next one
    int t[2], nt;
    targets(best, t, &nt);
    for (int k = 0; k < nt; ++k)
      if (!in_q(t[k])) Q.push_back(t[k]);
    schedule(best);
  }
  return order;
}

struct LayoutTracker {
  int LB = 0;
  int lane_bits = 5;       // 6 when a vector spans two warps (every relayout is then a PERM)
  int perm_min = 4;        // exchanges from which one PERM beats a chain of SWAPs (measured)
  std::vector<int> pos;    // position of qubit q inside the tile, or -1
  std::vector<int> qat;    // qubit at tile position j
  std::vector<rg::ROp>* rops = nullptr;

  void exchange(int a, int b) {       // bookkeeping of one transposition of positions
    const int qa = qat[a], qb = qat[b];
    qat[a] = qb; qat[b] = qa;
    if (qb >= 0) pos[qb] = a;
    if (qa >= 0) pos[qa] = b;
  }

  void emit_swap(int local, int lane) {
    rops->push_back({rg::R_SWAP, local, lane, 0, 0, -1, 0, 0});
    exchange(local, lane);
  }

  // transpositions (any two positions each), applied in order
  void emit_perm(const std::vector<std::pair<int, int>>& tr) {
    if (tr.empty()) return;
    if ((int)tr.size() < perm_min) {
      // few exchanges: shuffle swaps are cheaper than the shared-memory round trip
      for (const auto& t : tr) {
        const int a = std::min(t.first, t.second), b = std::max(t.first, t.second);
        if (a == b) continue;
        if (a < LB && b >= LB) { emit_swap(a, b); continue; }
        if (a >= LB) {                   // lane <-> lane through local position 0
          emit_swap(0, a); emit_swap(0, b); emit_swap(0, a);
        } else {                         // local <-> local through the first lane position
          emit_swap(a, LB); emit_swap(b, LB); emit_swap(a, LB);
        }
      }
      return;
    }
    int cur[16];                       // cur[j] = source position of what ends up at j
    for (int j = 0; j < 16; ++j) cur[j] = j;
    for (const auto& t : tr) {
      std::swap(cur[t.first], cur[t.second]);
      exchange(t.first, t.second);
    }
    int icur[16];
    for (int j = 0; j < 16; ++j) icur[cur[j]] = j;
    // PermMasks of `src` (src[j] = source position of destination position j)
    auto masks = [&](const int* src, rg::ROp& op) {
      // bank swizzle: a local source bit x that becomes lane bit y toggles column bit y
      unsigned col[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int y = 0; y < lane_bits; ++y)
        if (src[LB + y] < LB) col[src[LB + y]] = 1u << y;
      auto slot_mask = [&](int sp) {          // slot bits toggled by the bit of source position sp
        return sp < LB ? (((1u << lane_bits) << sp) | col[sp]) : (1u << (sp - LB));
      };
      uint32_t w[6] = {0, 0, 0, 0, 0, 0};
      for (int x = 0; x < LB; ++x) w[x / 3] |= slot_mask(x) << (10 * (x % 3));
      for (int j = 0; j < LB + lane_bits; ++j) w[2 + j / 3] |= slot_mask(src[j]) << (10 * (j % 3));
      op.pc = (int32_t)w[0]; op.type = (int32_t)w[1]; op.g = (int32_t)w[2]; op.p = (int32_t)w[3];
      op.m = (int32_t)w[4]; op.pad = (int32_t)w[5];
    };
    rg::ROp fwd{rg::R_PERM, 0, 0, 0, 0, 0, 0, 0}, bwd{rg::R_PERMB, 0, 0, 0, 0, 0, 0, 0};
    masks(cur, fwd);
    masks(icur, bwd);
    rops->push_back(fwd);
    rops->push_back(bwd);
  }

  // Bring lane-resident qubit q (needed by gate g_cur) to a local position, together with the
  // following lane-resident dense targets while that evicts only qubits needed later than them.
  //   targets(g, t, &nt): dense targets of gate g (nt = 0 for gates that need none)
  template <typename TargetsFn>
  void make_local(int q, int g_cur, int g_limit, TargetsFn targets) {
    if (pos[q] < LB) return;
    auto next_use = [&](int qq, int from) {
      for (int g = from; g < g_limit; ++g) {
        int t[2], nt;
        targets(g, t, &nt);
        for (int k = 0; k < nt; ++k)
          if (t[k] == qq) return g;
      }
      return g_limit + 1;
    };
    // without PERM every exchange is a shuffle swap of its own: bring in only what is needed now
    const int max_in = perm_min > LB ? 1 : LB;
    std::vector<int> incoming{q}, in_use{g_cur};
    for (int g = g_cur + 1; g < g_limit && (int)incoming.size() < max_in; ++g) {
      int t[2], nt;
      targets(g, t, &nt);
      for (int k = 0; k < nt && (int)incoming.size() < max_in; ++k)
        if (pos[t[k]] >= LB && std::find(incoming.begin(), incoming.end(), t[k]) == incoming.end()) {
          incoming.push_back(t[k]);
          in_use.push_back(g);
        }
    }
    std::vector<std::pair<int, int>> loc;          // (next use, local position), farthest first
    for (int x = 0; x < LB; ++x) loc.push_back({next_use(qat[x], g_cur + 1), x});
    std::sort(loc.begin(), loc.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) {
      return a.first != b.first ? a.first > b.first : a.second < b.second;
    });
    std::vector<std::pair<int, int>> tr;
    for (size_t k = 0; k < incoming.size() && k < loc.size(); ++k) {
      if (k > 0 && loc[k].first <= in_use[k]) break;   // that local qubit is needed sooner
      tr.push_back({loc[k].second, pos[incoming[k]]});
    }
    emit_perm(tr);
  }

  // put qubit q on position X (any), displacing whatever sits there
  void move_to(int q, int X) {
    if (pos[q] == X) return;
    emit_perm({{pos[q], X}});
  }
};

}  // namespace qcp
