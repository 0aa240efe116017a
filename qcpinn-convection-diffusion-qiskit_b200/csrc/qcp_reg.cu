// qcpinn_b200 -- engine R host side: physical-program compiler, phase-table builder, gradient
// projection of the diagonal blocks, launch plumbing.  Device code: qcp_reg.cuh.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "qcp_layout.hpp"
#include "qcp_reg.cuh"

namespace qcp {

using namespace rg;

// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------
struct BlkPos {
  int32_t pos[kMaxQubitsReg];
};

// phase tables of the diagonal blocks: D[(blk << n) + i * G + lane] = exp(i * sum_g angle_g(k)),
// RZ(t) = diag(e^{-it/2}, e^{+it/2}), CRZ likewise on the control = 1 half
template <typename T>
__global__ void rg_diag_build_kernel(int n, int LB, int n_blk, const BlkPos* __restrict__ bpos,
                                     const DiagGate* __restrict__ dg, int n_dg,
                                     const T* __restrict__ theta, C2A<T>* __restrict__ out) {
  const int G = 1 << (n - LB);
  const int total = n_blk << n;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int blk = idx >> n, e = idx & ((1 << n) - 1);
    const int phys = ((e % G) << LB) | (e / G);
    const BlkPos bp = bpos[blk];
    double ang = 0.0;
    for (int g = 0; g < n_dg; ++g) {
      const DiagGate d = dg[g];
      if (d.blk != blk) continue;
      const double half = 0.5 * (double)theta[d.p];
      if (d.kind == QCP_GATE_RZ) {
        ang += ((phys >> bp.pos[d.a]) & 1) ? half : -half;
      } else if ((phys >> bp.pos[d.a]) & 1) {
        ang += ((phys >> bp.pos[d.b]) & 1) ? half : -half;
      }
    }
    double s, c;
    sincos(ang, &s, &c);
    out[idx] = {(T)c, (T)s};
  }
}

template <typename T>
__global__ void rg_reduce_theta_kernel(const double* __restrict__ partials, int grid, int n_theta,
                                       T* __restrict__ gtheta) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_theta) return;
  double s = 0.0;
  for (int g = 0; g < grid; ++g) s += partials[(size_t)g * n_theta + p];
  gtheta[p] = (T)s;
}

template <typename T>
__global__ void rg_wsum_kernel(const T* __restrict__ wpart, int grid, int total, double* __restrict__ wsum) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= total) return;
  double s = 0.0;
  for (int g = 0; g < grid; ++g) s += (double)wpart[(size_t)g * total + j];
  wsum[j] = s;
}

// dL/dtheta_g = 1/2 sum_k z_g(k) W_blk[k], z_g = eigenvalue of the generator (Z, or |1><1| (x) Z)
template <typename T>
__global__ void rg_diag_grad_kernel(int n, int LB, const BlkPos* __restrict__ bpos,
                                    const DiagGate* __restrict__ dg, const double* __restrict__ wsum,
                                    T* __restrict__ gtheta) {
  __shared__ double red[8];
  const DiagGate d = dg[blockIdx.x];
  const BlkPos bp = bpos[d.blk];
  const int G = 1 << (n - LB);
  double acc = 0.0;
  for (int e = threadIdx.x; e < (1 << n); e += blockDim.x) {
    const int phys = ((e % G) << LB) | (e / G);
    double z;
    if (d.kind == QCP_GATE_RZ) z = ((phys >> bp.pos[d.a]) & 1) ? -1.0 : 1.0;
    else z = ((phys >> bp.pos[d.a]) & 1) ? (((phys >> bp.pos[d.b]) & 1) ? -1.0 : 1.0) : 0.0;
    acc += z * wsum[((size_t)d.blk << n) + e];
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

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
struct RegPlan {
  int WV;                     // warps per stream vector (2 when n - LB = 6)
  int n, enc, dtype, LB, G, n_gates, n_theta, n_consts, n_rops, n_blk, n_dg, num_sms;
  int kind_count[8];
  int meas_pos[kMaxQubitsReg];
  ROp* d_rops;
  const GateOp* d_gates;      // borrowed from the owning plan
  const double2* d_consts;    // borrowed
  BlkPos* d_bpos;
  DiagGate* d_dg;
  void* d_diag;
  void* d_wpart;
  size_t wpart_bytes;
  double* d_wsum;
  double* d_tpart;
  size_t tpart_bytes;
  int occ[2][2];              // blocks/SM cache [S == 6][backward]
  const void* theta;          // device copy owned by the plan (set by reg_prepare)
};

static size_t es_of(int dtype) { return dtype == QCP_F64 ? 8 : 4; }

int reg_supported(int n, int dtype) {
  const char* env = std::getenv("QCP_ENGINE");
  if (env && (env[0] == 'L' || env[0] == 'l')) return 0;
  if (n < 5) return 0;
  return n <= kMaxQubitsReg;     // float64 at n = 10: two warps per vector
}

// Translate the logical gate list into physical ops (see the header comment of qcp_reg.cuh).
// reorder: run the gates in the dependency-respecting order of dag_order() (gates of the same few
// qubits together -> fewer local<->lane relayouts) instead of program order
static void compile_physical(const GateOp* ops_in, int n_ops_in, int n, int LB, bool reorder,
                             std::vector<ROp>& rops, std::vector<BlkPos>& bpos, std::vector<DiagGate>& dgs,
                             int* meas_pos) {
  std::vector<GateOp> vops;
  std::vector<int> orig;
  int n_blk = 0;
  fold_diagonals(ops_in, n_ops_in, true, vops, orig, dgs, &n_blk);
  const GateOp* ops = vops.data();
  const int n_ops = (int)vops.size();
  LayoutTracker lt;
  lt.LB = LB;
  if (n - LB > 5) { lt.lane_bits = 6; lt.perm_min = 1; }   // warp-pair vectors: relayout = PERM only
  lt.pos.assign(n, -1);
  lt.qat.assign(16, -1);
  lt.rops = &rops;
  for (int q = 0; q < n; ++q) { lt.pos[q] = n - 1 - q; lt.qat[n - 1 - q] = q; }
  std::vector<int>& pos = lt.pos;
  auto gate_targets = [&](int g, int* t, int* nt) {
    *nt = 0;
    switch (ops[g].kind) {
      case QCP_GATE_RX: case QCP_GATE_RY: case QCP_GATE_H: t[(*nt)++] = ops[g].a; break;
      case QCP_GATE_CRX: case QCP_GATE_CNOT: t[(*nt)++] = ops[g].b; break;
      case QCP_GATE_U4: t[(*nt)++] = ops[g].a; t[(*nt)++] = ops[g].b; break;
      default: break;     // diagonal gates live in phase tables: no local target needed
    }
  };
  std::vector<int> order(n_ops);
  for (int g = 0; g < n_ops; ++g) order[g] = g;
  if (reorder) order = dag_order(ops, n_ops, n, LB, gate_targets);
  auto targets = [&](int gi, int* t, int* nt) { gate_targets(order[gi], t, nt); };   // by position
  auto make_local = [&](int q, int g_cur) { lt.make_local(q, g_cur, n_ops, targets); };
  auto move_to = [&](int q, int X) { lt.move_to(q, X); };

  // pairs h (target bit removed) whose amplitudes have the LOCAL control bit set
  auto pair_mask = [&](int pt, int pc) {
    if (pc >= LB) return 0;
    int m = 0;
    for (int h = 0; h < (1 << (LB - 1)); ++h) {
      const int i0 = ((h >> pt) << (pt + 1)) | (h & ((1 << pt) - 1));
      if ((i0 >> pc) & 1) m |= 1 << h;
    }
    return m;
  };
  for (int gi = 0; gi < n_ops; ++gi) {
    const int g = order[gi];
    const GateOp op = ops[g];
    if (op.kind == kDiagMarker) {          // phase table of block op.a, in the layout of this point
      BlkPos bp{};
      for (int q = 0; q < n; ++q) bp.pos[q] = pos[q];
      if ((int)bpos.size() <= op.a) bpos.resize(op.a + 1);
      bpos[op.a] = bp;
      rops.push_back({R_DIAG, 0, -1, 0, op.a, -1, 0, 0});
      continue;
    }
    switch (op.kind) {
      case QCP_GATE_RX: case QCP_GATE_RY: case QCP_GATE_H:
        make_local(op.a, gi);
        rops.push_back({R_L1, pos[op.a], -1, op.kind == QCP_GATE_RX ? T_X : T_R, orig[g],
                        op.kind == QCP_GATE_H ? -1 : op.p, 0, 0});
        break;
      case QCP_GATE_CRX:
        make_local(op.b, gi);
        rops.push_back({R_L1, pos[op.b], pos[op.a], T_X, orig[g], op.p, pair_mask(pos[op.b], pos[op.a]), 0});
        break;
      case QCP_GATE_CNOT:
        make_local(op.b, gi);
        rops.push_back({R_CX, pos[op.b], pos[op.a], 0, g, -1, pair_mask(pos[op.b], pos[op.a]), 0});
        break;
      default:   // U4 on (wire_hi = a, wire_lo = b): local positions (1, 0)
        move_to(op.a, 1);
        move_to(op.b, 0);
        rops.push_back({R_U4, 0, -1, 0, op.p, -1, 0, 0});
        break;
    }
  }
  for (int q = 0; q < n; ++q) meas_pos[q] = pos[q];
}

// both gate orders are compiled; the shorter physical program wins (ties keep the program order).
// QCP_REG_ORDER=program / dag forces one.
static void compile_best(const GateOp* ops, int n_ops, int n, int LB, std::vector<ROp>& rops,
                         std::vector<BlkPos>& bpos, std::vector<DiagGate>& dgs, int* meas_pos) {
  const char* env = std::getenv("QCP_REG_ORDER");
  const bool force_program = env && env[0] == 'p', force_dag = env && env[0] == 'd';
  std::vector<ROp> r2;
  std::vector<BlkPos> b2;
  std::vector<DiagGate> g2;
  int m2[16];
  if (!force_dag) compile_physical(ops, n_ops, n, LB, false, rops, bpos, dgs, meas_pos);
  if (force_program) return;
  compile_physical(ops, n_ops, n, LB, true, r2, b2, g2, m2);
  if (force_dag || r2.size() < rops.size()) {
    rops.swap(r2); bpos.swap(b2); dgs.swap(g2);
    for (int q = 0; q < n; ++q) meas_pos[q] = m2[q];
  }
}

// host-only entry for qcp_plancheck.cu
void reg_compile_host(const GateOp* ops, int n_ops, int n, int LB, std::vector<ROp>& rops,
                      std::vector<std::vector<int>>& blk_pos, std::vector<DiagGate>& dgs, int* meas_pos) {
  std::vector<BlkPos> bpos;
  compile_best(ops, n_ops, n, LB, rops, bpos, dgs, meas_pos);
  for (const BlkPos& b : bpos) blk_pos.emplace_back(b.pos, b.pos + n);
}

RegPlan* reg_create(int n, int enc, int dtype, const GateOp* host_ops, int n_ops, int n_theta,
                    int n_consts, const GateOp* d_ops, const double2* d_consts, int num_sms) {
  if (!reg_supported(n, dtype)) return nullptr;
  RegPlan* r = new RegPlan();
  memset(r, 0, sizeof(*r));
  r->n = n; r->enc = enc; r->dtype = dtype; r->n_gates = n_ops; r->n_theta = n_theta;
  r->n_consts = n_consts; r->num_sms = num_sms;
  const int lbmax = dtype == QCP_F64 ? 4 : 5;
  r->LB = n - 1 < lbmax ? n - 1 : lbmax;
  r->G = 1 << (n - r->LB);
  r->WV = r->G > 32 ? 2 : 1;
  r->d_gates = d_ops; r->d_consts = d_consts;
  std::vector<ROp> rops;
  std::vector<BlkPos> bpos;
  std::vector<DiagGate> dgs;
  compile_best(host_ops, n_ops, n, r->LB, rops, bpos, dgs, r->meas_pos);
  r->n_rops = (int)rops.size(); r->n_blk = (int)bpos.size(); r->n_dg = (int)dgs.size();
  for (const ROp& o : rops) r->kind_count[o.kind & 7]++;
  const size_t es = es_of(dtype);
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** ptr, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(ptr, bytes ? bytes : 16); };
  alloc((void**)&r->d_rops, sizeof(ROp) * rops.size());
  alloc((void**)&r->d_bpos, sizeof(BlkPos) * bpos.size());
  alloc((void**)&r->d_dg, sizeof(DiagGate) * dgs.size());
  alloc(&r->d_diag, 2 * es * ((size_t)r->n_blk << n));
  alloc((void**)&r->d_wsum, sizeof(double) * ((size_t)r->n_blk << n));
  auto put = [&](void* dst, const void* src, size_t bytes) {
    if (e == cudaSuccess && bytes) e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
  };
  put(r->d_rops, rops.data(), sizeof(ROp) * rops.size());
  put(r->d_bpos, bpos.data(), sizeof(BlkPos) * bpos.size());
  put(r->d_dg, dgs.data(), sizeof(DiagGate) * dgs.size());
  if (e != cudaSuccess) {
    set_error("engine R: CUDA allocation/copy failed: %s", cudaGetErrorString(e));
    reg_destroy(r);
    return nullptr;
  }
  return r;
}

void reg_destroy(RegPlan* r) {
  if (!r) return;
  cudaFree(r->d_rops); cudaFree(r->d_bpos); cudaFree(r->d_dg); cudaFree(r->d_diag);
  cudaFree(r->d_wpart); cudaFree(r->d_wsum); cudaFree(r->d_tpart);
  delete r;
}

int reg_prepare(RegPlan* r, const void* d_theta, cudaStream_t s) {
  r->theta = d_theta;
  if (r->n_blk == 0) return 0;
  const int total = r->n_blk << r->n;
  const int blocks = (total + 255) / 256;
  if (r->dtype == QCP_F64)
    rg_diag_build_kernel<double><<<blocks, 256, 0, s>>>(r->n, r->LB, r->n_blk, r->d_bpos, r->d_dg, r->n_dg,
        static_cast<const double*>(d_theta), static_cast<C2A<double>*>(r->d_diag));
  else
    rg_diag_build_kernel<float><<<blocks, 256, 0, s>>>(r->n, r->LB, r->n_blk, r->d_bpos, r->d_dg, r->n_dg,
        static_cast<const float*>(d_theta), static_cast<C2A<float>*>(r->d_diag));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("engine R: table build launch failed: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

int reg_describe(const RegPlan* r, char* buf, int len) {
  return snprintf(buf, len, "engine=register n=%d LB=%d lanes_per_vector=%d ops=%d (dense=%d cnot=%d swap=%d diag_blocks=%d "
                  "covering %d diagonal gates, haar=%d)", r->n, r->LB, r->G, r->n_rops, r->kind_count[R_L1],
                  r->kind_count[R_CX], r->kind_count[R_SWAP], r->n_blk, r->n_dg, r->kind_count[R_U4]);
}

long long reg_state_elems(const RegPlan* r, long long B, int S) {
  return 2LL * S * B << r->n;
}

int reg_launch_count(const RegPlan* r, bool backward) {
  if (!backward) return 1;
  return r->n_blk > 0 ? 4 : 2;
}

static int grow(void** ptr, size_t* have, size_t want) {
  if (want <= *have) return 0;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr; *have = 0;
  cudaError_t e = cudaMalloc(ptr, want);
  if (e != cudaSuccess) { set_error("engine R: cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); return 1; }
  *have = want;
  return 0;
}

int reg_run(RegPlan* r, int S, bool backward, void* ws, long long B, void* state, void* grad_theta,
            cudaStream_t s) {
  if (S != 1 && S != 6) { set_error("engine R: bad stream count %d", S); return 1; }
  if (!r->theta && r->n_theta > 0) { set_error("engine R: qcp_prepare() has not run"); return 1; }
  const size_t es = es_of(r->dtype);
  const SmemLayout L = rg_layout(es, r->LB, S, r->n, r->n_rops, r->n_gates, r->n_consts, r->n_theta,
                                 r->n_blk, backward);
  if (L.total > 227 * 1024) {
    set_error("engine R: gate program needs %d bytes of shared memory", L.total);
    return 1;
  }
  int& occ = r->occ[S == 6][backward];
  if (occ == 0) {
    const int rc = r->dtype == QCP_F64 ? rg_occupancy<double>(r->LB, r->WV, S, backward, L.total, &occ)
                                       : rg_occupancy<float>(r->LB, r->WV, S, backward, L.total, &occ);
    if (rc || occ < 1) { set_error("engine R: kernel does not fit on an SM (smem %d)", L.total); occ = 0; return 1; }
  }
  const int PP = r->WV == 2 ? 1 : 32 / r->G, NPT = S == 6 ? PP : (rg_warps(S, r->WV) / r->WV) * PP;
  long long want = (B + NPT - 1) / NPT;
  int grid = r->num_sms * occ;
  if ((long long)grid > want) grid = (int)want;
  if (grid < 1) grid = 1;

  RgArgs a{};
  a.lay = L;
  a.n = r->n; a.enc = r->enc; a.n_rops = r->n_rops; a.n_gates = r->n_gates; a.n_theta = r->n_theta;
  a.n_blk = r->n_blk; a.n_consts = r->n_consts;
  for (int q = 0; q < r->n; ++q) a.meas_pos[q] = r->meas_pos[q];
  a.rops = r->d_rops; a.gates = r->d_gates; a.consts = r->d_consts; a.theta = r->theta;
  a.diag = r->d_diag; a.ws = ws; a.B = B; a.state = state;
  const int rows = grid * rg_warps(S, r->WV);      // theta partials: one row per warp (deterministic sums)
  if (backward) {
    const int nt = r->n_theta > 0 ? r->n_theta : 1;
    if (grow((void**)&r->d_tpart, &r->tpart_bytes, sizeof(double) * (size_t)rows * nt)) return 1;
    if (grow(&r->d_wpart, &r->wpart_bytes, es * ((size_t)grid * (r->n_blk > 0 ? r->n_blk : 1) << r->n))) return 1;
    a.theta_partials = r->d_tpart;
    a.w_partials = r->d_wpart;
    cudaError_t ez = cudaMemsetAsync(r->d_tpart, 0, sizeof(double) * (size_t)rows * nt, s);
    if (ez == cudaSuccess && r->n_blk > 0)
      ez = cudaMemsetAsync(r->d_wpart, 0, es * ((size_t)grid * r->n_blk << r->n), s);
    if (ez != cudaSuccess) { set_error("engine R: cudaMemsetAsync failed: %s", cudaGetErrorString(ez)); return 1; }
  }
  const int rc = r->dtype == QCP_F64 ? rg_launch<double>(r->LB, r->WV, S, backward, a, grid, L.total, s)
                                     : rg_launch<float>(r->LB, r->WV, S, backward, a, grid, L.total, s);
  if (rc) return rc;
  if (!backward) return 0;

  const int tb = (r->n_theta + 127) / 128;
  const int total = r->n_blk << r->n;
  if (r->dtype == QCP_F64) {
    if (tb) rg_reduce_theta_kernel<double><<<tb, 128, 0, s>>>(r->d_tpart, rows, r->n_theta, static_cast<double*>(grad_theta));
    if (total) {
      rg_wsum_kernel<double><<<(total + 127) / 128, 128, 0, s>>>(static_cast<const double*>(r->d_wpart), grid, total, r->d_wsum);
      rg_diag_grad_kernel<double><<<r->n_dg, 256, 0, s>>>(r->n, r->LB, r->d_bpos, r->d_dg, r->d_wsum, static_cast<double*>(grad_theta));
    }
  } else {
    if (tb) rg_reduce_theta_kernel<float><<<tb, 128, 0, s>>>(r->d_tpart, rows, r->n_theta, static_cast<float*>(grad_theta));
    if (total) {
      rg_wsum_kernel<float><<<(total + 127) / 128, 128, 0, s>>>(static_cast<const float*>(r->d_wpart), grid, total, r->d_wsum);
      rg_diag_grad_kernel<float><<<r->n_dg, 256, 0, s>>>(r->n, r->LB, r->d_bpos, r->d_dg, r->d_wsum, static_cast<float*>(grad_theta));
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("engine R: reduction launch failed: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // namespace qcp
