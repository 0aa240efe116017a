// qcpinn_b200 -- fused collocation-point mapping + analytic target (reference
// data/diffusion_dataset.py:12-38).  One kernel replaces the ~15 (u) / ~30 (r) element-wise torch
// launches of Sampler.sample(): points = lo + (hi - lo) * rand, then the Gaussian-pulse solution u or
// the closed-form forcing r evaluated at those points.  The random numbers still come from
// torch.rand, so the sampling stream is the reference's.
#include "qcp_common.cuh"

namespace qcp {

struct Box {
  float lo[3], hi[3];
};

__global__ void sample_targets_kernel(const float* __restrict__ rnd, long long n, Box box, int kind,
                                      float D, float vx, float vy, float* __restrict__ X,
                                      float* __restrict__ y) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    float c[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      c[d] = box.lo[d] + (box.hi[d] - box.lo[d]) * rnd[3 * p + d];
      X[3 * p + d] = c[d];
    }
    const float dx = c[1] - 0.5f, dy = c[2] - 0.5f;
    // same operation order as the torch expressions, in float32 like the reference
    const float u = expf(-100.0f * (dx * dx + dy * dy)) * expf(-c[0]);
    float out = u;
    if (kind == 1) {
      const float ut = -u;
      const float ux = -200.0f * dx * u;
      const float uy = -200.0f * dy * u;
      const float uxx = (40000.0f * dx * dx - 400.0f) * u;   // the reference's closed form (sic)
      const float uyy = (40000.0f * dy * dy - 400.0f) * u;
      out = ut + vx * ux + vy * uy - D * (uxx + uyy);
    }
    y[p] = out;
  }
}

}  // namespace qcp

extern "C" int qcp_sample_targets(const float* rnd, long long n, const float* lo_hi, int kind,
                                  double diffusion, double v_x, double v_y, float* X, float* y,
                                  void* stream) {
  using namespace qcp;
  if (!lo_hi || (n > 0 && (!rnd || !X || !y)) || (kind != 0 && kind != 1)) {
    set_error("qcp_sample_targets: bad argument");
    return 1;
  }
  if (n <= 0) return 0;
  Box box;
  for (int d = 0; d < 3; ++d) {
    box.lo[d] = lo_hi[d];
    box.hi[d] = lo_hi[3 + d];
  }
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  sample_targets_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      rnd, n, box, kind, (float)diffusion, (float)v_x, (float)v_y, X, y);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("qcp_sample_targets: launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Loss seed and gradient finishing of the fused train step (reference trainer/diffusion_train.py
// :44-47 MSELoss terms, :81-87 backward / clip_grad_norm_): the MSE value and its cotangent
// 2 w (pred - target) / n in ONE launch per term, and the rank average + norm clip of the flat
// parameter gradient in ONE single-CTA launch, instead of ~25 element-wise torch kernels.
// ---------------------------------------------------------------------------------------------
namespace qcp {

__global__ void mse_seed_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                long long n, float gscale, double inv_n, float* __restrict__ grad,
                                double* __restrict__ loss_slot) {
  __shared__ double red[8];
  double acc = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    const float d = pred[p] - target[p];
    grad[p] = gscale * d;
    acc += (double)d * (double)d;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    atomicAdd(loss_slot, s * inv_n);
  }
}

// flat = [n_grad gradients | n_extra scalars]: everything is scaled by pre_scale (1 / world size
// after a sum all-reduce), then the gradient part is clipped to max_norm like
// torch.nn.utils.clip_grad_norm_ (coef = max_norm / (norm + 1e-6), capped at 1)
__global__ void clip_grads_kernel(float* __restrict__ flat, int n_grad, int n_extra, float pre_scale,
                                  float max_norm) {
  __shared__ double red[32];
  __shared__ float coef_s;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n_grad; i += blockDim.x) {
    const double g = (double)flat[i] * (double)pre_scale;
    acc += g * g;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    const float norm = (float)sqrt(s);
    float coef = max_norm / (norm + 1e-6f);
    coef_s = coef < 1.0f ? coef : 1.0f;
  }
  __syncthreads();
  const float c = coef_s * pre_scale;
  for (int i = threadIdx.x; i < n_grad; i += blockDim.x) flat[i] *= c;
  for (int i = threadIdx.x; i < n_extra; i += blockDim.x) flat[n_grad + i] *= pre_scale;
}

// gradient vector of the adjoint kernels (plan dtype) -> float32 flat buffer of the optimizer, with the
// weighted objective and its three terms behind it: flat = [grads | loss | loss_r | loss_bc | loss_ic]
template <typename T>
__global__ void pack_step_kernel(const T* __restrict__ grads, int n_grad, const double* __restrict__ terms,
                                 double w_r, double w_bc, double w_ic, float* __restrict__ flat) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_grad; i += gridDim.x * blockDim.x)
    flat[i] = (float)grads[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double t0 = terms[0], t1 = terms[1], t2 = terms[2];
    flat[n_grad] = (float)(t0 * w_r + t1 * w_bc + t2 * w_ic);
    flat[n_grad + 1] = (float)t0;
    flat[n_grad + 2] = (float)t1;
    flat[n_grad + 3] = (float)t2;
  }
}

// torch.optim.lr_scheduler.ReduceLROnPlateau.step(metric) on the device, in the same double
// arithmetic as the Python original; state = {best, num_bad_epochs, cooldown_counter, last_epoch,
// recorded, reductions}
__global__ void plateau_step_kernel(const float* __restrict__ metric, double* __restrict__ state,
                                    float* __restrict__ lr, float* __restrict__ history,
                                    long long history_cap, int mode_max, int threshold_abs,
                                    double threshold, double factor, double patience, double cooldown,
                                    double min_lr, double eps) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const double cur = (double)metric[0];
  double best = state[0], bad = state[1], cool = state[2];
  state[3] += 1.0;
  bool better;
  if (!mode_max && !threshold_abs) better = cur < best * (1.0 - threshold);
  else if (!mode_max) better = cur < best - threshold;
  else if (!threshold_abs) better = cur > best * (threshold + 1.0);
  else better = cur > best + threshold;
  if (better) { best = cur; bad = 0.0; } else { bad += 1.0; }
  if (cool > 0.0) { cool -= 1.0; bad = 0.0; }
  if (bad > patience) {
    const double old_lr = (double)lr[0];
    const double scaled = old_lr * factor;
    const double new_lr = scaled > min_lr ? scaled : min_lr;       // Python max(a, b)
    if (old_lr - new_lr > eps) { lr[0] = (float)new_lr; state[5] += 1.0; }
    cool = cooldown;
    bad = 0.0;
  }
  state[0] = best; state[1] = bad; state[2] = cool;
  if (history) {
    const long long k = (long long)state[4];
    if (k < history_cap) history[k] = metric[0];
    state[4] = (double)(k + 1);
  }
}

}  // namespace qcp

extern "C" int qcp_mse_seed(const float* pred, const float* target, long long n, double weight,
                            float* grad, double* loss_slot, void* stream) {
  using namespace qcp;
  if (!loss_slot || (n > 0 && (!pred || !target || !grad))) { set_error("qcp_mse_seed: NULL argument"); return 1; }
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  mse_seed_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, target, n, (float)(2.0 * weight / (double)n), 1.0 / (double)n, grad, loss_slot);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("qcp_mse_seed: launch failed: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

extern "C" int qcp_clip_grads(float* flat, int n_grad, int n_extra, double pre_scale, double max_norm,
                              void* stream) {
  using namespace qcp;
  if (!flat || n_grad < 0 || n_extra < 0) { set_error("qcp_clip_grads: bad argument"); return 1; }
  clip_grads_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(flat, n_grad, n_extra, (float)pre_scale,
                                                                       (float)max_norm);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("qcp_clip_grads: launch failed: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

extern "C" int qcp_pack_step(const void* grads, int dtype, int n_grad, const double* terms, double w_r,
                             double w_bc, double w_ic, float* flat, void* stream) {
  using namespace qcp;
  if ((!grads && n_grad > 0) || !terms || !flat || n_grad < 0 || (dtype != QCP_F32 && dtype != QCP_F64)) {
    set_error("qcp_pack_step: bad argument");
    return 1;
  }
  const int blocks = n_grad > 4096 ? 8 : 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == QCP_F64)
    pack_step_kernel<double><<<blocks, 1024, 0, st>>>(static_cast<const double*>(grads), n_grad, terms, w_r,
                                                      w_bc, w_ic, flat);
  else
    pack_step_kernel<float><<<blocks, 1024, 0, st>>>(static_cast<const float*>(grads), n_grad, terms, w_r,
                                                     w_bc, w_ic, flat);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("qcp_pack_step: launch failed: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

extern "C" int qcp_plateau_step(const float* metric, double* state, float* lr, float* history,
                                long long history_cap, int mode_max, int threshold_abs, double threshold,
                                double factor, long long patience, long long cooldown, double min_lr,
                                double eps, void* stream) {
  using namespace qcp;
  if (!metric || !state || !lr || (history && history_cap <= 0)) {
    set_error("qcp_plateau_step: bad argument");
    return 1;
  }
  plateau_step_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      metric, state, lr, history, history_cap, mode_max, threshold_abs, threshold, factor, (double)patience,
      (double)cooldown, min_lr, eps);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("qcp_plateau_step: launch failed: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}
