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
