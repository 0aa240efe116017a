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
