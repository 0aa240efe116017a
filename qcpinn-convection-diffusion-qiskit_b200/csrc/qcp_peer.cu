// qcpinn_b200 -- data-parallel gradient exchange fused with the clip, over NVLink peer memory.
//
// The reference has no distributed code; the data-parallel step of this framework (SURVEY.md section
// 8e) averages a ~3 KB flat gradient across the ranks of one NVSwitch node and then clips it
// (reference trainer/diffusion_train.py:85, clip_grad_norm_).  A 3 KB all-reduce is pure latency, so
// instead of an NCCL launch followed by the clip kernel, ONE single-CTA kernel per rank
//   1. stores its vector into slot [parity][rank] of EVERY rank's symmetric buffer (P2P stores),
//   2. releases a sequence-numbered flag on every rank and acquires the flags of all ranks,
//   3. sums the `world` slots of its own buffer in rank order (every rank adds the same numbers in
//      the same order: the result is bit-identical on all ranks and independent of timing),
//   4. scales by 1 / world, computes the gradient norm and applies the clip coefficient in place.
// Two slot sets (step parity) make the exchange race free without a second barrier: a rank can
// only start writing step k + 2 after every peer has SENT its k + 1 flag, i.e. after that peer has
// finished reading step k.  The wait is bounded (clock64): a peer that never arrives sets *error and
// the kernel returns instead of hanging the GPU.
#include "qcp_common.cuh"

namespace qcp {

constexpr int kMaxPeers = 16;

struct PeerBufs {
  float* p[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys(unsigned* addr, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* addr) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}

__global__ void __launch_bounds__(1024)
peer_allreduce_clip_kernel(float* __restrict__ flat, int n_grad, int n_extra, PeerBufs peers, int rank,
                           int world, int n_pad, unsigned* __restrict__ seq_ctr, float max_norm,
                           int* __restrict__ error, long long timeout_cycles) {
  __shared__ unsigned seq_s;
  __shared__ double red[32];
  __shared__ float coef_s;
  const int n = n_grad + n_extra;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) {
    seq_s = *seq_ctr + 1u;
    *seq_ctr = seq_s;
  }
  __syncthreads();
  const unsigned seq = seq_s;
  const int par = (int)(seq & 1u);
  const size_t flag_off = (size_t)2 * world * n_pad;          // flags follow the slots (u32 each)

  // 1. my vector into slot [par][rank] of every rank (peer stores over NVLink; self included)
  for (int idx = tid; idx < world * n; idx += nt) {
    const int p = idx / n, i = idx - p * n;
    peers.p[p][((size_t)par * world + rank) * n_pad + i] = flat[i];
  }
  __threadfence_system();
  __syncthreads();

  // 2. flag exchange: one thread per peer
  if (tid < world) {
    unsigned* theirs = reinterpret_cast<unsigned*>(peers.p[tid] + flag_off) + par * world + rank;
    st_release_sys(theirs, seq);
    const unsigned* mine = reinterpret_cast<const unsigned*>(peers.p[rank] + flag_off) + par * world + tid;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) != seq) {
      if (clock64() - t0 > timeout_cycles) {
        atomicExch(error, 1 + tid);
        break;
      }
    }
  }
  __syncthreads();

  // 3. + 4. ordered sum, average, clip (torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6) <= 1)
  const float* own = peers.p[rank] + (size_t)par * world * n_pad;
  const float inv_world = 1.0f / (float)world;
  double acc = 0.0;
  for (int i = tid; i < n; i += nt) {
    float s = 0.0f;
    for (int r = 0; r < world; ++r) s += __ldcg(own + (size_t)r * n_pad + i);
    s *= inv_world;
    flat[i] = s;
    if (i < n_grad) acc += (double)s * (double)s;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) red[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < (nt >> 5); ++w) s += red[w];
    const float norm = (float)sqrt(s);
    const float coef = max_norm / (norm + 1e-6f);
    coef_s = coef < 1.0f ? coef : 1.0f;
  }
  __syncthreads();
  const float c = coef_s;
  for (int i = tid; i < n_grad; i += nt) flat[i] *= c;
}

}  // namespace qcp

extern "C" long long qcp_peer_allreduce_floats(int n_values, int world) {
  if (n_values <= 0 || world <= 0 || world > qcp::kMaxPeers) return -1;
  const long long n_pad = (n_values + 31) / 32 * 32;
  return 2 * world * n_pad + 2 * world + 32;
}

extern "C" int qcp_peer_allreduce_clip(float* flat, int n_grad, int n_extra, const void* const* peer_bufs,
                                       int rank, int world, unsigned int* seq, double max_norm, int* error,
                                       double timeout_s, void* stream) {
  using namespace qcp;
  if (!flat || !peer_bufs || !seq || !error || n_grad < 0 || n_extra < 0 || n_grad + n_extra <= 0 ||
      world < 1 || world > kMaxPeers || rank < 0 || rank >= world) {
    set_error("qcp_peer_allreduce_clip: bad argument");
    return 1;
  }
  PeerBufs pb{};
  for (int r = 0; r < world; ++r) {
    if (!peer_bufs[r]) { set_error("qcp_peer_allreduce_clip: NULL peer buffer %d", r); return 1; }
    pb.p[r] = static_cast<float*>(const_cast<void*>(peer_bufs[r]));
  }
  const int n_pad = (n_grad + n_extra + 31) / 32 * 32;
  int dev = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const long long cycles = (long long)(timeout_s * 1e3 * (khz > 0 ? khz : 1500000));
  peer_allreduce_clip_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      flat, n_grad, n_extra, pb, rank, world, n_pad, seq, (float)max_norm, error, cycles);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("qcp_peer_allreduce_clip: launch failed: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}
