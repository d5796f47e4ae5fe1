// optim.cu -- fused Adam over the flat fp32 parameter buffer.
// Replaces torch.optim.Adam(lr=1e-4).step() (/root/reference/nerf/run_nerf_acc.py:206,305-307) with the same
// update formula torch uses (bias corrections folded the same way), one HBM-bound pass: 16 B/param read,
// 12 B/param written.  grad_scale folds the 1/world_size of the data-parallel gradient mean.
#include <stdlib.h>

#include "common.cuh"

namespace {
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                   float step_size, float bc2_sqrt, float gscale, const float* __restrict__ active) {
  if (active && *active == 0.0f) return;                        // iteration without samples: no optimiser step (run_nerf_acc.py:289)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = m[i] + (gi - m[i]) * (1.0f - b1);          // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * b2 + (1.0f - b2) * (gi * gi);         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;             // (exp_avg_sq.sqrt() / sqrt(bias_correction2)).add_(eps)
    p[i] = p[i] - step_size * (mi / denom);                    // param.addcdiv_(exp_avg, denom, value=-step_size)
  }
}

// ---- fused gradient all-reduce + Adam over NVLink peer memory (one process per GPU, gradients in symmetric memory).
// Every rank's backward leaves its 1/global-batch-scaled gradient in its own buffer and publishes a step tag to every peer
// (signal_peers_kernel: system-scope fence, then one flag store per peer over NVLink).  The optimiser kernel of each rank waits
// for all tags, then reads the W gradient buffers directly through their peer mappings, sums them in rank order (so every
// rank computes bit-identical sums and the replicas never drift) and applies the Adam update in the same pass -- no separate
// collective, no reduced gradient ever written.  The buffers alternate with the step parity, so a fast rank cannot overwrite a
// gradient a slow peer is still reading (it would need the slow peer's tag of the NEXT step first).
constexpr int kMaxPeers = 16;
struct PeerPtrs { const float* grad[kMaxPeers]; };
struct PeerFlags { uint32_t* flags[kMaxPeers]; };

__global__ void signal_peers_kernel(PeerFlags peers, int world, int rank, uint32_t tag) {
  __threadfence_system();                                     // this rank's gradient stores are visible system-wide ...
  if ((int)threadIdx.x < world) {
    volatile uint32_t* f = peers.flags[threadIdx.x] + rank;   // ... before its tag lands in peer threadIdx.x's flag array
    *f = tag;
  }
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) adam_allreduce_kernel(float* __restrict__ p, PeerPtrs peers, int world, const uint32_t* my_flags,
                                                             uint32_t tag, float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                             float lr, float b1, float b2, float eps, float step_size, float bc2_sqrt,
                                                             float gscale, int64_t active_index, unsigned long long timeout_ns,
                                                             unsigned long long* __restrict__ wait_stats) {
  __shared__ float s_active;
  if (threadIdx.x == 0) {
    const unsigned long long t_start = global_timer_ns();
    for (int r = 0; r < world; ++r) {
      unsigned spins = 0;
      while ((int32_t)(ld_acquire_sys(my_flags + r) - tag) < 0) {      // wrap-safe "flag >= tag"
        // a peer died: fail loudly instead of hanging (wall-clock bound: rank 0 may legitimately be busy with an evaluation)
        if ((++spins & 1023u) == 0 && global_timer_ns() - t_start > timeout_ns) __trap();
        __nanosleep(64);
      }
    }
    if (wait_stats && blockIdx.x == 0) {          // how long this rank stalled on its slowest peer (bench.py: allreduce_wait_us)
      const unsigned long long w = global_timer_ns() - t_start;
      wait_stats[0] += w;
      wait_stats[1] += 1ull;
      if (w > wait_stats[2]) wait_stats[2] = w;
    }
    float a = 1.0f;
    if (active_index >= 0) {
      a = 0.0f;
      for (int r = 0; r < world; ++r) a += __ldcv(peers.grad[r] + active_index);
    }
    s_active = a;
  }
  __syncthreads();
  if (s_active == 0.0f) return;                                // no rank kept a sample: no optimiser step (run_nerf_acc.py:289)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float g = 0.0f;
    for (int r = 0; r < world; ++r) g += __ldcv(peers.grad[r] + i);   // fixed rank order: identical on every rank
    const float gi = g * gscale;
    const float mi = m[i] + (gi - m[i]) * (1.0f - b1);
    const float vi = v[i] * b2 + (1.0f - b2) * (gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}
}  // namespace

extern "C" int angio_signal_peers(void* const* peer_flags_host, int32_t world, int32_t rank, uint32_t tag, void* stream) {
  ANGIO_REQUIRE(peer_flags_host && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "angio_signal_peers: bad arguments");
  PeerFlags f;
  for (int r = 0; r < world; ++r) f.flags[r] = reinterpret_cast<uint32_t*>(peer_flags_host[r]);
  angio::note_launch("signal_peers_kernel"); signal_peers_kernel<<<1, 32, 0, angio::as_stream(stream)>>>(f, world, rank, tag);
  return angio::finish_launch("angio_signal_peers");
}

extern "C" int angio_adam_step_allreduce(float* params, const void* const* peer_grads_host, int32_t world, const uint32_t* my_flags,
                                         uint32_t tag, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2,
                                         float eps, int32_t step, float grad_scale, int64_t active_index, unsigned long long* wait_stats,
                                         void* stream) {
  ANGIO_REQUIRE(params && peer_grads_host && my_flags && exp_avg && exp_avg_sq && n >= 0 && step >= 1 && world >= 1 && world <= kMaxPeers,
                "angio_adam_step_allreduce: bad arguments");
  if (n == 0) return 0;
  PeerPtrs pp;
  for (int r = 0; r < world; ++r) pp.grad[r] = reinterpret_cast<const float*>(peer_grads_host[r]);
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  static const unsigned long long timeout_ns = [] {
    const char* e = getenv("ANGIO_PEER_TIMEOUT_S");
    const double s = e ? atof(e) : 600.0;
    return (unsigned long long)((s > 0.0 ? s : 600.0) * 1e9);
  }();
  int blocks = angio::blocks_for(n, 256);
  const int cap = angio::sm_count() * 8;
  angio::note_launch("adam_allreduce_kernel"); adam_allreduce_kernel<<<blocks > cap ? cap : blocks, 256, 0, angio::as_stream(stream)>>>(
      params, pp, world, my_flags, tag, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, (float)((double)lr / bc1), (float)sqrt(bc2), grad_scale,
      active_index, timeout_ns, wait_stats);
  return angio::finish_launch("angio_adam_step_allreduce");
}

extern "C" int angio_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                               float beta1, float beta2, float eps, int32_t step, float grad_scale, const float* active, void* stream) {
  ANGIO_REQUIRE(params && grads && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "angio_adam_step: bad arguments");
  if (n == 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  int blocks = angio::blocks_for(n, 256);
  int cap = angio::sm_count() * 8;
  angio::note_launch("adam_kernel"); adam_kernel<<<blocks > cap ? cap : blocks, 256, 0, angio::as_stream(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1,
                                                                                beta2, eps, (float)((double)lr / bc1), (float)sqrt(bc2), grad_scale, active);
  return angio::finish_launch("angio_adam_step");
}
