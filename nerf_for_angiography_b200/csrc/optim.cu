// optim.cu -- fused Adam over the flat fp32 parameter buffer.
// Replaces torch.optim.Adam(lr=1e-4).step() (/root/reference/nerf/run_nerf_acc.py:206,305-307) with the same
// update formula torch uses (bias corrections folded the same way), one HBM-bound pass: 16 B/param read,
// 12 B/param written.  grad_scale folds the 1/world_size of the data-parallel gradient mean.
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
}  // namespace

extern "C" int angio_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                               float beta1, float beta2, float eps, int32_t step, float grad_scale, const float* active, void* stream) {
  ANGIO_REQUIRE(params && grads && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "angio_adam_step: bad arguments");
  if (n == 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  int blocks = angio::blocks_for(n, 256);
  int cap = angio::sm_count() * 8;
  angio::note_launch(); adam_kernel<<<blocks > cap ? cap : blocks, 256, 0, angio::as_stream(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1,
                                                                                beta2, eps, (float)((double)lr / bc1), (float)sqrt(bc2), grad_scale, active);
  return angio::finish_launch("angio_adam_step");
}
