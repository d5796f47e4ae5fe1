// sampler.cu -- weighted ray sampling WITHOUT replacement over the whole ray pool.
// Replaces sample_pixel_rays' `DataFrame.sample(n, weights)` (/root/reference/nerf/nerf_helpers.py:137-150; 27.5 ms of
// serial pandas work per iteration at reference scale).  Exponential-race formulation of sampling without replacement
// (Efraimidis-Spirakis): key_i = -log(u_i) / w_i, the n smallest keys are the sample.  One pass over the pool computes
// the keys from a counter-based hash RNG and keeps only candidates below a threshold tau chosen so that ~n + 8 sigma
// survive; survivors are appended with warp-aggregated atomics.  The caller finishes with a top-n over the few
// survivors.  HBM-bound: 4 B/ray when weights are given, nothing at all for uniform weights.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) sample_candidates_kernel(const float* __restrict__ weights, int64_t n_pool, uint64_t seed,
                                                                float tau, int32_t capacity, float* __restrict__ cand_keys,
                                                                int64_t* __restrict__ cand_ids, int32_t* __restrict__ counter) {
  const int lane = threadIdx.x % 32;
  for (int64_t base = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) - lane; base < n_pool; base += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = base + lane;
    bool take = false;
    float key = 0.0f;
    if (i < n_pool) {
      const uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1));
      const float u = ((float)(uint32_t)(h >> 40) + 0.5f) * (1.0f / 16777216.0f);   // (0, 1)
      const float w = weights ? weights[i] : 1.0f;
      key = -__logf(u) / w;
      take = (w > 0.0f) && (key < tau);
    }
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (m) {
      int slot = 0;
      if (lane == (__ffs(m) - 1)) slot = atomicAdd(counter, __popc(m));
      slot = __shfl_sync(0xffffffffu, slot, __ffs(m) - 1);
      if (take) {
        const int pos = slot + __popc(m & ((1u << lane) - 1u));
        if (pos < capacity) { cand_keys[pos] = key; cand_ids[pos] = i; }
      }
    }
  }
}

}  // namespace

extern "C" int angio_sample_candidates(const float* weights, int64_t n_pool, uint64_t seed, float tau, int32_t capacity,
                                       float* cand_keys, int64_t* cand_ids, int32_t* counter, void* stream) {
  ANGIO_REQUIRE(n_pool > 0 && capacity > 0 && cand_keys && cand_ids && counter && tau > 0.0f, "angio_sample_candidates: bad arguments");
  int blocks = angio::blocks_for(n_pool, 256);
  const int cap = angio::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  angio::note_launch(); sample_candidates_kernel<<<blocks, 256, 0, angio::as_stream(stream)>>>(weights, n_pool, seed, tau, capacity, cand_keys, cand_ids, counter);
  return angio::finish_launch("angio_sample_candidates");
}
