// visibility.cu -- transmittance-based visibility filter and stream compaction of ray samples.
// Replaces nerfacc's render_visibility and the three boolean-mask gathers inside nerfacc.ray_marching
// (run after alpha_fn, /root/reference/nerf/nerf_helpers_acc.py:11-25,29).
//
// One warp per ray.  The transmittance T_{i+1} = T_i * (1 - alpha_i) must be accumulated in sample order to
// stay bit-identical to the serial definition, so the chain is walked with warp shuffles (every lane follows
// the same 32-step dependent chain; lane j latches T_j) while the loads stay coalesced; thousands of rays in
// flight hide the chain latency.  Compaction uses __ballot_sync/__popc prefix sums: the kept samples of a ray
// are written contiguously at new_offsets[ray], no atomics.  HBM traffic: 4 B/sample in + 1 B/sample out for
// the mask; 9 B/sample in + 12 B/kept sample out for the compaction.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) visibility_mask_kernel(const float* __restrict__ alphas,
                                                              const int32_t* __restrict__ offsets, int64_t n_rays,
                                                              float eps, float thre, uint8_t* __restrict__ keep,
                                                              int32_t* __restrict__ kept_counts, const float* __restrict__ t_init,
                                                              const int32_t* __restrict__ base_counts, const float* __restrict__ thre_cap) {
  const int lane = threadIdx.x % 32;
  if (thre_cap) thre = fminf(thre, *thre_cap);    // nerfacc: alpha_thre = min(alpha_thre, mean(grid.occs)), the mean stays on the device
  const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  const int64_t n_warps = (int64_t)gridDim.x * blockDim.x / 32;
  for (int64_t r = warp_global; r < n_rays; r += n_warps) {
    const int beg = offsets[r], end = offsets[r + 1];
    float T = t_init ? t_init[r] : 1.0f;          // tail of a lazily marched ray: transmittance behind its head samples
    int kept = 0;
    for (int base = beg; base < end; base += 32) {
      const int i = base + lane;
      const bool valid = i < end;
      const float a = valid ? alphas[i] : 0.0f;
      float myT = 1.0f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float aj = __shfl_sync(0xffffffffu, a, j);
        if (lane == j) myT = T;
        T = __fmul_rn(T, __fsub_rn(1.0f, aj));  // padded lanes multiply by exactly 1
      }
      bool vis = valid && (myT >= eps);
      if (thre > 0.0f) vis = vis && (a >= thre);
      if (valid) keep[i] = vis ? 1 : 0;
      kept += __popc(__ballot_sync(0xffffffffu, vis));
      if (T < eps) {
        // early ray termination: alpha in [0, 1] makes T non-increasing, so nothing behind this chunk can be visible;
        // the rest of the ray is marked invisible WITHOUT reading its alphas (the two-phase visibility pass never
        // computes them)
        for (int k = base + 32 + lane; k < end; k += 32) keep[k] = 0;
        break;
      }
    }
    if (lane == 0) kept_counts[r] = kept + (base_counts ? base_counts[r] : 0);
  }
}

// Thread-per-ray variant for LONG segments (the tail of lazily marched rays in the steady training regime: ~270 samples per ray,
// nothing terminates early).  The transmittance chain is serial per ray either way; walking it with warp shuffles costs 32
// dependent shuffle + multiply steps per 32 samples (190 us for 17 M samples, latency-bound), whereas one thread per ray runs the
// same chain as plain register arithmetic with its alphas prefetched eight at a time -- the 32 lanes of a warp stream 32
// different segments, each lane consuming its own 128-byte lines out of L1.  Same arithmetic, same order: bit-identical flags.
__global__ void __launch_bounds__(128) visibility_mask_ray_kernel(const float* __restrict__ alphas, const int32_t* __restrict__ offsets,
                                                                  int64_t n_rays, float eps, float thre, uint8_t* __restrict__ keep,
                                                                  int32_t* __restrict__ kept_counts, const float* __restrict__ t_init,
                                                                  const int32_t* __restrict__ base_counts, const float* __restrict__ thre_cap) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  if (thre_cap) thre = fminf(thre, *thre_cap);
  const int beg = offsets[r], end = offsets[r + 1];
  float T = t_init ? t_init[r] : 1.0f;
  int kept = 0;
  int i = beg;
  bool dead = false;
  auto one = [&](float a) -> uint32_t {          // the serial step: flag of this sample, then T *= (1 - alpha)
    bool vis = T >= eps;
    if (thre > 0.0f) vis = vis && (a >= thre);
    kept += vis ? 1 : 0;
    T = __fmul_rn(T, __fsub_rn(1.0f, a));
    return vis ? 1u : 0u;
  };
  // scalar head up to a 16-byte boundary of the alpha array, then 16 samples per round as four float4 loads (a lane's loads and
  // its 4-byte flag stores stay inside its own cache lines: a quarter of the memory transactions of scalar accesses)
  for (; i < end && (i & 3) != 0; ++i) keep[i] = (uint8_t)one(alphas[i]);
  dead = T < eps;
  while (i + 16 <= end && !dead) {
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const float4*>(alphas + i + 4 * k);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t f = one(v[k].x);
      f |= one(v[k].y) << 8;
      f |= one(v[k].z) << 16;
      f |= one(v[k].w) << 24;
      *reinterpret_cast<uint32_t*>(keep + i + 4 * k) = f;
    }
    i += 16;
    // alpha in [0, 1] makes T non-increasing: nothing behind this point can be visible
    dead = T < eps;
  }
  for (; i < end && !dead; ++i) { keep[i] = (uint8_t)one(alphas[i]); dead = T < eps; }
  for (; i < end; ++i) keep[i] = 0;
  kept_counts[r] = kept + (base_counts ? base_counts[r] : 0);
}

__global__ void __launch_bounds__(256) compact_kernel(const uint8_t* __restrict__ keep, const int32_t* __restrict__ offsets,
                                                      const int32_t* __restrict__ new_offsets, int64_t n_rays,
                                                      const float* __restrict__ t_starts, const float* __restrict__ t_ends,
                                                      int64_t capacity, int32_t* __restrict__ ray_idx_out, float* __restrict__ t0_out,
                                                      float* __restrict__ t1_out) {
  const int lane = threadIdx.x % 32;
  const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  const int64_t n_warps = (int64_t)gridDim.x * blockDim.x / 32;
  for (int64_t r = warp_global; r < n_rays; r += n_warps) {
    const int beg = offsets[r], end = offsets[r + 1];
    int dst = new_offsets[r];
    const int dst_end = new_offsets[r + 1];
    if (dst == dst_end) continue;
    for (int base = beg; base < end; base += 32) {
      const int i = base + lane;
      const bool k = (i < end) && keep[i];
      const unsigned m = __ballot_sync(0xffffffffu, k);
      if (k) {
        const int pos = dst + __popc(m & ((1u << lane) - 1u));
        if (pos < capacity) {                     // never write past the caller's arrays
          ray_idx_out[pos] = (int32_t)r;
          t0_out[pos] = t_starts[i];
          t1_out[pos] = t_ends[i];
        }
      }
      dst += __popc(m);
      if (dst == dst_end) break;   // every kept sample of this ray has been written
    }
  }
}

// ---- lazily marched rays (march_head_kernel): visibility of the head samples (ray r: head_cnt[r] <= k0 <= 32 samples from slot
// head_base[r]).  One warp per ray, one chunk of the chain above.  Besides the keep flags and the kept count it leaves the
// transmittance behind the head and whether the ray has to be continued: only if its head used the whole budget (more samples
// may follow) and it is not opaque yet.
__global__ void __launch_bounds__(256) visibility_head_mask_kernel(const float* __restrict__ alphas, const int32_t* __restrict__ head_cnt,
                                                                   const int32_t* __restrict__ head_base, int64_t n_rays, int k0, float eps,
                                                                   float thre, uint8_t* __restrict__ keep, int32_t* __restrict__ kept_counts,
                                                                   float* __restrict__ t_end, uint8_t* __restrict__ alive,
                                                                   const float* __restrict__ thre_cap) {
  const int lane = threadIdx.x % 32;
  if (thre_cap) thre = fminf(thre, *thre_cap);
  const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  const int64_t n_warps = (int64_t)gridDim.x * blockDim.x / 32;
  for (int64_t r = warp_global; r < n_rays; r += n_warps) {
    const int cnt = head_cnt[r], base = head_base[r];
    const bool valid = lane < cnt;
    const float a = valid ? alphas[base + lane] : 0.0f;
    float T = 1.0f, myT = 1.0f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float aj = __shfl_sync(0xffffffffu, a, j);
      if (lane == j) myT = T;
      T = __fmul_rn(T, __fsub_rn(1.0f, aj));
    }
    bool vis = valid && (myT >= eps);
    if (thre > 0.0f) vis = vis && (a >= thre);
    if (valid) keep[base + lane] = vis ? 1 : 0;
    const int kept = __popc(__ballot_sync(0xffffffffu, vis));
    if (lane == 0) {
      kept_counts[r] = kept;
      t_end[r] = T;
      alive[r] = (cnt == k0 && T >= eps) ? 1 : 0;
    }
  }
}

// compaction of a lazily marched batch: kept head samples first, then the kept tail samples (packed by tail_offsets), per ray
__global__ void __launch_bounds__(256) compact_head_tail_kernel(const uint8_t* __restrict__ keep_head, const int32_t* __restrict__ head_cnt,
                                                                const int32_t* __restrict__ head_base, const float* __restrict__ head_t0,
                                                                const float* __restrict__ head_t1,
                                                                const uint8_t* __restrict__ keep_tail, const int32_t* __restrict__ tail_offsets,
                                                                const float* __restrict__ tail_t0, const float* __restrict__ tail_t1,
                                                                const int32_t* __restrict__ new_offsets, int64_t n_rays, int64_t capacity,
                                                                int32_t* __restrict__ ray_idx_out, float* __restrict__ t0_out,
                                                                float* __restrict__ t1_out) {
  const int lane = threadIdx.x % 32;
  const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  const int64_t n_warps = (int64_t)gridDim.x * blockDim.x / 32;
  for (int64_t r = warp_global; r < n_rays; r += n_warps) {
    int dst = new_offsets[r];
    const int dst_end = new_offsets[r + 1];
    if (dst == dst_end) continue;
    {
      const int hb = head_base[r];
      const bool k = lane < head_cnt[r] && keep_head[hb + lane];
      const unsigned m = __ballot_sync(0xffffffffu, k);
      const int pos = dst + __popc(m & ((1u << lane) - 1u));
      if (k && pos < capacity) {
        ray_idx_out[pos] = (int32_t)r;
        t0_out[pos] = head_t0[hb + lane];
        t1_out[pos] = head_t1[hb + lane];
      }
      dst += __popc(m);
    }
    const int beg = tail_offsets[r], end = tail_offsets[r + 1];
    for (int base = beg; base < end && dst < dst_end; base += 32) {
      const int i = base + lane;
      const bool k = (i < end) && keep_tail[i];
      const unsigned m = __ballot_sync(0xffffffffu, k);
      const int pos = dst + __popc(m & ((1u << lane) - 1u));
      if (k && pos < capacity) {
        ray_idx_out[pos] = (int32_t)r;
        t0_out[pos] = tail_t0[i];
        t1_out[pos] = tail_t1[i];
      }
      dst += __popc(m);
    }
  }
}

// ---- two-phase visibility pass (early ray termination).  The reference evaluates alpha_fn on EVERY marched sample and
// filters afterwards (nerf_helpers_acc.py:11-29); a sample behind the point where the transmittance fell below
// early_stop_eps can never be kept, whatever its alpha.  Phase A evaluates the first k0 samples of every ray, phase B the
// remaining samples of the rays that are still alive after them; the kept set is bit-identical to the full evaluation.

// counts[r] = number of samples of ray r in [skip, skip + limit) (limit < 0: no limit); 0 for rays with alive[r] == 0
__global__ void __launch_bounds__(256) segment_counts_kernel(const int32_t* __restrict__ offsets, int64_t n_rays, int skip, int limit,
                                                             const uint8_t* __restrict__ alive, int32_t* __restrict__ counts) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  int c = offsets[r + 1] - offsets[r] - skip;
  if (c < 0) c = 0;
  if (limit >= 0 && c > limit) c = limit;
  if (alive && !alive[r]) c = 0;
  counts[r] = c;
}

// sample_ids[seg_offsets[r] + j] = offsets[r] + skip + j   for j < seg_offsets[r+1] - seg_offsets[r]   (warp per ray)
__global__ void __launch_bounds__(256) segment_ids_kernel(const int32_t* __restrict__ offsets, const int32_t* __restrict__ seg_offsets,
                                                          int64_t n_rays, int skip, int32_t* __restrict__ sample_ids) {
  const int lane = threadIdx.x % 32;
  const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  const int64_t n_warps = (int64_t)gridDim.x * blockDim.x / 32;
  for (int64_t r = warp_global; r < n_rays; r += n_warps) {
    const int dst = seg_offsets[r], m = seg_offsets[r + 1] - dst, src = offsets[r] + skip;
    for (int j = lane; j < m; j += 32) sample_ids[dst + j] = src + j;
  }
}

// alive[r] = ray r has more than k0 samples and its transmittance after the first k0 is still >= eps (thread per ray; the
// product is accumulated in sample order exactly like visibility_mask_kernel)
__global__ void __launch_bounds__(256) visibility_head_kernel(const float* __restrict__ alphas, const int32_t* __restrict__ offsets,
                                                              int64_t n_rays, int k0, float eps, uint8_t* __restrict__ alive) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const int beg = offsets[r], cnt = offsets[r + 1] - beg;
  if (cnt <= k0) { alive[r] = 0; return; }
  float T = 1.0f;
  for (int j = 0; j < k0; ++j) T = __fmul_rn(T, __fsub_rn(1.0f, alphas[beg + j]));
  alive[r] = (T >= eps) ? 1 : 0;
}


int warp_grid(int64_t n_rays) {
  int64_t blocks = (n_rays + 7) / 8;  // 8 warps per 256-thread block, one ray per warp per pass
  int64_t cap = (int64_t)angio::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" int angio_visibility_mask(const float* alphas, const int32_t* offsets, int64_t n_rays, float early_stop_eps,
                                     float alpha_thre, uint8_t* keep, int32_t* kept_counts, const float* t_init,
                                     const int32_t* base_counts, const float* alpha_thre_cap, void* stream) {
  ANGIO_REQUIRE(offsets && kept_counts && n_rays >= 0, "angio_visibility_mask: bad arguments");
  if (n_rays == 0) return 0;
  // Long segments (a resumed march: t_init given): one thread per ray; otherwise the warp-per-ray kernel, whose chunked early
  // exit never reads alphas behind the first chunk that terminates a ray (the two-phase pass leaves those undefined).
  if (t_init != nullptr && (reinterpret_cast<uintptr_t>(alphas) & 15) == 0 && (reinterpret_cast<uintptr_t>(keep) & 3) == 0) {
    angio::note_launch("visibility_mask_ray_kernel");
    visibility_mask_ray_kernel<<<angio::blocks_for(n_rays, 128), 128, 0, angio::as_stream(stream)>>>(alphas, offsets, n_rays, early_stop_eps, alpha_thre,
                                                                                                   keep, kept_counts, t_init, base_counts, alpha_thre_cap);
    return angio::finish_launch("angio_visibility_mask");
  }
  angio::note_launch("visibility_mask_kernel"); visibility_mask_kernel<<<warp_grid(n_rays), 256, 0, angio::as_stream(stream)>>>(alphas, offsets, n_rays, early_stop_eps,
                                                                                 alpha_thre, keep, kept_counts, t_init, base_counts, alpha_thre_cap);
  return angio::finish_launch("angio_visibility_mask");
}

extern "C" int angio_ray_segment_counts(const int32_t* offsets, int64_t n_rays, int32_t skip, int32_t limit, const uint8_t* alive,
                                       int32_t* counts, void* stream) {
  ANGIO_REQUIRE(offsets && counts && n_rays >= 0 && skip >= 0, "angio_ray_segment_counts: bad arguments");
  if (n_rays == 0) return 0;
  angio::note_launch("segment_counts_kernel"); segment_counts_kernel<<<angio::blocks_for(n_rays, 256), 256, 0, angio::as_stream(stream)>>>(offsets, n_rays, skip, limit, alive, counts);
  return angio::finish_launch("angio_ray_segment_counts");
}

extern "C" int angio_ray_segment_ids(const int32_t* offsets, const int32_t* seg_offsets, int64_t n_rays, int32_t skip, int32_t* sample_ids,
                                     void* stream) {
  ANGIO_REQUIRE(offsets && seg_offsets && sample_ids && n_rays >= 0 && skip >= 0, "angio_ray_segment_ids: bad arguments");
  if (n_rays == 0) return 0;
  angio::note_launch("segment_ids_kernel"); segment_ids_kernel<<<warp_grid(n_rays), 256, 0, angio::as_stream(stream)>>>(offsets, seg_offsets, n_rays, skip, sample_ids);
  return angio::finish_launch("angio_ray_segment_ids");
}

extern "C" int angio_visibility_head(const float* alphas, const int32_t* offsets, int64_t n_rays, int32_t k0, float early_stop_eps,
                                     uint8_t* alive, void* stream) {
  ANGIO_REQUIRE(alphas && offsets && alive && n_rays >= 0 && k0 > 0, "angio_visibility_head: bad arguments");
  if (n_rays == 0) return 0;
  angio::note_launch("visibility_head_kernel"); visibility_head_kernel<<<angio::blocks_for(n_rays, 256), 256, 0, angio::as_stream(stream)>>>(alphas, offsets, n_rays, k0, early_stop_eps, alive);
  return angio::finish_launch("angio_visibility_head");
}

extern "C" int angio_visibility_head_mask(const float* alphas, const int32_t* head_cnt, const int32_t* head_base, int64_t n_rays, int32_t k0,
                                          float early_stop_eps, float alpha_thre, uint8_t* keep, int32_t* kept_counts, float* t_end,
                                          uint8_t* alive, const float* alpha_thre_cap, void* stream) {
  ANGIO_REQUIRE(alphas && head_cnt && head_base && keep && kept_counts && t_end && alive && n_rays >= 0 && k0 >= 1 && k0 <= 32,
                "angio_visibility_head_mask: bad arguments");
  if (n_rays == 0) return 0;
  angio::note_launch("visibility_head_mask_kernel"); visibility_head_mask_kernel<<<warp_grid(n_rays), 256, 0, angio::as_stream(stream)>>>(
      alphas, head_cnt, head_base, n_rays, k0, early_stop_eps, alpha_thre, keep, kept_counts, t_end, alive, alpha_thre_cap);
  return angio::finish_launch("angio_visibility_head_mask");
}

extern "C" int angio_compact_head_tail(const uint8_t* keep_head, const int32_t* head_cnt, const int32_t* head_base, const float* head_t0,
                                       const float* head_t1, const uint8_t* keep_tail, const int32_t* tail_offsets, const float* tail_t0,
                                       const float* tail_t1, const int32_t* new_offsets, int64_t n_rays, int64_t capacity,
                                       int32_t* ray_idx_out, float* t_starts_out, float* t_ends_out, void* stream) {
  ANGIO_REQUIRE(keep_head && head_cnt && head_base && head_t0 && head_t1 && tail_offsets && new_offsets && ray_idx_out && t_starts_out &&
                    t_ends_out && n_rays >= 0,
                "angio_compact_head_tail: bad arguments");
  if (n_rays == 0) return 0;
  angio::note_launch("compact_head_tail_kernel"); compact_head_tail_kernel<<<warp_grid(n_rays), 256, 0, angio::as_stream(stream)>>>(
      keep_head, head_cnt, head_base, head_t0, head_t1, keep_tail, tail_offsets, tail_t0, tail_t1, new_offsets, n_rays,
      capacity > 0 ? capacity : INT64_MAX, ray_idx_out, t_starts_out, t_ends_out);
  return angio::finish_launch("angio_compact_head_tail");
}

extern "C" int angio_compact_samples(const uint8_t* keep, const int32_t* offsets, const int32_t* new_offsets, int64_t n_rays,
                                     const float* t_starts, const float* t_ends, int64_t capacity, int32_t* ray_idx_out,
                                     float* t_starts_out, float* t_ends_out, void* stream) {
  ANGIO_REQUIRE(offsets && new_offsets && n_rays >= 0, "angio_compact_samples: bad arguments");
  if (n_rays == 0) return 0;
  angio::note_launch("compact_kernel"); compact_kernel<<<warp_grid(n_rays), 256, 0, angio::as_stream(stream)>>>(keep, offsets, new_offsets, n_rays, t_starts, t_ends,
                                                                         capacity > 0 ? capacity : INT64_MAX, ray_idx_out, t_starts_out, t_ends_out);
  return angio::finish_launch("angio_compact_samples");
}
