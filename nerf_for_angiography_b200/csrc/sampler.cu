// sampler.cu -- weighted ray sampling WITHOUT replacement over the whole ray pool, then a uniform shuffle.
// Replaces sample_pixel_rays' `DataFrame.sample(n, weights).sample(frac=1)` (/root/reference/nerf/nerf_helpers.py:137-150;
// 27.5 ms of serial pandas work per iteration at reference scale).
//
// Exponential-race formulation of sampling without replacement (Efraimidis-Spirakis): key_i = E_i / w_i with
// E_i ~ Exp(1); the n smallest keys are the sample.
//   pass 1 (all SMs)   one sweep over the pool: keys from a counter-based hash RNG, only candidates below a threshold tau
//                      (chosen by the caller so that ~n + 8 sigma survive) are appended with warp-aggregated atomics.
//                      HBM-bound: 4 B/ray when weights are given, nothing at all for uniform weights.
//   pass 2 (4 small kernels over the ~n candidates, L2 resident) exact radix select of the n-th smallest key, then the
//                      survivors are ordered by an independent hash of their ray id (bucket sort) -- a uniformly random
//                      permutation that is a pure function of (seed, selected set), i.e. reproducible run to run although
//                      pass 1 appends in arbitrary order.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// E ~ Exp(1) from 48 hash bits.  E = -log(1 - v), v uniform in (0, 1) built so that SMALL v (the keys that can win the
// race) keep 24 significant bits: a 24-bit grid alone would give only ~n distinct winning keys.
__device__ __forceinline__ float exp_variate(uint64_t h) {
  const float hi = (float)(uint32_t)(h >> 40);             // 24 bits
  const float lo = (float)(uint32_t)((h >> 16) & 0xFFFFFFu);
  const float v = (hi + (lo + 0.5f) * (1.0f / 16777216.0f)) * (1.0f / 16777216.0f);
  if (v < 0.015625f) return v * (1.0f + v * (0.5f + v * (1.0f / 3.0f)));   // series of -log(1-v), rel. error < v^3/4
  return -__logf(fmaxf(1.0f - v, 1e-30f));
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
      const float w = weights ? weights[i] : 1.0f;
      key = exp_variate(h) / w;
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

// ---- pass 2: exact selection + shuffle over the m ~ n + 8 sqrt(n) candidates (L2 resident)
//   radix_hist_kernel x3 + select_finish_kernel: radix select of the n-th smallest key -> ctl[1] = its bit pattern, ctl[2] = ties to take
//   mark_kernel     (many CTAs) flag the sample (the flag is the ray's shuffle hash, it replaces the key) + bucket histogram
//   scatter_kernel  (many CTAs) every CTA scans the <= 4096 bucket counts itself, then scatters (hash, id) into its bucket
//   bucket_sort_kernel (warp per bucket) orders each ~32-ray bucket by (hash, id) and writes the final ids
constexpr int kSelThreads = 1024;
constexpr int kMaxBuckets = 4096;

__device__ __forceinline__ uint32_t shuffle_hash(int64_t id, uint64_t seed) {
  return (uint32_t)(mix64((seed ^ 0xD1B54A32D192ED03ull) + 0x9E3779B97F4A7C15ull * (uint64_t)(id + 1)) >> 32);
}

// ctl: [0] candidate counter, [1] threshold key bits, [2] ties to take, [3] ties taken
// Exact radix select of the n-th smallest key (keys are positive floats: their bit patterns order like the values), most
// significant bits first in three passes of 12 + 12 + 8 bits.  Each pass is a multi-CTA histogram into a global table; a CTA of
// the next pass first locates the bin that holds the wanted rank in the previous tables (<= 4096 bins: one block scan), so no
// pass needs a separate "find" launch.  (The single-CTA version of this took 79 us; the winning keys share their leading bits,
// hence the warp-aggregated shared-memory histogram.)
__host__ __device__ constexpr int hist_bits(int pass) { return pass == 2 ? 8 : 12; }
__host__ __device__ constexpr int hist_shift(int pass) { return pass == 0 ? 20 : (pass == 1 ? 8 : 0); }
__host__ __device__ constexpr int hist_off(int pass) { return pass == 0 ? 0 : (pass == 1 ? 4096 : 8192); }
constexpr int kHistWords = 4096 + 4096 + 256;

// all kSelThreads threads: bin and remaining rank such that cum(bin - 1) < need <= cum(bin); nbins <= 4096
__device__ __forceinline__ void find_bin(const uint32_t* __restrict__ hist, int nbins, uint32_t need, uint32_t* s_scan, uint32_t* s_out) {
  const int tid = threadIdx.x;
  uint32_t v[4], sum = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) { const int bin = tid * 4 + k; v[k] = (bin < nbins) ? hist[bin] : 0u; sum += v[k]; }
  uint32_t inc = sum;
  for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if ((tid & 31) >= o) inc += t; }
  if ((tid & 31) == 31) s_scan[tid / 32] = inc;
  __syncthreads();
  if (tid < 32) {
    uint32_t w = s_scan[tid], winc = w;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o); if (tid >= o) winc += t; }
    s_scan[tid] = winc - w;
  }
  __syncthreads();
  uint32_t cum = s_scan[tid / 32] + inc - sum;           // keys in the bins before this thread's four
  if (cum < need && need <= cum + sum) {
    int k = 0;
    for (; k < 4; ++k) { if (cum + v[k] >= need) break; cum += v[k]; }
    s_out[0] = (uint32_t)(tid * 4 + k);
    s_out[1] = need - cum;
  }
  __syncthreads();
}

template <int PASS>
__global__ void __launch_bounds__(kSelThreads) radix_hist_kernel(const float* __restrict__ cand_keys, const int32_t* __restrict__ ctl,
                                                                int32_t capacity, int32_t n, uint32_t* __restrict__ hists) {
  __shared__ uint32_t s_hist[4096];
  __shared__ uint32_t s_scan[32];
  __shared__ uint32_t s_out[2];
  const int tid = threadIdx.x;
  const int m = min(ctl[0], capacity);
  if (m < n) return;
  uint32_t prefix = 0;                                    // bits already fixed by the earlier passes (right-aligned)
  if (PASS >= 1) {
    find_bin(hists + hist_off(0), 1 << hist_bits(0), (uint32_t)n, s_scan, s_out);
    prefix = s_out[0];
    if (PASS >= 2) {
      const uint32_t need1 = s_out[1];
      __syncthreads();
      find_bin(hists + hist_off(1), 1 << hist_bits(1), need1, s_scan, s_out);
      prefix = (prefix << hist_bits(1)) | s_out[0];
    }
  }
  constexpr int nb = 1 << hist_bits(PASS);
  for (int b = tid; b < nb; b += kSelThreads) s_hist[b] = 0;
  __syncthreads();
  const uint32_t* kb = reinterpret_cast<const uint32_t*>(cand_keys);
  for (int i0 = blockIdx.x * kSelThreads; i0 < m; i0 += gridDim.x * kSelThreads) {
    const int i = i0 + tid;
    const uint32_t b = (i < m) ? kb[i] : 0u;
    const bool hit = (i < m) && (PASS == 0 || (b >> (hist_shift(PASS) + hist_bits(PASS))) == prefix);
    const uint32_t digit = hit ? (b >> hist_shift(PASS)) & (uint32_t)(nb - 1) : 0xFFFFFFFFu;
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    if (hit && (tid & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[digit], (uint32_t)__popc(peers));
  }
  __syncthreads();
  for (int b = tid; b < nb; b += kSelThreads)
    if (s_hist[b]) atomicAdd(&hists[hist_off(PASS) + b], s_hist[b]);
}

// one CTA: the three bins -> threshold bit pattern + ties to take; also the status words of the call
constexpr int kTieCap = 1024;

__global__ void __launch_bounds__(kSelThreads) select_finish_kernel(int32_t* __restrict__ ctl, int32_t capacity, int32_t n,
                                                                   uint32_t* __restrict__ hists, int32_t* __restrict__ status,
                                                                   int64_t* __restrict__ ids_out, const float* __restrict__ cand_keys,
                                                                   const int64_t* __restrict__ cand_ids) {
  __shared__ uint32_t s_scan[32];
  __shared__ uint32_t s_out[2];
  __shared__ int64_t s_tie[kTieCap];
  __shared__ uint32_t s_ntie;
  __shared__ long long s_id_thr;
  const int count = ctl[0];
  const int m = min(count, capacity);
  if (threadIdx.x == 0) {
    status[0] = count;
    status[1] = (count > capacity || m < n) ? 1 : 0;        // overflow / too few candidates: the caller re-draws with a larger tau
  }
  if (m < n) {
    // failed draw (fewer candidates than requested): the later kernels return early, so leave VALID ids behind (ray 0) --
    // the gather that follows dereferences them before any host code has looked at status[1]
    for (int i = threadIdx.x; i < n; i += kSelThreads) ids_out[i] = 0;
    return;
  }
  uint32_t prefix = 0, need = (uint32_t)n;
  for (int pass = 0; pass < 3; ++pass) {
    find_bin(hists + hist_off(pass), 1 << hist_bits(pass), need, s_scan, s_out);
    prefix = (prefix << hist_bits(pass)) | s_out[0];
    need = s_out[1];
    __syncthreads();
  }
  // Candidates whose key EQUALS the threshold: `need` of them belong to the draw.  With fp32 keys this is not a measure-zero event
  // (about n * 2^-24 per draw: every few hundred draws at n = 65 536), and the candidate list is in atomic-append order, so the
  // ties are resolved by ray id -- the `need` smallest ids -- which makes the drawn SET a function of the seed alone.
  if (threadIdx.x == 0) { s_ntie = 0; s_id_thr = 0x7FFFFFFFFFFFFFFFll; }
  __syncthreads();
  const uint32_t* kb = reinterpret_cast<const uint32_t*>(cand_keys);
  for (int i = threadIdx.x; i < m; i += kSelThreads)
    if (kb[i] == prefix) {
      const uint32_t p = atomicAdd(&s_ntie, 1u);
      if (p < (uint32_t)kTieCap) s_tie[p] = cand_ids[i];
    }
  __syncthreads();
  const uint32_t nt = s_ntie;
  if (need < nt && nt <= (uint32_t)kTieCap) {
    for (uint32_t t = threadIdx.x; t < nt; t += kSelThreads) {
      const int64_t mine = s_tie[t];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < nt; ++j) rank += s_tie[j] < mine;
      if (rank == need - 1) s_id_thr = mine;
    }
  } else if (need < nt) {
    // thousands of equal keys (degenerate weights): bisect the smallest id threshold with count(tied ids <= threshold) >= need
    long long lo = -1, hi = 0x3FFFFFFFFFFFFFFFll;
    while (hi - lo > 1) {
      const long long mid = lo + (hi - lo) / 2;
      __syncthreads();
      if (threadIdx.x == 0) s_ntie = 0;
      __syncthreads();
      uint32_t c = 0;
      for (int i = threadIdx.x; i < m; i += kSelThreads) c += (kb[i] == prefix && cand_ids[i] <= mid) ? 1u : 0u;
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_ntie, c);
      __syncthreads();
      if (s_ntie >= need) hi = mid; else lo = mid;
    }
    if (threadIdx.x == 0) s_id_thr = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ctl[1] = (int32_t)prefix; ctl[2] = (int32_t)need; ctl[3] = 0;
    *reinterpret_cast<long long*>(hists) = s_id_thr;        // the histograms are consumed: their first words carry the id threshold
  }
}

__device__ __forceinline__ uint32_t bucket_of(uint32_t flag, int log2_buckets) { return log2_buckets ? flag >> (32 - log2_buckets) : 0u; }

__global__ void __launch_bounds__(256) mark_kernel(float* __restrict__ cand_keys, const int64_t* __restrict__ cand_ids, int32_t* __restrict__ ctl,
                                                   int32_t capacity, int32_t n, uint64_t seed, int32_t log2_buckets,
                                                   uint32_t* __restrict__ bcount, const uint32_t* __restrict__ hists) {
  const int m = min(ctl[0], capacity);
  if (m < n) return;
  const uint32_t T = (uint32_t)ctl[1], need_ties = (uint32_t)ctl[2];
  const long long id_thr = *reinterpret_cast<const long long*>(hists);
  uint32_t* kb = reinterpret_cast<uint32_t*>(cand_keys);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const uint32_t b = kb[i];
    bool sel = b < T;
    if (b == T) sel = cand_ids[i] <= id_thr && (uint32_t)atomicAdd(&ctl[3], 1) < need_ties;   // ties at the threshold: smallest ids first
    uint32_t flag = 0xFFFFFFFFu;
    if (sel) {
      flag = shuffle_hash(cand_ids[i], seed);
      if (flag == 0xFFFFFFFFu) flag = 0xFFFFFFFEu;
      atomicAdd(&bcount[bucket_of(flag, log2_buckets)], 1u);
    }
    kb[i] = flag;
  }
}

__global__ void __launch_bounds__(kSelThreads) scatter_kernel(const float* __restrict__ cand_keys, const int64_t* __restrict__ cand_ids,
                                                             const int32_t* __restrict__ ctl, int32_t capacity, int32_t n, int32_t n_buckets,
                                                             int32_t log2_buckets, const uint32_t* __restrict__ bcount, uint32_t* __restrict__ bfill,
                                                             uint32_t* __restrict__ bstart_out, uint32_t* __restrict__ tmp_hash,
                                                             int64_t* __restrict__ tmp_ids) {
  __shared__ uint32_t bstart[kMaxBuckets + 1];
  __shared__ uint32_t scan_tmp[kSelThreads / 32];
  const int tid = threadIdx.x;
  const int m = min(ctl[0], capacity);
  if (m < n) return;
  {   // exclusive scan of the bucket counts (n_buckets <= 4096 = 4 per thread), redone by every CTA
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int b = tid * 4 + k; v[k] = (b < n_buckets) ? bcount[b] : 0u; sum += v[k]; }
    uint32_t inc = sum;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if ((tid & 31) >= o) inc += t; }
    if ((tid & 31) == 31) scan_tmp[tid / 32] = inc;
    __syncthreads();
    if (tid < 32) {
      uint32_t w = scan_tmp[tid], winc = w;
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o); if (tid >= o) winc += t; }
      scan_tmp[tid] = winc - w;
    }
    __syncthreads();
    uint32_t run = scan_tmp[tid / 32] + inc - sum;
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int b = tid * 4 + k; if (b < n_buckets) { bstart[b] = run; run += v[k]; } }
    if (tid == kSelThreads - 1) bstart[n_buckets] = run;
    __syncthreads();
    if (blockIdx.x == 0)
      for (int b = tid; b <= n_buckets; b += kSelThreads) bstart_out[b] = bstart[b];
  }
  const uint32_t* kb = reinterpret_cast<const uint32_t*>(cand_keys);
  for (int i = blockIdx.x * blockDim.x + tid; i < m; i += gridDim.x * blockDim.x) {
    const uint32_t flag = kb[i];
    if (flag != 0xFFFFFFFFu) {
      const uint32_t b = bucket_of(flag, log2_buckets);
      const uint32_t pos = bstart[b] + atomicAdd(&bfill[b], 1u);        // arbitrary order inside a bucket; the sort fixes it
      tmp_hash[pos] = flag;
      tmp_ids[pos] = cand_ids[i];
    }
  }
}

// Buckets hold ~32 rays: the common case (<= 64) keeps them in registers and compares through shuffles.
__global__ void __launch_bounds__(256) bucket_sort_kernel(const int32_t* __restrict__ ctl, int32_t capacity, int32_t n, int32_t n_buckets,
                                                          const uint32_t* __restrict__ bstart, const uint32_t* __restrict__ tmp_hash,
                                                          const int64_t* __restrict__ tmp_ids, int64_t* __restrict__ ids_out) {
  if (min(ctl[0], capacity) < n) return;
  const int lane = threadIdx.x % 32;
  const int b = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (b >= n_buckets) return;
  const uint32_t s0 = bstart[b], cnt = bstart[b + 1] - s0;
  if (cnt == 0) return;
  if (cnt <= 64) {
    const bool v0 = (uint32_t)lane < cnt, v1 = (uint32_t)lane + 32 < cnt;
    const uint32_t h0 = v0 ? tmp_hash[s0 + lane] : 0u, h1 = v1 ? tmp_hash[s0 + lane + 32] : 0u;
    const int64_t i0 = v0 ? tmp_ids[s0 + lane] : 0, i1 = v1 ? tmp_ids[s0 + lane + 32] : 0;
    uint32_t r0 = 0, r1 = 0;
    const int c0 = cnt < 32 ? (int)cnt : 32;
    for (int j = 0; j < c0; ++j) {
      const uint32_t hj = __shfl_sync(0xffffffffu, h0, j);
      const int64_t ij = __shfl_sync(0xffffffffu, i0, j);
      r0 += (hj < h0) || (hj == h0 && ij < i0);
      r1 += (hj < h1) || (hj == h1 && ij < i1);
    }
    for (int j = 32; j < (int)cnt; ++j) {
      const uint32_t hj = __shfl_sync(0xffffffffu, h1, j - 32);
      const int64_t ij = __shfl_sync(0xffffffffu, i1, j - 32);
      r0 += (hj < h0) || (hj == h0 && ij < i0);
      r1 += (hj < h1) || (hj == h1 && ij < i1);
    }
    if (v0) ids_out[s0 + r0] = i0;
    if (v1) ids_out[s0 + r1] = i1;
  } else {
    for (uint32_t e = lane; e < cnt; e += 32) {
      const uint32_t he = tmp_hash[s0 + e];
      const int64_t ie = tmp_ids[s0 + e];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < cnt; ++j) {
        const uint32_t hj = tmp_hash[s0 + j];
        rank += (hj < he) || (hj == he && tmp_ids[s0 + j] < ie);
      }
      ids_out[s0 + rank] = ie;
    }
  }
}

__global__ void __launch_bounds__(256) raygen_flat_kernel(const double* __restrict__ cam2world, const int64_t* __restrict__ ids, int64_t n,
                                                          int img_w, int img_h, double focal, const float* __restrict__ pixels,
                                                          float* __restrict__ rays_o, float* __restrict__ rays_d, float* __restrict__ pix_out) {
  // same arithmetic as raygen_kernel (raygen.cu): float64 in the reference's operation order, one rounding to fp32
  const double half_w = (double)img_w / 2.0, half_h = (double)img_h / 2.0;
  const int64_t hw = (int64_t)img_w * img_h;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t id = ids[i];
    const int v = (int)(id / hw);
    const int rem = (int)(id - (int64_t)v * hw);
    const int y = rem / img_w, x = rem - y * img_w;
    const double* M = cam2world + (int64_t)v * 16;
    const double d0 = __ddiv_rn(__dsub_rn((double)x, half_w), focal);
    const double d1 = -__ddiv_rn(__dsub_rn((double)y, half_h), focal);
    const double d2 = -1.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double s = __dadd_rn(__dadd_rn(__dmul_rn(d0, M[k * 4 + 0]), __dmul_rn(d1, M[k * 4 + 1])), __dmul_rn(d2, M[k * 4 + 2]));
      rays_d[i * 3 + k] = (float)s;
      rays_o[i * 3 + k] = (float)M[k * 4 + 3];
    }
    if (pix_out) pix_out[i] = pixels[id];
  }
}

inline int64_t align256(int64_t b) { return (b + 255) / 256 * 256; }

}  // namespace

extern "C" int angio_sample_candidates(const float* weights, int64_t n_pool, uint64_t seed, float tau, int32_t capacity,
                                       float* cand_keys, int64_t* cand_ids, int32_t* counter, void* stream) {
  ANGIO_REQUIRE(n_pool > 0 && capacity > 0 && cand_keys && cand_ids && counter && tau > 0.0f, "angio_sample_candidates: bad arguments");
  int blocks = angio::blocks_for(n_pool, 256);
  const int cap = angio::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  angio::note_launch("sample_candidates_kernel"); sample_candidates_kernel<<<blocks, 256, 0, angio::as_stream(stream)>>>(weights, n_pool, seed, tau, capacity, cand_keys, cand_ids, counter);
  return angio::finish_launch("angio_sample_candidates");
}

extern "C" int64_t angio_sample_rays_workspace_bytes(int32_t capacity, int64_t n) {
  if (capacity <= 0 || n < 0) return ANGIO_ERR_INVALID_ARG;
  // control words + bucket counters | candidate keys | candidate ids | bucketed hashes | bucketed ids
  return align256(16 + (3 * kMaxBuckets + 1 + kHistWords) * 4) + align256((int64_t)capacity * 4) + align256((int64_t)capacity * 8) + align256(n * 4) +
         align256(n * 8);
}

namespace {

struct SampleWorkspace {
  int32_t* ctl;
  uint32_t *bcount, *bfill, *hists, *bstart, *tmp_hash;
  float* keys;
  int64_t *ids, *tmp_ids;
};

SampleWorkspace carve_workspace(void* workspace, int32_t capacity, int64_t n) {
  SampleWorkspace w;
  char* wb = reinterpret_cast<char*>(workspace);
  const int64_t head = align256(16 + (3 * kMaxBuckets + 1 + kHistWords) * 4);
  w.ctl = reinterpret_cast<int32_t*>(wb);
  w.bcount = reinterpret_cast<uint32_t*>(wb + 16);
  w.bfill = w.bcount + kMaxBuckets;
  w.hists = w.bfill + kMaxBuckets;                                     // zeroed together with the counters
  w.bstart = w.hists + kHistWords;
  wb += head;
  w.keys = reinterpret_cast<float*>(wb); wb += align256((int64_t)capacity * 4);
  w.ids = reinterpret_cast<int64_t*>(wb); wb += align256((int64_t)capacity * 8);
  w.tmp_hash = reinterpret_cast<uint32_t*>(wb); wb += align256(n * 4);
  w.tmp_ids = reinterpret_cast<int64_t*>(wb);
  return w;
}

// candidates in w.keys / w.ids, their count in w.ctl[0] -> the n smallest keys, shuffled, in ids_out
int select_and_shuffle(const SampleWorkspace& w, int32_t capacity, int64_t n, uint64_t seed, int64_t* ids_out, int32_t* status, cudaStream_t st) {
  int n_buckets = 1, log2b = 0;                                        // ~32 rays per shuffle bucket
  while (n_buckets < kMaxBuckets && (int64_t)n_buckets * 32 < n) { n_buckets <<= 1; ++log2b; }
  int sweep_blocks = angio::blocks_for(capacity, 1024);
  if (sweep_blocks > angio::sm_count()) sweep_blocks = angio::sm_count();
  angio::note_launch("radix_hist_kernel<0>"); radix_hist_kernel<0><<<sweep_blocks, kSelThreads, 0, st>>>(w.keys, w.ctl, capacity, (int32_t)n, w.hists);
  angio::note_launch("radix_hist_kernel<1>"); radix_hist_kernel<1><<<sweep_blocks, kSelThreads, 0, st>>>(w.keys, w.ctl, capacity, (int32_t)n, w.hists);
  angio::note_launch("radix_hist_kernel<2>"); radix_hist_kernel<2><<<sweep_blocks, kSelThreads, 0, st>>>(w.keys, w.ctl, capacity, (int32_t)n, w.hists);
  angio::note_launch("select_finish_kernel"); select_finish_kernel<<<1, kSelThreads, 0, st>>>(w.ctl, capacity, (int32_t)n, w.hists, status, ids_out, w.keys, w.ids);
  angio::note_launch("mark_kernel"); mark_kernel<<<sweep_blocks * 4, 256, 0, st>>>(w.keys, w.ids, w.ctl, capacity, (int32_t)n, seed, log2b, w.bcount, w.hists);
  angio::note_launch("scatter_kernel"); scatter_kernel<<<sweep_blocks, kSelThreads, 0, st>>>(w.keys, w.ids, w.ctl, capacity, (int32_t)n, n_buckets, log2b, w.bcount,
                                                                          w.bfill, w.bstart, w.tmp_hash, w.tmp_ids);
  angio::note_launch("bucket_sort_kernel"); bucket_sort_kernel<<<angio::blocks_for(n_buckets, 8), 256, 0, st>>>(w.ctl, capacity, (int32_t)n, n_buckets, w.bstart,
                                                                                        w.tmp_hash, w.tmp_ids, ids_out);
  return 0;
}

}  // namespace

extern "C" int angio_sample_rays(const float* weights, int64_t n_pool, int64_t n, uint64_t seed, float tau, int32_t capacity,
                                 int64_t* ids_out, int32_t* status, void* workspace, int64_t workspace_bytes, void* stream) {
  ANGIO_REQUIRE(n_pool > 0 && n > 0 && n <= n_pool && capacity >= n && tau > 0.0f && ids_out && status && workspace,
                "angio_sample_rays: bad arguments");
  ANGIO_REQUIRE(n <= (int64_t)1 << 24, "angio_sample_rays: at most 2^24 rays per call");
  if (workspace_bytes < angio_sample_rays_workspace_bytes(capacity, n)) {
    angio::set_error("angio_sample_rays: workspace too small");
    return ANGIO_ERR_WORKSPACE;
  }
  cudaStream_t st = angio::as_stream(stream);
  const SampleWorkspace w = carve_workspace(workspace, capacity, n);
  ANGIO_CUDA(cudaMemsetAsync(w.ctl, 0, 16 + (2 * kMaxBuckets + kHistWords) * 4, st));   // counters, bucket counts / fills, radix histograms
  if (int rc = angio_sample_candidates(weights, n_pool, seed, tau, capacity, w.keys, w.ids, w.ctl, stream)) return rc;
  select_and_shuffle(w, capacity, n, seed, ids_out, status, st);
  return angio::finish_launch("angio_sample_rays");
}

extern "C" int angio_sample_select(const float* cand_keys, const int64_t* cand_ids, int32_t m, int64_t n, uint64_t seed, int64_t* ids_out,
                                   int32_t* status, void* workspace, int64_t workspace_bytes, void* stream) {
  ANGIO_REQUIRE(cand_keys && cand_ids && m > 0 && n > 0 && ids_out && status && workspace, "angio_sample_select: bad arguments");
  ANGIO_REQUIRE(n <= (int64_t)1 << 24, "angio_sample_select: at most 2^24 rays per call");
  if (workspace_bytes < angio_sample_rays_workspace_bytes(m, n)) {
    angio::set_error("angio_sample_select: workspace too small");
    return ANGIO_ERR_WORKSPACE;
  }
  cudaStream_t st = angio::as_stream(stream);
  const SampleWorkspace w = carve_workspace(workspace, m, n);
  ANGIO_CUDA(cudaMemsetAsync(w.ctl, 0, 16 + (2 * kMaxBuckets + kHistWords) * 4, st));
  ANGIO_CUDA(cudaMemcpyAsync(w.keys, cand_keys, (size_t)m * 4, cudaMemcpyDeviceToDevice, st));
  ANGIO_CUDA(cudaMemcpyAsync(w.ids, cand_ids, (size_t)m * 8, cudaMemcpyDeviceToDevice, st));
  ANGIO_CUDA(cudaMemcpyAsync(w.ctl, &m, 4, cudaMemcpyHostToDevice, st));               // pageable source: copied before the call returns
  select_and_shuffle(w, m, n, seed, ids_out, status, st);
  return angio::finish_launch("angio_sample_select");
}

extern "C" int angio_raygen_flat(const double* cam2world, const int64_t* ids, int64_t n, int32_t img_w, int32_t img_h, double focal,
                                 const float* pixels, float* rays_o, float* rays_d, float* pix_out, void* stream) {
  ANGIO_REQUIRE(cam2world && ids && rays_o && rays_d, "angio_raygen_flat: null pointer");
  ANGIO_REQUIRE(n >= 0 && img_w > 0 && img_h > 0 && focal != 0.0, "angio_raygen_flat: bad sizes");
  ANGIO_REQUIRE((pix_out == nullptr) || (pixels != nullptr), "angio_raygen_flat: pix_out needs pixels");
  if (n == 0) return 0;
  int blocks = angio::blocks_for(n, 256);
  const int cap = angio::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  angio::note_launch("raygen_flat_kernel"); raygen_flat_kernel<<<blocks, 256, 0, angio::as_stream(stream)>>>(cam2world, ids, n, img_w, img_h, focal, pixels, rays_o, rays_d, pix_out);
  return angio::finish_launch("angio_raygen_flat");
}
