// march.cu -- occupancy-grid ray marching and sample compaction.
// Replaces nerfacc.ray_marching's slab test and two-pass _C.ray_marching kernel, reached from
// acc_ray_marching (/root/reference/nerf/nerf_helpers_acc.py:29).  The per-ray arithmetic mirrors
// oracle/march_ref.c op for op: every fp32 operation individually rounded (__fadd_rn/__fmul_rn/__fdiv_rn, so
// nvcc cannot contract into FMA), IEEE division, truncating float->int, NaN-dropping fminf/fmaxf.
//
// B200 mapping: marching is a data-dependent serial walk per ray (t0 <- t1 <- t0+dt must be accumulated in
// order to stay bit-exact), so parallelism comes from rays: one thread per ray, a warp marches 32 rays in
// lockstep.  The write pass stages each ray's samples in shared memory and the whole warp flushes one ray
// segment at a time, so global stores are contiguous 128-byte runs instead of 32 scattered 4-byte stores.
// The 128^3 byte grid (2 MB) stays L2-resident; output is 12 B/sample (int32 ray id + t0 + t1).
#include <stdlib.h>

#include "common.cuh"

namespace {

using angio::Roi;

struct MarchParams {
  Roi roi;
  float ext[3];  // roi.hi - roi.lo (rounded once, as the reference recomputes it identically each time)
  float resf;
  int res;
  float dt;
};

__device__ __forceinline__ void ray_aabb(const float o[3], const float d[3], const float* aabb, float& near_, float& far_) {
  float tmin = __fdiv_rn(__fsub_rn(aabb[0], o[0]), d[0]);
  float tmax = __fdiv_rn(__fsub_rn(aabb[3], o[0]), d[0]);
  if (tmin > tmax) { float t = tmin; tmin = tmax; tmax = t; }
  float tymin = __fdiv_rn(__fsub_rn(aabb[1], o[1]), d[1]);
  float tymax = __fdiv_rn(__fsub_rn(aabb[4], o[1]), d[1]);
  if (tymin > tymax) { float t = tymin; tymin = tymax; tymax = t; }
  if (tmin > tymax || tymin > tmax) { near_ = 1e10f; far_ = 1e10f; return; }
  if (tymin > tmin) tmin = tymin;
  if (tymax < tmax) tmax = tymax;
  float tzmin = __fdiv_rn(__fsub_rn(aabb[2], o[2]), d[2]);
  float tzmax = __fdiv_rn(__fsub_rn(aabb[5], o[2]), d[2]);
  if (tzmin > tzmax) { float t = tzmin; tzmin = tzmax; tzmax = t; }
  if (tmin > tzmax || tzmin > tmax) { near_ = 1e10f; far_ = 1e10f; return; }
  if (tzmin > tmin) tmin = tzmin;
  if (tzmax < tmax) tmax = tzmax;
  near_ = tmin;
  far_ = tmax;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ bool occupied_at(float x, float y, float z, const MarchParams& p, const uint8_t* __restrict__ binary) {
  if (x < p.roi.lo[0] || x > p.roi.hi[0] || y < p.roi.lo[1] || y > p.roi.hi[1] || z < p.roi.lo[2] || z > p.roi.hi[2])
    return false;
  const float ux = __fdiv_rn(__fsub_rn(x, p.roi.lo[0]), p.ext[0]);
  const float uy = __fdiv_rn(__fsub_rn(y, p.roi.lo[1]), p.ext[1]);
  const float uz = __fdiv_rn(__fsub_rn(z, p.roi.lo[2]), p.ext[2]);
  const int ix = clampi(__float2int_rz(__fmul_rn(ux, p.resf)), 0, p.res - 1);
  const int iy = clampi(__float2int_rz(__fmul_rn(uy, p.resf)), 0, p.res - 1);
  const int iz = clampi(__float2int_rz(__fmul_rn(uz, p.resf)), 0, p.res - 1);
  return __ldg(binary + ((int64_t)ix * p.res + iy) * p.res + iz) != 0;
}

__device__ __forceinline__ float axis_dist(float pos, float d, float inv_d, float lo, float ext, float resf) {
  const float u = __fmul_rn(__fdiv_rn(__fsub_rn(pos, lo), ext), resf);
  const float s = copysignf(1.0f, d);
  const float f = floorf(__fadd_rn(__fadd_rn(u, 0.5f), __fmul_rn(0.5f, s)));
  return __fmul_rn(__fdiv_rn(__fmul_rn(__fsub_rn(f, u), inv_d), resf), ext);
}

// Per-ray marching state machine; `advance` runs until `budget` more samples were emitted or the ray ends.
struct Marcher {
  float o[3], d[3], inv[3];
  float t0, t1, tm, tmax;

  __device__ __forceinline__ void init(const float* ro, const float* rd, float tmin_, float tmax_, float dt) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { o[k] = ro[k]; d[k] = rd[k]; inv[k] = __fdiv_rn(1.0f, rd[k]); }
    tmax = tmax_;
    t0 = tmin_;
    t1 = __fadd_rn(t0, dt);
    tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
  }
  __device__ __forceinline__ bool done() const { return !(tm < tmax); }

  // emits at most `budget` samples through emit(j, t0, t1); on_skip() is called whenever empty space is skipped (the next
  // sample then starts a new run: its t0 no longer equals the previous t1).  Returns the number emitted.
  template <class Emit, class Skip>
  __device__ __forceinline__ int advance(const MarchParams& p, const uint8_t* __restrict__ binary, int budget, Emit emit, Skip on_skip) {
    int j = 0;
    while (tm < tmax && j < budget) {
      const float x = __fadd_rn(o[0], __fmul_rn(tm, d[0]));
      const float y = __fadd_rn(o[1], __fmul_rn(tm, d[1]));
      const float z = __fadd_rn(o[2], __fmul_rn(tm, d[2]));
      if (occupied_at(x, y, z, p, binary)) {
        emit(j, t0, t1);
        ++j;
        t0 = t1;
        t1 = __fadd_rn(t0, p.dt);
        tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
      } else {
        const float tx = axis_dist(x, d[0], inv[0], p.roi.lo[0], p.ext[0], p.resf);
        const float ty = axis_dist(y, d[1], inv[1], p.roi.lo[1], p.ext[1], p.resf);
        const float tz = axis_dist(z, d[2], inv[2], p.roi.lo[2], p.ext[2], p.resf);
        const float t = fmaxf(fminf(fminf(tx, ty), tz), 0.0f);
        const float target = __fadd_rn(tm, t);
        float _t = tm;
        do { _t = __fadd_rn(_t, p.dt); } while (_t < target);
        tm = _t;
        const float h = __fmul_rn(p.dt, 0.5f);
        t0 = __fsub_rn(tm, h);
        t1 = __fadd_rn(tm, h);
        on_skip();
      }
    }
    return j;
  }
};

// ---- closed form of the marcher's t-chain.  Inside a run the marcher walks t0 <- t1, t1 <- fl(t0 + dt) serially.  When every
// value of a ray's chain lies in ONE binade [2^e, 2^(e+1)) all of them are multiples of u = 2^(e-23), and fl(x + dt) = x + r*u with
// r = round(dt / u) the SAME for every x (the rounding could only depend on x if dt / u ended in exactly .5, which is excluded
// below).  Sample k >= 1 of a run that starts with (t0_s, t1_s) is then (t1_s + (k-1) D, t1_s + k D) with D = r*u -- products and
// sums of small integers times u, exact in fp32 -- so the write pass can compute every sample independently instead of replaying
// the chain one sample at a time per thread.  Rays outside the guard (other geometry) keep the serial replay: bit-identical
// either way.
__device__ __forceinline__ bool linear_chain(float t_lo, float t_hi, float dt, float* delta) {
  const uint32_t a = __float_as_uint(t_lo), b = __float_as_uint(t_hi);
  const uint32_t e = a >> 23;                                       // sign + exponent
  if (e != (b >> 23) || e == 0u || e >= 255u) return false;          // positive, normal, same binade
  if (e < 24u) return false;
  const float u = __uint_as_float((e - 23u) << 23);                  // ulp of the binade
  const float s2 = __fmul_rn(__fdiv_rn(dt, u), 2.0f);                // 2 dt / u: scaling by powers of two is exact
  if (!(s2 < 8388608.0f)) return false;                              // dt must be small against the binade (else s2 is not exact)
  if (s2 == floorf(s2) && fmodf(s2, 2.0f) == 1.0f) return false;     // dt / u ends in .5: round-to-even would depend on x
  *delta = __fsub_rn(__fadd_rn(t_lo, dt), t_lo);                     // = r * u, exact
  return *delta > 0.0f;
}

// all values a ray's chain can take: run starts >= t_min - dt/2 (after a skip t0 = tm - dt/2), last t1 <= t_max + 2 dt
__device__ __forceinline__ bool ray_chain_is_linear(float t_min, float t_max, float dt, float* delta) {
  return linear_chain(__fsub_rn(t_min, dt), __fadd_rn(t_max, __fmul_rn(2.0f, dt)), dt, delta);
}

struct Aabb6 { float v[6]; };
constexpr int kMaxRuns = 8;   // recorded runs per ray (count pass -> write pass)

__global__ void __launch_bounds__(128) march_count_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                          int64_t n_rays, Aabb6 aabb, MarchParams p,
                                                          const uint8_t* __restrict__ binary, float near_plane,
                                                          float far_plane, float* __restrict__ t_min,
                                                          float* __restrict__ t_max, int32_t* __restrict__ counts,
                                                          float* __restrict__ run_t0, float* __restrict__ run_t1,
                                                          int32_t* __restrict__ run_n, int32_t* __restrict__ n_runs,
                                                          const uint8_t* __restrict__ resume_alive, int skip_linear) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_rays) return;
  float o[3] = {rays_o[i * 3], rays_o[i * 3 + 1], rays_o[i * 3 + 2]};
  float d[3] = {rays_d[i * 3], rays_d[i * 3 + 1], rays_d[i * 3 + 2]};
  float a, b;
  if (resume_alive) {
    // continuation of a head march (march_head_kernel): t_min / t_max are INPUTS -- the head stopped right after emitting a
    // sample, where the marcher's state (t0, t0 + dt, their midpoint) is exactly what init() builds from t_min = that t0
    if (!resume_alive[i]) { counts[i] = 0; if (n_runs) n_runs[i] = 0; return; }
    a = t_min[i];
    b = t_max[i];
  } else {
    ray_aabb(o, d, aabb.v, a, b);
    a = a < near_plane ? near_plane : a;  // torch.clamp(t_min, min=near_plane)
    b = b > far_plane ? far_plane : b;    // torch.clamp(t_max, max=far_plane)
    t_min[i] = a;
    t_max[i] = b;
  }
  if (skip_linear) {                      // counted by march_count_warp_kernel (one warp per ray, closed-form t-chain)
    float delta;
    if (ray_chain_is_linear(a, b, p.dt, &delta)) return;
  }
  Marcher m;
  m.init(o, d, a, b, p.dt);
  // Runs: maximal stretches of consecutive samples (t0 of a sample = t1 of the one before).  The first kMaxRuns of a ray are
  // recorded (first sample's t0 and t1, length) so the write pass replays the fp32 t-chain without touching the grid again; a ray with more
  // runs is flagged (n_runs = -1) and re-marched there.
  int total = 0, nr = 0, cur = 0;
  bool open = false;
  while (!m.done())
    total += m.advance(p, binary, 1 << 30,
                       [&](int, float t0, float t1) {
                         if (!open) {
                           open = true;
                           cur = 0;
                           // the first sample of a run carries its own (t0, t1): t_min / t_min + dt at the ray start,
                           // tm -+ dt/2 after a skip; from then on t0 = previous t1, t1 = t0 + dt
                           if (run_t0 && nr < kMaxRuns) { run_t0[(int64_t)nr * n_rays + i] = t0; run_t1[(int64_t)nr * n_rays + i] = t1; }
                         }
                         ++cur;
                       },
                       [&]() {
                         if (open) {
                           if (run_n && nr < kMaxRuns) run_n[(int64_t)nr * n_rays + i] = cur;
                           ++nr;
                           open = false;
                         }
                       });
  if (open) {
    if (run_n && nr < kMaxRuns) run_n[(int64_t)nr * n_rays + i] = cur;
    ++nr;
  }
  counts[i] = total;
  if (n_runs) n_runs[i] = nr <= kMaxRuns ? nr : -1;
}

// Count pass, one WARP per ray, for rays whose t-chain is linear (linear_chain): the 32 lanes test 32 CONSECUTIVE candidate samples
// at once -- candidate k of a stretch that starts with (T0, T1) is (T1 + (k-1) D, T1 + k D), its midpoint and grid lookup are
// independent of the others -- and a ballot finds the first candidate that is empty or behind t_max.  Empty space is skipped with
// the same closed form: the serial marcher adds dt to the midpoint until it reaches the next voxel boundary, i.e. j =
// ceil((target - tm) / D) steps in exact integer arithmetic on multiples of the binade's ulp.  Emits exactly the counts and the
// run table of march_count_kernel (which skips the rays counted here); 65 536 serial rays are only 14 warps per SM.
__global__ void __launch_bounds__(256) march_count_warp_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                               int64_t n_rays, MarchParams p, const uint8_t* __restrict__ binary,
                                                               const float* __restrict__ t_min, const float* __restrict__ t_max,
                                                               int32_t* __restrict__ counts, float* __restrict__ run_t0,
                                                               float* __restrict__ run_t1, int32_t* __restrict__ run_n,
                                                               int32_t* __restrict__ n_runs, const uint8_t* __restrict__ resume_alive) {
  const int lane = threadIdx.x % 32;
  const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  const int64_t n_warps = (int64_t)gridDim.x * blockDim.x / 32;
  for (int64_t i = warp_global; i < n_rays; i += n_warps) {
    if (resume_alive && !resume_alive[i]) continue;         // march_count_kernel writes the zero count
    const float o[3] = {rays_o[i * 3], rays_o[i * 3 + 1], rays_o[i * 3 + 2]};
    const float d[3] = {rays_d[i * 3], rays_d[i * 3 + 1], rays_d[i * 3 + 2]};
    const float ta = t_min[i], tb = t_max[i];              // stored by march_count_kernel (launched before, same stream)
    float D;
    if (!ray_chain_is_linear(ta, tb, p.dt, &D)) continue;   // left to the serial kernel
    const float u = __uint_as_float(((__float_as_uint(ta) >> 23) - 23u) << 23);      // ulp of the ray's binade
    const int Di = __float2int_rn(__fdiv_rn(D, u));         // D / u, exact
    float inv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) inv[k] = __fdiv_rn(1.0f, d[k]);
    const float h = __fmul_rn(p.dt, 0.5f);
    float T0 = ta, T1 = __fadd_rn(ta, p.dt);                // the stretch's first candidate and its midpoint
    float TM = __fmul_rn(__fadd_rn(T0, T1), 0.5f);
    int total = 0, nr = 0, cur = 0;
    bool open = false;
    while (true) {
      // candidate `lane` of the stretch
      float c0 = T0, c1 = T1;
      if (lane > 0) {
        c0 = __fadd_rn(T1, __fmul_rn((float)(lane - 1), D));
        c1 = __fadd_rn(T1, __fmul_rn((float)lane, D));
      }
      const float tm = lane > 0 ? __fmul_rn(__fadd_rn(c0, c1), 0.5f) : TM;
      const bool active = tm < tb;
      bool occ = false;
      if (active) {
        const float x = __fadd_rn(o[0], __fmul_rn(tm, d[0]));
        const float y = __fadd_rn(o[1], __fmul_rn(tm, d[1]));
        const float z = __fadd_rn(o[2], __fmul_rn(tm, d[2]));
        occ = occupied_at(x, y, z, p, binary);
      }
      const unsigned stop = __ballot_sync(0xffffffffu, !occ);
      const int f = stop ? __ffs(stop) - 1 : 32;            // candidates 0 .. f-1 are samples
      if (f > 0) {
        if (!open) {
          open = true;
          cur = 0;
          if (lane == 0 && run_t0 && nr < kMaxRuns) { run_t0[(int64_t)nr * n_rays + i] = T0; run_t1[(int64_t)nr * n_rays + i] = T1; }
        }
        cur += f;
        total += f;
      }
      if (f == 32) {                                        // the stretch goes on: candidate 32 becomes candidate 0
        T0 = __fadd_rn(T1, __fmul_rn(31.0f, D));
        T1 = __fadd_rn(T1, __fmul_rn(32.0f, D));
        TM = __fmul_rn(__fadd_rn(T0, T1), 0.5f);
        continue;
      }
      const float tm_f = __shfl_sync(0xffffffffu, tm, f);
      if (!(tm_f < tb)) break;                              // candidate f lies behind t_max: the ray is done
      // candidate f is empty: skip to the next voxel boundary in steps of dt (Marcher::advance, else-branch)
      const float x = __fadd_rn(o[0], __fmul_rn(tm_f, d[0]));
      const float y = __fadd_rn(o[1], __fmul_rn(tm_f, d[1]));
      const float z = __fadd_rn(o[2], __fmul_rn(tm_f, d[2]));
      const float tx = axis_dist(x, d[0], inv[0], p.roi.lo[0], p.ext[0], p.resf);
      const float ty = axis_dist(y, d[1], inv[1], p.roi.lo[1], p.ext[1], p.resf);
      const float tz = axis_dist(z, d[2], inv[2], p.roi.lo[2], p.ext[2], p.resf);
      const float t = fmaxf(fminf(fminf(tx, ty), tz), 0.0f);
      const float target = __fadd_rn(tm_f, t);
      if (open) {
        if (lane == 0 && run_n && nr < kMaxRuns) run_n[(int64_t)nr * n_rays + i] = cur;
        ++nr;
        open = false;
      }
      // serial: _t = tm; do { _t += dt } while (_t < target).  The loop ends at the first _t >= target; every _t >= t_max ends the
      // ray whatever its exact value, so a target at / behind t_max (or not finite) needs no arithmetic at all.
      if (!(target < tb)) break;
      const int diff = __float2int_rn(__fdiv_rn(__fsub_rn(target, tm_f), u));        // (target - tm) / u, exact: same binade
      int j = (diff + Di - 1) / Di;
      if (j < 1) j = 1;
      const float tm_new = __fadd_rn(tm_f, __fmul_rn((float)j, D));
      TM = tm_new;                                          // after a skip the midpoint is the walked value itself
      T0 = __fsub_rn(tm_new, h);
      T1 = __fadd_rn(tm_new, h);
    }
    if (open) {
      if (lane == 0 && run_n && nr < kMaxRuns) run_n[(int64_t)nr * n_rays + i] = cur;
      ++nr;
    }
    if (lane == 0) {
      counts[i] = total;
      if (n_runs) n_runs[i] = nr <= kMaxRuns ? nr : -1;
    }
  }
}

// Write pass from the run table, one WARP per ray, every sample computed independently (see linear_chain): 32 consecutive
// samples per store instruction, no shared-memory staging, no serial replay.  Rays whose chain is not linear or that have more
// than kMaxRuns runs are left to march_write_kernel (which skips the rays handled here).
__global__ void __launch_bounds__(256) march_write_runs_kernel(int64_t n_rays, float dt, const float* __restrict__ t_min,
                                                               const float* __restrict__ t_max, const int32_t* __restrict__ offsets,
                                                               const float* __restrict__ run_t0, const float* __restrict__ run_t1,
                                                               const int32_t* __restrict__ run_n, const int32_t* __restrict__ n_runs,
                                                               int64_t capacity, int32_t* __restrict__ ray_idx,
                                                               float* __restrict__ t_starts, float* __restrict__ t_ends) {
  const int lane = threadIdx.x % 32;
  const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  const int64_t n_warps = (int64_t)gridDim.x * blockDim.x / 32;
  for (int64_t i = warp_global; i < n_rays; i += n_warps) {
    const int nr = n_runs[i];
    const int base = offsets[i], total = offsets[i + 1] - base;
    if (nr <= 0 || total <= 0) continue;
    float delta;
    if (!ray_chain_is_linear(t_min[i], t_max[i], dt, &delta)) continue;
    // run r covers output samples [start_r, start_r + len_r): at most kMaxRuns runs, their starts in registers
    int start[kMaxRuns + 1];
    start[0] = 0;
#pragma unroll
    for (int r = 0; r < kMaxRuns; ++r) start[r + 1] = start[r] + (r < nr ? run_n[(int64_t)r * n_rays + i] : 0);
    for (int j = lane; j < total; j += 32) {
      int r = 0;
#pragma unroll
      for (int q = 1; q < kMaxRuns; ++q) r += (q < nr && j >= start[q]) ? 1 : 0;
      const int k = j - start[r];
      const float s0 = run_t0[(int64_t)r * n_rays + i], s1 = run_t1[(int64_t)r * n_rays + i];
      float a = s0, b = s1;
      if (k > 0) {
        a = __fadd_rn(s1, __fmul_rn((float)(k - 1), delta));
        b = __fadd_rn(s1, __fmul_rn((float)k, delta));
      }
      if ((int64_t)base + j < capacity) {                 // never write past the caller's arrays
        ray_idx[base + j] = (int32_t)i;
        t_starts[base + j] = a;
        t_ends[base + j] = b;
      }
    }
  }
}

constexpr int kStage = 32;            // samples staged per ray per round
constexpr int kStagePad = kStage + 1; // +1 float: lanes writing the same slot hit different banks
constexpr int kWarpsPerBlock = 4;

__global__ void __launch_bounds__(32 * kWarpsPerBlock) march_write_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, int64_t n_rays, MarchParams p,
    const uint8_t* __restrict__ binary, const float* __restrict__ t_min, const float* __restrict__ t_max,
    const int32_t* __restrict__ offsets, const float* __restrict__ run_t0, const float* __restrict__ run_t1,
    const int32_t* __restrict__ run_n, const int32_t* __restrict__ n_runs, int64_t capacity, int32_t* __restrict__ ray_idx,
    float* __restrict__ t_starts, float* __restrict__ t_ends, int skip_linear) {
  __shared__ float s_t0[kWarpsPerBlock][32][kStagePad];
  __shared__ float s_t1[kWarpsPerBlock][32][kStagePad];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t ray0 = (blockIdx.x * (int64_t)kWarpsPerBlock + warp) * 32;
  if (ray0 >= n_rays) return;
  const int64_t i = ray0 + lane;
  bool valid = i < n_rays;
  Marcher m;
  int written = 0, base = 0;
  // replay state: run index, samples left in the current run, running t0
  int nr = 0, run = 0, left = 0;
  float rt0 = 0.0f, rt1 = 0.0f;
  bool replay = false;
  m.tm = 1.0f; m.tmax = 0.0f;  // done
  if (valid && skip_linear && n_runs && n_runs[i] >= 0) {
    float delta;
    if (ray_chain_is_linear(t_min[i], t_max[i], p.dt, &delta)) valid = false;      // written by march_write_runs_kernel
  }
  if (valid) {
    base = offsets[i];
    nr = n_runs ? n_runs[i] : -1;
    replay = nr >= 0;
    if (!replay) {             // no run table (or more than kMaxRuns runs): march this ray again
      float o[3] = {rays_o[i * 3], rays_o[i * 3 + 1], rays_o[i * 3 + 2]};
      float d[3] = {rays_d[i * 3], rays_d[i * 3 + 1], rays_d[i * 3 + 2]};
      m.init(o, d, t_min[i], t_max[i], p.dt);
    }
  }
  float(*st0)[kStagePad] = s_t0[warp];
  float(*st1)[kStagePad] = s_t1[warp];
  while (__any_sync(0xffffffffu, !m.done() || (replay && (left > 0 || run < nr)))) {
    int cnt = 0;
    if (replay) {
      // the same fp32 chain the marcher walks inside a run: t1 = t0 + dt, next t0 = t1
      while (cnt < kStage && (left > 0 || run < nr)) {
        if (left == 0) {
          rt0 = run_t0[(int64_t)run * n_rays + i]; rt1 = run_t1[(int64_t)run * n_rays + i];
          left = run_n[(int64_t)run * n_rays + i];
          ++run;
        }
        st0[lane][cnt] = rt0; st1[lane][cnt] = rt1;
        rt0 = rt1;
        rt1 = __fadd_rn(rt0, p.dt);
        --left;
        ++cnt;
      }
    } else if (!m.done()) {
      cnt = m.advance(p, binary, kStage, [&](int j, float a, float b) { st0[lane][j] = a; st1[lane][j] = b; }, []() {});
    }
    __syncwarp();
    // cooperative flush: one ray segment at a time, 32 consecutive samples per store instruction
#pragma unroll 1
    for (int r = 0; r < 32; ++r) {
      const int c = __shfl_sync(0xffffffffu, cnt, r);
      if (c == 0) continue;
      const int dst = __shfl_sync(0xffffffffu, base + written, r);
      if (lane < c && (int64_t)dst + lane < capacity) {   // never write past the caller's arrays
        ray_idx[dst + lane] = (int32_t)(ray0 + r);
        t_starts[dst + lane] = st0[r][lane];
        t_ends[dst + lane] = st1[r][lane];
      }
    }
    written += cnt;
    __syncwarp();
  }
}

// Lazy marching, head pass: the first k0 (<= 32) samples of every ray, one pass, no count / scan: with early ray termination
// most rays never need more (a dense field is opaque after a few samples, a pruned grid leaves few samples per ray), and the
// rays that do continue from t_resume with the ordinary count -> scan -> write passes (angio_march_count with resume_alive).
// Same Marcher, so head + tail are the samples of the full march, bit for bit.  The samples are PACKED without a scan: a warp
// sums its 32 rays' counts, reserves that many slots with one atomicAdd on the running total and records each ray's first
// slot in head_base -- the order of the rays in memory depends on the order the warps arrive, every consumer goes through
// head_base, so results do not.  The warp stages its rays' samples in shared memory and flushes one ray segment per round.
__global__ void __launch_bounds__(32 * kWarpsPerBlock) march_head_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, int64_t n_rays, Aabb6 aabb, MarchParams p,
    const uint8_t* __restrict__ binary, float near_plane, float far_plane, int k0, int32_t* __restrict__ head_idx,
    float* __restrict__ head_t0, float* __restrict__ head_t1, int32_t* __restrict__ head_cnt, int32_t* __restrict__ head_base,
    int32_t* __restrict__ head_total, float* __restrict__ t_resume, float* __restrict__ t_max) {
  __shared__ float s_t0[kWarpsPerBlock][32][kStagePad];
  __shared__ float s_t1[kWarpsPerBlock][32][kStagePad];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t ray0 = (blockIdx.x * (int64_t)kWarpsPerBlock + warp) * 32;
  if (ray0 >= n_rays) return;
  const int64_t i = ray0 + lane;
  float(*st0)[kStagePad] = s_t0[warp];
  float(*st1)[kStagePad] = s_t1[warp];
  int cnt = 0;
  if (i < n_rays) {
    float o[3] = {rays_o[i * 3], rays_o[i * 3 + 1], rays_o[i * 3 + 2]};
    float d[3] = {rays_d[i * 3], rays_d[i * 3 + 1], rays_d[i * 3 + 2]};
    float a, b;
    ray_aabb(o, d, aabb.v, a, b);
    a = a < near_plane ? near_plane : a;
    b = b > far_plane ? far_plane : b;
    Marcher m;
    m.init(o, d, a, b, p.dt);
    cnt = m.advance(p, binary, k0, [&](int j, float t0, float t1) { st0[lane][j] = t0; st1[lane][j] = t1; }, []() {});
    head_cnt[i] = cnt;
    // budget reached: the last action was an emit, the state is init(t0); otherwise the ray is finished
    t_resume[i] = (cnt == k0) ? m.t0 : b;
    t_max[i] = b;
  }
  // slots of this warp's rays: inclusive warp scan of the counts, one atomic reservation per warp
  int incl = cnt;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, s);
    if (lane >= s) incl += v;
  }
  const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
  int warp_base = 0;
  if (lane == 0 && warp_total > 0) warp_base = atomicAdd(head_total, warp_total);
  warp_base = __shfl_sync(0xffffffffu, warp_base, 0);
  const int my_base = warp_base + incl - cnt;
  if (i < n_rays) head_base[i] = my_base;
  __syncwarp();
#pragma unroll 1
  for (int r = 0; r < 32; ++r) {
    const int c = __shfl_sync(0xffffffffu, cnt, r);
    if (c == 0) continue;
    const int dst = __shfl_sync(0xffffffffu, my_base, r) + lane;
    if (lane < c) {
      head_idx[dst] = (int32_t)(ray0 + r);
      head_t0[dst] = st0[r][lane];
      head_t1[dst] = st1[r][lane];
    }
  }
}

// Exclusive scan of per-ray counts (n = rays per batch: 64 K .. 1 M) by ONE thread-block cluster of 8 CTAs (portable cluster
// size): a round covers 8 x 1024 x 8 = 64 K elements -- a whole config-3 batch -- with lane-contiguous 32-byte loads / stores,
// the CTA totals are exchanged through distributed shared memory (every CTA writes its total into all eight CTAs' shared
// memory), one cluster barrier, no global scratch and no second launch.  (A single CTA needed two 9 us rounds for 64 K counts.)
constexpr int kScanCtas = 8;
constexpr int kScanPerThread = 8;
constexpr int kScanRound = kScanCtas * 1024 * kScanPerThread;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_shared_cluster(uint32_t local_smem_addr, uint32_t cta, int32_t v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(cta));
  asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(remote), "r"(v) : "memory");
}

__global__ void __cluster_dims__(kScanCtas, 1, 1) __launch_bounds__(1024) exclusive_scan_kernel(const int32_t* __restrict__ counts, int64_t n,
                                                                                                 int32_t* __restrict__ offsets,
                                                                                                 int32_t* __restrict__ total_out) {
  __shared__ int32_t s_warp[32];
  __shared__ int32_t s_cta_tot[2][kScanCtas];     // [round parity][cta]: written by every CTA of the cluster (DSMEM)
  const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
  const uint32_t cta = cluster_ctarank();
  const bool vec_ok = (reinterpret_cast<uintptr_t>(counts) % 16 == 0) && (reinterpret_cast<uintptr_t>(offsets) % 16 == 0);
  int32_t carry = 0;                               // sum of all earlier rounds (every thread of every CTA tracks it)
  int round = 0;
  for (int64_t base = 0; base < n; base += kScanRound, ++round) {
    const int64_t i0 = base + ((int64_t)cta * 1024 + tid) * kScanPerThread;
    int32_t v[kScanPerThread];
    if (vec_ok && i0 + kScanPerThread <= n) {
      const int4 q0 = *reinterpret_cast<const int4*>(counts + i0), q1 = *reinterpret_cast<const int4*>(counts + i0 + 4);
      v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
    } else {
#pragma unroll
      for (int k = 0; k < kScanPerThread; ++k) v[k] = (i0 + k < n) ? counts[i0 + k] : 0;
    }
    int32_t local = 0;
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) { const int32_t t = v[k]; v[k] = local; local += t; }   // exclusive within the thread
    int32_t incl = local;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, incl, s);
      if (lane >= s) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int32_t w = s_warp[lane];
      int32_t wi = w;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const int32_t t = __shfl_up_sync(0xffffffffu, wi, s);
        if (lane >= s) wi += t;
      }
      s_warp[lane] = wi - w;                       // exclusive prefix of warp sums
      if (lane == 31) {                            // wi = this CTA's total: publish it to every CTA of the cluster
        const uint32_t slot = (uint32_t)__cvta_generic_to_shared(&s_cta_tot[round & 1][cta]);
        for (uint32_t c = 0; c < kScanCtas; ++c) st_shared_cluster(slot, c, wi);
      }
    }
    cluster_sync_all();                            // all eight totals are in everybody's shared memory (also a CTA barrier)
    int32_t before = 0, all = 0;
#pragma unroll
    for (int c = 0; c < kScanCtas; ++c) {
      const int32_t t = s_cta_tot[round & 1][c];
      if (c < (int)cta) before += t;
      all += t;
    }
    const int32_t pre = carry + before + s_warp[warp] + incl - local;
    if (vec_ok && i0 + kScanPerThread <= n) {
      *reinterpret_cast<int4*>(offsets + i0) = make_int4(pre + v[0], pre + v[1], pre + v[2], pre + v[3]);
      *reinterpret_cast<int4*>(offsets + i0 + 4) = make_int4(pre + v[4], pre + v[5], pre + v[6], pre + v[7]);
    } else {
#pragma unroll
      for (int k = 0; k < kScanPerThread; ++k)
        if (i0 + k < n) offsets[i0 + k] = pre + v[k];
    }
    carry += all;
    __syncthreads();                               // s_warp is rewritten next round (s_cta_tot alternates by parity)
  }
  cluster_sync_all();                              // no CTA exits while a peer may still write into its shared memory
  if (cta == 0 && tid == 0) {
    offsets[n] = carry;
    if (total_out) *total_out = carry;
  }
}

__global__ void __launch_bounds__(256) grid_query_kernel(const float* __restrict__ pts, int64_t n, MarchParams p,
                                                         const uint8_t* __restrict__ binary, float* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = occupied_at(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2], p, binary) ? 1.0f : 0.0f;
}

// run table layout (angio_march_runs_bytes): [kMaxRuns][n_rays] float t0 | same for t1 | [kMaxRuns][n_rays] int32 length | [n_rays] int32 count
float* run_table_t0(void* runs, int64_t n_rays) { (void)n_rays; return reinterpret_cast<float*>(runs); }
float* run_table_t1(void* runs, int64_t n_rays) { return runs ? reinterpret_cast<float*>(runs) + (int64_t)kMaxRuns * n_rays : nullptr; }
int32_t* run_table_n(void* runs, int64_t n_rays) { return runs ? reinterpret_cast<int32_t*>(runs) + 2 * (int64_t)kMaxRuns * n_rays : nullptr; }
int32_t* run_table_count(void* runs, int64_t n_rays) { return runs ? reinterpret_cast<int32_t*>(runs) + 3 * (int64_t)kMaxRuns * n_rays : nullptr; }

MarchParams make_params(const float* roi_host, int res, float dt) {
  MarchParams p;
  p.roi = angio::make_roi(roi_host);
  for (int k = 0; k < 3; ++k) p.ext[k] = p.roi.hi[k] - p.roi.lo[k];
  p.res = res;
  p.resf = (float)res;
  p.dt = dt;
  return p;
}

}  // namespace

extern "C" int angio_march_count(const float* rays_o, const float* rays_d, int64_t n_rays, const float* aabb_host,
                                 const float* roi_host, int32_t res, const uint8_t* binary, float near_plane,
                                 float far_plane, float step_size, float* t_min, float* t_max, int32_t* counts,
                                 void* runs, const uint8_t* resume_alive, void* stream) {
  ANGIO_REQUIRE(rays_o && rays_d && aabb_host && roi_host && binary && t_min && t_max && counts, "angio_march_count: null pointer");
  ANGIO_REQUIRE(n_rays >= 0 && res > 0 && step_size > 0.0f, "angio_march_count: bad sizes (step_size must be > 0)");
  if (n_rays == 0) return 0;
  Aabb6 aabb;
  for (int k = 0; k < 6; ++k) aabb.v[k] = aabb_host[k];
  // ANGIO_MARCH_SERIAL_COUNT=1 keeps the thread-per-ray walk for every ray (A/B measurements; bit-identical output)
  const char* ev = getenv("ANGIO_MARCH_SERIAL_COUNT");
  const int fast = !(ev && ev[0] == '1') ? 1 : 0;
  const MarchParams mp = make_params(roi_host, res, step_size);
  // the serial kernel first: it stores t_min / t_max and the counts of the rays it keeps (and of rays that are not alive)
  angio::note_launch("march_count_kernel"); march_count_kernel<<<angio::blocks_for(n_rays, 128), 128, 0, angio::as_stream(stream)>>>(
      rays_o, rays_d, n_rays, aabb, mp, binary, near_plane, far_plane, t_min, t_max, counts,
      run_table_t0(runs, n_rays), run_table_t1(runs, n_rays), run_table_n(runs, n_rays), run_table_count(runs, n_rays), resume_alive, fast);
  if (int rc = angio::finish_launch("angio_march_count")) return rc;
  if (fast) {
    int64_t blocks = (n_rays + 7) / 8;                    // 8 warps per block, one ray per warp per pass
    const int64_t cap_blocks = (int64_t)angio::sm_count() * 32;
    if (blocks > cap_blocks) blocks = cap_blocks;
    angio::note_launch("march_count_warp_kernel");
    march_count_warp_kernel<<<(int)blocks, 256, 0, angio::as_stream(stream)>>>(
        rays_o, rays_d, n_rays, mp, binary, t_min, t_max, counts, run_table_t0(runs, n_rays),
        run_table_t1(runs, n_rays), run_table_n(runs, n_rays), run_table_count(runs, n_rays), resume_alive);
  }
  return angio::finish_launch("angio_march_count");
}

extern "C" int angio_march_head(const float* rays_o, const float* rays_d, int64_t n_rays, const float* aabb_host, const float* roi_host,
                                int32_t res, const uint8_t* binary, float near_plane, float far_plane, float step_size, int32_t k0,
                                int32_t* head_idx, float* head_t0, float* head_t1, int32_t* head_cnt, int32_t* head_base, int32_t* head_total,
                                float* t_resume, float* t_max, void* stream) {
  ANGIO_REQUIRE(rays_o && rays_d && aabb_host && roi_host && binary && head_idx && head_t0 && head_t1 && head_cnt && head_base &&
                    head_total && t_resume && t_max,
                "angio_march_head: null pointer");
  ANGIO_REQUIRE(n_rays >= 0 && res > 0 && step_size > 0.0f && k0 >= 1 && k0 <= kStage, "angio_march_head: bad sizes (1 <= k0 <= 32)");
  if (n_rays == 0) return 0;
  Aabb6 aabb;
  for (int k = 0; k < 6; ++k) aabb.v[k] = aabb_host[k];
  const int rays_per_block = 32 * kWarpsPerBlock;
  angio::note_launch("march_head_kernel"); march_head_kernel<<<angio::blocks_for(n_rays, rays_per_block), rays_per_block, 0, angio::as_stream(stream)>>>(
      rays_o, rays_d, n_rays, aabb, make_params(roi_host, res, step_size), binary, near_plane, far_plane, k0, head_idx, head_t0, head_t1,
      head_cnt, head_base, head_total, t_resume, t_max);
  return angio::finish_launch("angio_march_head");
}

extern "C" int64_t angio_march_runs_bytes(int64_t n_rays) { return n_rays < 0 ? ANGIO_ERR_INVALID_ARG : (3 * (int64_t)kMaxRuns + 1) * n_rays * 4; }

extern "C" int angio_exclusive_scan_i32(const int32_t* counts, int64_t n, int32_t* offsets, int32_t* total_out, void* stream) {
  ANGIO_REQUIRE(offsets && (counts || n == 0) && n >= 0, "angio_exclusive_scan_i32: bad arguments");
  angio::note_launch("exclusive_scan_kernel"); exclusive_scan_kernel<<<kScanCtas, 1024, 0, angio::as_stream(stream)>>>(counts, n, offsets, total_out);
  return angio::finish_launch("angio_exclusive_scan_i32");
}

extern "C" int angio_march_write(const float* rays_o, const float* rays_d, int64_t n_rays, const float* roi_host, int32_t res,
                                 const uint8_t* binary, float step_size, const float* t_min, const float* t_max,
                                 const int32_t* offsets, const void* runs, int64_t capacity, int32_t* ray_idx, float* t_starts,
                                 float* t_ends, void* stream) {
  ANGIO_REQUIRE(rays_o && rays_d && roi_host && binary && t_min && t_max && offsets && ray_idx && t_starts && t_ends,
                "angio_march_write: null pointer");
  ANGIO_REQUIRE(n_rays >= 0 && res > 0 && step_size > 0.0f, "angio_march_write: bad sizes");
  if (n_rays == 0) return 0;
  const int rays_per_block = 32 * kWarpsPerBlock;
  const int64_t cap = capacity > 0 ? capacity : INT64_MAX;
  void* rt = const_cast<void*>(runs);
  // ANGIO_MARCH_SERIAL_WRITE=1 keeps the serial replay for every ray (A/B measurements; bit-identical output)
  const char* ev = getenv("ANGIO_MARCH_SERIAL_WRITE");
  const int fast = (runs != nullptr && !(ev && ev[0] == '1')) ? 1 : 0;
  if (fast) {
    int64_t blocks = (n_rays + 7) / 8;                    // 8 warps per block, one ray per warp per pass
    const int64_t cap_blocks = (int64_t)angio::sm_count() * 32;
    if (blocks > cap_blocks) blocks = cap_blocks;
    angio::note_launch("march_write_runs_kernel");
    march_write_runs_kernel<<<(int)blocks, 256, 0, angio::as_stream(stream)>>>(n_rays, step_size, t_min, t_max, offsets, run_table_t0(rt, n_rays),
                                                                               run_table_t1(rt, n_rays), run_table_n(rt, n_rays),
                                                                               run_table_count(rt, n_rays), cap, ray_idx, t_starts, t_ends);
    if (int rc = angio::finish_launch("march_write_runs_kernel")) return rc;
  }
  angio::note_launch("march_write_kernel"); march_write_kernel<<<angio::blocks_for(n_rays, rays_per_block), rays_per_block, 0, angio::as_stream(stream)>>>(
      rays_o, rays_d, n_rays, make_params(roi_host, res, step_size), binary, t_min, t_max, offsets,
      run_table_t0(rt, n_rays), run_table_t1(rt, n_rays), run_table_n(rt, n_rays), run_table_count(rt, n_rays),
      cap, ray_idx, t_starts, t_ends, fast);
  return angio::finish_launch("angio_march_write");
}

extern "C" int angio_grid_query(const float* points, int64_t n, const float* roi_host, int32_t res, const uint8_t* binary,
                                float* out, void* stream) {
  ANGIO_REQUIRE(points && roi_host && binary && out && n >= 0 && res > 0, "angio_grid_query: bad arguments");
  if (n == 0) return 0;
  angio::note_launch("grid_query_kernel"); grid_query_kernel<<<angio::blocks_for(n, 256), 256, 0, angio::as_stream(stream)>>>(points, n, make_params(roi_host, res, 1.0f), binary, out);
  return angio::finish_launch("angio_grid_query");
}
