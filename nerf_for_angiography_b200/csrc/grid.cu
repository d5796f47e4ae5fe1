// grid.cu -- occupancy-grid refresh.  Replaces nerfacc OccupancyGrid._update (pure torch ops in the library),
// reached from acc_update_n_step (/root/reference/nerf/nerf_helpers_acc.py:65-78):
//   x = (cell_coords + U[0,1)^3) / res mapped into the AABB      -> angio_grid_cell_points
//   occ = sigmoid(MLP(x))                                        -> angio_mlp_forward(ANGIO_OUT_SIGMA)
//   occs[cell] = max(occs[cell] * decay, occ)                    -> angio_grid_ema_update
//   binary = occs > min(mean(occs), occ_thre)                    -> angio_grid_threshold
// All HBM-bound elementwise / reduction work over 128^3 = 2 M cells (8 MB of fp32 occs).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) cell_points_kernel(const int64_t* __restrict__ cells, const float* __restrict__ jitter,
                                                          int64_t n, angio::Roi roi, int res, float* __restrict__ pts) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t c = cells ? cells[i] : i;
  const int iz = (int)(c % res), iy = (int)((c / res) % res), ix = (int)(c / ((int64_t)res * res));
  const float resf = (float)res;
  const int idx[3] = {ix, iy, iz};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // (grid_coords + rand) / resolution, then contract_inv: x * (hi - lo) + lo   (no FMA contraction)
    const float u = __fdiv_rn(__fadd_rn((float)idx[k], jitter[i * 3 + k]), resf);
    pts[i * 3 + k] = __fadd_rn(__fmul_rn(u, __fsub_rn(roi.hi[k], roi.lo[k])), roi.lo[k]);
  }
}

// First pass: scale the touched cells by `decay` exactly once (mark with the sign bit), second pass: atomic max.
// occ >= 0 and occs >= 0, so the int ordering of the float bit patterns equals the float ordering.
__global__ void __launch_bounds__(256) ema_decay_kernel(float* __restrict__ occs, const int64_t* __restrict__ cells, int64_t n,
                                                        float decay, uint32_t* __restrict__ touched_bits) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t c = cells ? cells[i] : i;
  const uint32_t bit = 1u << (c & 31);
  const uint32_t old = atomicOr(touched_bits + (c >> 5), bit);
  if (!(old & bit)) occs[c] = __fmul_rn(occs[c], decay);
}

__global__ void __launch_bounds__(256) ema_max_kernel(float* __restrict__ occs, const int64_t* __restrict__ cells,
                                                      const float* __restrict__ occ, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t c = cells ? cells[i] : i;
  atomicMax(reinterpret_cast<int*>(occs) + c, __float_as_int(fmaxf(occ[i], 0.0f)));
}

// all-cells fast path (warm-up phase: every cell exactly once, no duplicates): one fused elementwise pass
__global__ void __launch_bounds__(256) ema_all_kernel(float* __restrict__ occs, const float* __restrict__ occ, int64_t n, float decay) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  occs[i] = fmaxf(__fmul_rn(occs[i], decay), occ[i]);
}

// deterministic two-level mean: per-block partial sums (fixed order), then one block folds the partials
__global__ void __launch_bounds__(256) partial_sum_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ partials) {
  __shared__ float s[256];
  float acc = 0.0f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += x[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) s[threadIdx.x] += s[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = s[0];
}

__global__ void __launch_bounds__(256) final_mean_kernel(const float* __restrict__ partials, int n_partials, int64_t n,
                                                         float* __restrict__ mean_out) {
  __shared__ float s[256];
  float acc = 0.0f;
  for (int i = threadIdx.x; i < n_partials; i += 256) acc += partials[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) s[threadIdx.x] += s[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) *mean_out = s[0] / (float)n;
}

__global__ void __launch_bounds__(256) threshold_kernel(const float* __restrict__ occs, int64_t n, const float* __restrict__ mean,
                                                        float occ_thre, uint8_t* __restrict__ binary) {
  const float thre = fminf(*mean, occ_thre);  // torch.clamp(occs.mean(), max=occ_thre)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    binary[i] = occs[i] > thre ? 1 : 0;
}

constexpr int kPartials = 1024;

}  // namespace

extern "C" int angio_grid_cell_points(const int64_t* cells, const float* jitter, int64_t n, const float* roi_host, int32_t res,
                                      float* points, void* stream) {
  ANGIO_REQUIRE(jitter && roi_host && points && n >= 0 && res > 0, "angio_grid_cell_points: bad arguments");
  if (n == 0) return 0;
  angio::note_launch("cell_points_kernel"); cell_points_kernel<<<angio::blocks_for(n, 256), 256, 0, angio::as_stream(stream)>>>(cells, jitter, n, angio::make_roi(roi_host), res, points);
  return angio::finish_launch("angio_grid_cell_points");
}

// workspace: a bitset of ceil(n_cells/32) words (only needed when cells != NULL)
extern "C" int angio_grid_ema_update(float* occs, int64_t n_cells, const int64_t* cells, const float* occ, int64_t n, float decay,
                                     void* workspace, int64_t workspace_bytes, void* stream) {
  ANGIO_REQUIRE(occs && occ && n >= 0 && n_cells > 0, "angio_grid_ema_update: bad arguments");
  if (n == 0) return 0;
  cudaStream_t st = angio::as_stream(stream);
  if (!cells) {
    ANGIO_REQUIRE(n == n_cells, "angio_grid_ema_update: cells == NULL requires n == n_cells");
    angio::note_launch("ema_all_kernel"); ema_all_kernel<<<angio::blocks_for(n, 256), 256, 0, st>>>(occs, occ, n, decay);
    return angio::finish_launch("angio_grid_ema_update(all)");
  }
  const int64_t need = ((n_cells + 31) / 32) * 4;
  if (workspace_bytes < need || !workspace) {
    angio::set_error("angio_grid_ema_update: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need);
    return ANGIO_ERR_WORKSPACE;
  }
  ANGIO_CUDA(cudaMemsetAsync(workspace, 0, need, st));
  angio::note_launch("ema_decay_kernel"); ema_decay_kernel<<<angio::blocks_for(n, 256), 256, 0, st>>>(occs, cells, n, decay, reinterpret_cast<uint32_t*>(workspace));
  angio::note_launch("ema_max_kernel"); ema_max_kernel<<<angio::blocks_for(n, 256), 256, 0, st>>>(occs, cells, occ, n);
  return angio::finish_launch("angio_grid_ema_update");
}

extern "C" int angio_grid_threshold(const float* occs, int64_t n_cells, float occ_thre, uint8_t* binary, float* mean_out,
                                    void* workspace, int64_t workspace_bytes, void* stream) {
  ANGIO_REQUIRE(occs && binary && mean_out && n_cells > 0, "angio_grid_threshold: bad arguments");
  if (!workspace || workspace_bytes < (int64_t)kPartials * 4) {
    angio::set_error("angio_grid_threshold: workspace needs %d bytes", kPartials * 4);
    return ANGIO_ERR_WORKSPACE;
  }
  cudaStream_t st = angio::as_stream(stream);
  float* partials = reinterpret_cast<float*>(workspace);
  angio::note_launch("partial_sum_kernel"); partial_sum_kernel<<<kPartials, 256, 0, st>>>(occs, n_cells, partials);
  angio::note_launch("final_mean_kernel"); final_mean_kernel<<<1, 256, 0, st>>>(partials, kPartials, n_cells, mean_out);
  int blocks = angio::blocks_for(n_cells, 256);
  int cap = angio::sm_count() * 8;
  angio::note_launch("threshold_kernel"); threshold_kernel<<<blocks > cap ? cap : blocks, 256, 0, st>>>(occs, n_cells, mean_out, occ_thre, binary);
  return angio::finish_launch("angio_grid_threshold");
}
