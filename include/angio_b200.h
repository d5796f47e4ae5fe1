/*
 * angio_b200.h -- C ABI of libangio_b200.so, the B200 (sm_100a) implementation of the
 * kirstenmaas/nerf-for-angiography training / rendering hot path.
 *
 * The reference has no FFI of its own: its hot path is a sequence of Python calls into torch,
 * nerfacc and torch_scatter (nerf/run_nerf_acc.py:284-307).  Each entry point below names the
 * reference call it replaces.  The Python package nerf_for_angiography_b200 binds these with
 * ctypes; nothing here mentions torch.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - the caller allocates and owns every buffer (including workspaces); nothing is retained;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no implicit sync;
 *   - return value: 0 = OK, >0 = cudaError_t, <0 = ANGIO_ERR_*; angio_last_error_string() gives
 *     a thread-local description of the last failure;
 *   - sample arrays are "packed by ray": samples of ray r occupy [offsets[r], offsets[r+1]) in
 *     ascending t, rays in ascending order (nerfacc's packed layout).
 */
#ifndef ANGIO_B200_H
#define ANGIO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANGIO_B200_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define ANGIO_API __attribute__((visibility("default")))
#else
#define ANGIO_API
#endif

#define ANGIO_ERR_INVALID_ARG (-1)
#define ANGIO_ERR_UNSUPPORTED (-2)
#define ANGIO_ERR_WORKSPACE (-3)

ANGIO_API int angio_version(void);
ANGIO_API const char* angio_last_error_string(void);
/* number of SMs of the current device (grid sizing for persistent kernels) */
ANGIO_API int angio_sm_count(void);
/* total number of kernels this library has launched in this process (bench.py's gpu_launches) */
ANGIO_API int64_t angio_launch_count(void);

/* Per-launch timeline of this library's kernels (measurement only; bench.py's per-kernel rooflines).  Between
 * angio_profile_start(stream) and angio_profile_stop() every kernel launch of the library first records a CUDA event on
 * `stream` (which must be the stream the calls are issued on); the time between two consecutive events is the device time
 * of the earlier launch plus anything the caller enqueued in between.  angio_profile_stop waits for the last event and
 * returns the number of launches recorded (< 0: error); angio_profile_entry(i, ...) returns launch i's kernel name and
 * its milliseconds.  One profile at a time, single-threaded use. */
ANGIO_API int angio_profile_start(void* stream);
ANGIO_API int64_t angio_profile_stop(void);
ANGIO_API int angio_profile_entry(int64_t i, char* name_out, int32_t name_cap, float* ms_out);

/* ------------------------------------------------------------------------------------------------
 * Cone-beam ray generation.   Replaces phantomdata/helpers.py:156-175 (get_ray_values) + the
 * .float() cast at nerf/run_nerf_acc.py:88-89 / visualization/visualization.py:330-331.
 * cam2world: [n_views,4,4] float64 row-major (source_matrix output, proj_helpers.py:68-77).
 * Arithmetic is float64 in the reference's operation order, rounded once to float32.
 *   image mode : view_ids == NULL -> rays of view `view0`, pixel order row-major [H, W] (n = W*H)
 *   gather mode: view_ids/px/py [n] int32 (px = column = x_position, py = row = y_position)
 * pixels (optional): [n_views, H, W] float32 -> pix_out[n] (gather of the target value).
 */
ANGIO_API int angio_raygen(const double* cam2world, int32_t view0, const int32_t* view_ids, const int32_t* px,
                 const int32_t* py, int64_t n, int32_t img_w, int32_t img_h, double focal,
                 const float* pixels, float* rays_o, float* rays_d, float* pix_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Weighted ray sampling without replacement.   Replaces DataFrame.sample(n, weights) in sample_pixel_rays
 * (nerf/nerf_helpers.py:137-150).  Exponential race: key_i = -log(u_i)/w_i, u from a counter-based hash of
 * (seed, i); rays with key < tau are appended to (cand_keys, cand_ids) (capacity entries; *counter must be 0 on
 * entry and receives the number of candidates).  The caller takes the n smallest keys.  weights may be NULL
 * (uniform).  cand_keys should be pre-filled with +inf so unfilled slots never win.
 */
ANGIO_API int angio_sample_candidates(const float* weights, int64_t n_pool, uint64_t seed, float tau, int32_t capacity,
                                      float* cand_keys, int64_t* cand_ids, int32_t* counter, void* stream);

/* One-call sampler: candidates (above) -> exact selection of the n smallest keys (radix select) -> uniform shuffle
 * (the `.sample(frac=1)` of nerf/nerf_helpers.py:139), all stream-ordered, no host sync.  ids_out: [n] int64 flat ray ids
 * (view * H * W + y * W + x).  status: 2 x int32 on the device: [0] = number of candidates, [1] = 1 if the candidate
 * buffer overflowed or held fewer than n rays (ids_out is then undefined; re-draw with a larger tau / capacity).
 * The permutation depends only on (seed, selected set), and candidates whose key equals the n-th smallest are taken in ray-id
 * order: the draw is a function of the seed although candidates are appended in arbitrary order.
 */
ANGIO_API int64_t angio_sample_rays_workspace_bytes(int32_t capacity, int64_t n);
ANGIO_API int angio_sample_rays(const float* weights, int64_t n_pool, int64_t n, uint64_t seed, float tau, int32_t capacity,
                                int64_t* ids_out, int32_t* status, void* workspace, int64_t workspace_bytes, void* stream);
/* Check mode of the sampler (the second half of angio_sample_rays on caller-supplied candidates): the n candidates with the
 * smallest keys -- positive fp32 race keys, e.g. pre-drawn Exp(1)/weight variates a numpy reference of
 * DataFrame.sample(n, weights) (nerf/nerf_helpers.py:137-150) draws from the same uniforms -- shuffled as above.  Equal keys at
 * the selection threshold are resolved by ray id (smallest first), so the drawn set depends on (keys, ids) alone.
 * cand_keys / cand_ids: [m] on the device; workspace: angio_sample_rays_workspace_bytes(m, n).  status as above.
 */
ANGIO_API int angio_sample_select(const float* cand_keys, const int64_t* cand_ids, int32_t m, int64_t n, uint64_t seed,
                                  int64_t* ids_out, int32_t* status, void* workspace, int64_t workspace_bytes, void* stream);
/* angio_raygen in gather mode driven by flat ray ids (the sampler's output); pixels: [n_views, H, W] -> pix_out[n] */
ANGIO_API int angio_raygen_flat(const double* cam2world, const int64_t* ids, int64_t n, int32_t img_w, int32_t img_h,
                                double focal, const float* pixels, float* rays_o, float* rays_d, float* pix_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Occupancy-grid ray marching.   Replaces nerfacc.ray_marching's slab test + two-pass
 * _C.ray_marching (called at nerf/nerf_helpers_acc.py:29).  binary: [res,res,res] uint8 (torch.bool),
 * x-major.  roi / aabb: 6 floats (min xyz, max xyz) on the HOST.
 */
/* pass 0+1: t_min/t_max per ray (slab test clamped to [near,far]) and per-ray sample counts.
 * runs (optional, angio_march_runs_bytes(n_rays) bytes): the count pass records each ray's runs of consecutive samples
 * (start t0 + length, up to 8 per ray); given the same table, the write pass replays the fp32 t-chain of the runs instead of
 * marching the grid a second time (rays with more runs are re-marched).  The samples are bit-identical either way. */
ANGIO_API int64_t angio_march_runs_bytes(int64_t n_rays);
ANGIO_API int angio_march_count(const float* rays_o, const float* rays_d, int64_t n_rays, const float* aabb_host,
                      const float* roi_host, int32_t res, const uint8_t* binary, float near_plane,
                      float far_plane, float step_size, float* t_min, float* t_max, int32_t* counts,
                      void* runs, const uint8_t* resume_alive, void* stream);
/* Lazy marching.  angio_march_head marches only the first k0 (<= 32) samples of every ray, in ONE pass: ray r gets head_cnt[r]
 * samples in slots head_base[r] ... of head_idx / head_t0 / head_t1 (capacity n_rays * k0).  The samples are packed without a
 * scan -- every warp reserves the slots of its 32 rays with one atomic add on *head_total, which must be 0 on entry and holds
 * the number of head samples afterwards -- so the ORDER of the rays in memory is unspecified; everything downstream goes
 * through head_base.  t_resume / t_max per ray: with early ray termination most rays need nothing more; the others are
 * continued with angio_march_count(resume_alive = their flags, t_min = t_resume and t_max as INPUTS) -> scan ->
 * angio_march_write: head + tail are exactly the samples of the full march. */
ANGIO_API int angio_march_head(const float* rays_o, const float* rays_d, int64_t n_rays, const float* aabb_host,
                     const float* roi_host, int32_t res, const uint8_t* binary, float near_plane, float far_plane,
                     float step_size, int32_t k0, int32_t* head_idx, float* head_t0, float* head_t1,
                     int32_t* head_cnt, int32_t* head_base, int32_t* head_total, float* t_resume, float* t_max,
                     void* stream);
/* exclusive scan: offsets[n+1] int32 (offsets[n] = total); also copies the total to *total_out if not NULL */
ANGIO_API int angio_exclusive_scan_i32(const int32_t* counts, int64_t n, int32_t* offsets, int32_t* total_out,
                             void* stream);
/* pass 2: write samples.  ray_idx [n] int32, t_starts / t_ends [n] float32 */
ANGIO_API int angio_march_write(const float* rays_o, const float* rays_d, int64_t n_rays, const float* roi_host,
                      int32_t res, const uint8_t* binary, float step_size, const float* t_min,
                      const float* t_max, const int32_t* offsets, const void* runs, int64_t capacity,
                      int32_t* ray_idx, float* t_starts, float* t_ends, void* stream);
/* capacity (here and in angio_compact_samples): number of elements the output arrays hold; samples beyond it are dropped
 * instead of written (0 = unchecked).  The sync-free callers size the arrays by an upper bound and never hit it. */
/* nerfacc OccupancyGrid.query_occ (visualization/visualization.py:214): occupancy 0/1 at points [n,3] */
ANGIO_API int angio_grid_query(const float* points, int64_t n, const float* roi_host, int32_t res,
                     const uint8_t* binary, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Visibility filter + compaction.   Replaces nerfacc render_visibility + the three boolean-mask
 * gathers inside nerfacc.ray_marching (after alpha_fn, nerf/nerf_helpers_acc.py:11-25).
 * T_0 = 1, T_{i+1} = T_i * (1 - alpha_i) sequentially in fp32; keep iff T_i >= early_stop_eps and
 * (alpha_thre <= 0 or alpha_i >= alpha_thre).
 */
/* t_init (optional, [n_rays]): transmittance each ray starts with (tail of a lazily marched ray; NULL = 1);
 * base_counts (optional): added to kept_counts (the kept head samples);
 * alpha_thre_cap (optional, DEVICE float): the threshold used is min(alpha_thre, *alpha_thre_cap) -- nerfacc.ray_marching
 * clamps alpha_thre to mean(grid.occs); passing the device-resident mean keeps the training loop free of host syncs. */
ANGIO_API int angio_visibility_mask(const float* alphas, const int32_t* offsets, int64_t n_rays,
                          float early_stop_eps, float alpha_thre, uint8_t* keep, int32_t* kept_counts,
                          const float* t_init, const int32_t* base_counts, const float* alpha_thre_cap, void* stream);
/* Visibility of the head samples of angio_march_head: keep flags (indexed like the head arrays), kept count, transmittance
 * behind the head (t_end) and alive[r] = the head used its whole budget and t_end >= early_stop_eps (the ray must be
 * continued). */
ANGIO_API int angio_visibility_head_mask(const float* alphas, const int32_t* head_cnt, const int32_t* head_base,
                               int64_t n_rays, int32_t k0, float early_stop_eps, float alpha_thre, uint8_t* keep,
                               int32_t* kept_counts, float* t_end, uint8_t* alive, const float* alpha_thre_cap,
                               void* stream);
/* Compaction of a lazily marched batch into the packed layout: per ray the kept head samples, then the kept tail samples. */
ANGIO_API int angio_compact_head_tail(const uint8_t* keep_head, const int32_t* head_cnt, const int32_t* head_base,
                            const float* head_t0, const float* head_t1, const uint8_t* keep_tail,
                            const int32_t* tail_offsets, const float* tail_t0, const float* tail_t1,
                            const int32_t* new_offsets, int64_t n_rays, int64_t capacity, int32_t* ray_idx_out,
                            float* t_starts_out, float* t_ends_out, void* stream);
/* Two-phase visibility pass (early ray termination; the kept set is bit-identical to evaluating every sample): phase A
 * evaluates the first k0 samples of every ray, angio_visibility_head marks the rays whose transmittance after them is still
 * >= early_stop_eps, phase B evaluates the remaining samples of those rays only.  The sample subsets are handed to
 * angio_mlp_forward as index lists (angio_samples.sample_idx):
 *   counts[r]      = number of samples of ray r with local index in [skip, skip + limit)   (limit < 0: unbounded;
 *                    0 where alive[r] == 0, alive may be NULL)
 *   sample_ids[..] = their global indices, rays in order, after an exclusive scan of counts (seg_offsets)
 * angio_visibility_mask stops reading alphas behind the termination point (alphas must lie in [0, 1]).
 */
ANGIO_API int angio_ray_segment_counts(const int32_t* offsets, int64_t n_rays, int32_t skip, int32_t limit,
                             const uint8_t* alive, int32_t* counts, void* stream);
ANGIO_API int angio_ray_segment_ids(const int32_t* offsets, const int32_t* seg_offsets, int64_t n_rays, int32_t skip,
                          int32_t* sample_ids, void* stream);
ANGIO_API int angio_visibility_head(const float* alphas, const int32_t* offsets, int64_t n_rays, int32_t k0,
                          float early_stop_eps, uint8_t* alive, void* stream);
ANGIO_API int angio_compact_samples(const uint8_t* keep, const int32_t* offsets, const int32_t* new_offsets,
                          int64_t n_rays, const float* t_starts, const float* t_ends, int64_t capacity,
                          int32_t* ray_idx_out, float* t_starts_out, float* t_ends_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * CPPN MLP (model/CPPN.py:96-131,166-222; relu, no skip, no view directions).
 */
typedef struct angio_mlp_desc {
  int32_t enc;       /* 0 = none, 1 = fourier (model/CPPN.py:216-222) */
  int32_t enc_basis; /* pos_enc_basis (L); input width = 3 + 6*L when enc != 0 */
  int32_t width;     /* num_filters (H) */
  int32_t n_hidden;  /* num_early_layers: number of HxH layers after the input layer */
} angio_mlp_desc;

/* Flat fp32 parameter layout (angio_mlp_param_count floats):
 *   [fourier_coefficients (3*L) if enc] [W_0 (H x D_in)] [b_0 (H)] [W_1 (H x H)] [b_1] ... [W_n_hidden] [b_n_hidden]
 *   [W_out (1 x H)] [b_out (1)]        every W row-major [out][in] exactly like torch.nn.Linear.weight */
ANGIO_API int64_t angio_mlp_param_count(const angio_mlp_desc* desc);
ANGIO_API int32_t angio_mlp_input_width(const angio_mlp_desc* desc);

/* Where the MLP inputs come from.  Either explicit points, or ray samples whose midpoint
 * x = o[idx] + d[idx] * (t0 + t1) / 2 (nerf/run_nerf_acc.py:290-292) is formed in-kernel. */
typedef struct angio_samples {
  int64_t n;              /* number of samples */
  const float* points;    /* [n,3] or NULL */
  const float* rays_o;    /* [R,3] (when points == NULL) */
  const float* rays_d;    /* [R,3] */
  const int32_t* ray_idx; /* [n] */
  const float* t_starts;  /* [n] */
  const float* t_ends;    /* [n] */
  const int32_t* sample_idx; /* optional [n] index list: sample k of this call is element sample_idx[k] of ray_idx /
                                t_starts / t_ends, and its output goes to out[sample_idx[k]] (inference forward of the
                                bf16 path only; NULL = identity) */
  const int32_t* n_dev;   /* optional DEVICE-resident sample count: kernels process min(*n_dev, n) samples, so a marcher
                             that leaves its total on the device can feed the MLP without a host sync (n = capacity of
                             the arrays; tile-image layouts of the saved activations are strided by that capacity).
                             Supported by the bf16 (tcgen05) forward and backward; must be NULL for ANGIO_PREC_FP32. */
} angio_samples;

#define ANGIO_OUT_LOGIT 0 /* raw model output (CPPN.forward)                                   */
#define ANGIO_OUT_SIGMA 1 /* sigmoid(logit): occ_eval_fn, nerf/nerf_helpers_acc.py:66-70        */
#define ANGIO_OUT_ALPHA 2 /* 1 - exp(-sigmoid(logit)*(t1-t0)): alpha_fn, nerf_helpers_acc.py:11-25 */

#define ANGIO_PREC_FP32 0 /* fp32 SIMT check mode (reference arithmetic)                        */
#define ANGIO_PREC_BF16 1 /* bf16 tcgen05 tensor cores, fp32 accumulate                          */

/* bytes of caller-provided workspace for forward / backward at n samples */
ANGIO_API int64_t angio_mlp_workspace_bytes(const angio_mlp_desc* desc, int64_t n, int32_t precision, int32_t training);
/* bytes of the saved-activation buffer a training forward fills for the backward */
ANGIO_API int64_t angio_mlp_saved_bytes(const angio_mlp_desc* desc, int64_t n, int32_t precision);
/* bytes of the packed (bf16, UMMA-swizzled) weight image used by ANGIO_PREC_BF16 */
ANGIO_API int64_t angio_mlp_packed_bytes(const angio_mlp_desc* desc);
/* refresh the packed weight image from the fp32 master parameters (after every optimiser step) */
ANGIO_API int angio_mlp_pack_weights(const angio_mlp_desc* desc, const float* params, void* packed, void* stream);

/* Replaces CPPN.forward / get_predictions (nerf/nerf_helpers.py:24-45).  out: [n] float32.
 * saved != NULL makes it a training forward (activations kept for angio_mlp_backward). */
ANGIO_API int angio_mlp_forward(const angio_mlp_desc* desc, const float* params, const void* packed,
                      const angio_samples* in, int32_t out_mode, int32_t precision, float* out,
                      void* saved, void* workspace, int64_t workspace_bytes, void* stream);
/* Replaces autograd through CPPN.forward: grad_params (flat layout above) = d(sum(out*grad_out))/d(params).
 * grad_params is OVERWRITTEN.  grad_out: [n] float32 w.r.t. the raw logit. */
ANGIO_API int angio_mlp_backward(const angio_mlp_desc* desc, const float* params, const void* packed,
                       const angio_samples* in, const void* saved, const float* grad_out,
                       int32_t precision, float* grad_params, void* workspace, int64_t workspace_bytes,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Beer-Lambert attenuation line integral.   Replaces acc_render_volume_density
 * (nerf/nerf_helpers_acc.py:45-63): pix[r] = prod_i exp(-sigmoid(p_i) * (t1_i - t0_i)), rays without
 * samples render 1.  zero_mask (optional, [n] uint8): sigma forced to 0 where set (the `zero_idx` argument).
 */
ANGIO_API int angio_composite_forward(const float* logits, const float* t_starts, const float* t_ends,
                            const int32_t* offsets, int64_t n_rays, const uint8_t* zero_mask, float* pix,
                            void* stream);
/* analytic backward: grad_logits[i] = grad_pix[r] * pix[r] * (-(t1-t0)) * s * (1 - s) */
ANGIO_API int angio_composite_backward(const float* logits, const float* t_starts, const float* t_ends,
                             const int32_t* offsets, int64_t n_rays, const uint8_t* zero_mask,
                             const float* pix, const float* grad_pix, float* grad_logits, void* stream);
/* fused training tail: composite forward + mse_loss (nerf/run_nerf_acc.py:296-298) + d(loss)/d(logits).
 * loss_sum (1 float, must be zeroed by the caller) accumulates sum_r (pix_r - target_r)^2;
 * grad_logits is d(mean over n_rays_total)/d(logit) (n_rays_total = global batch for data parallel). */
ANGIO_API int angio_composite_mse_fused(const float* logits, const float* t_starts, const float* t_ends,
                              const int32_t* offsets, int64_t n_rays, const float* target,
                              int64_t n_rays_total, float* pix, float* grad_logits, float* loss_sum,
                              void* stream);

/* ------------------------------------------------------------------------------------------------
 * Occupancy-grid refresh.   Replaces nerfacc OccupancyGrid._update (reached from acc_update_n_step,
 * nerf/nerf_helpers_acc.py:65-78).
 */
/* x = (cell_coords + jitter) / res mapped to the AABB; cells == NULL means all cells 0..n-1 */
ANGIO_API int angio_grid_cell_points(const int64_t* cells, const float* jitter, int64_t n, const float* roi_host,
                           int32_t res, float* points, void* stream);
/* occs[cell] = max(occs[cell] * decay, occ)  (decay applied once per touched cell, max over duplicates).
 * cells == NULL: all cells (n == n_cells, no workspace).  Otherwise workspace = ceil(n_cells/32)*4 bytes. */
ANGIO_API int angio_grid_ema_update(float* occs, int64_t n_cells, const int64_t* cells, const float* occ, int64_t n,
                          float decay, void* workspace, int64_t workspace_bytes, void* stream);
/* binary = occs > min(mean(occs), occ_thre); mean_out (1 float) receives mean(occs) */
ANGIO_API int angio_grid_threshold(const float* occs, int64_t n_cells, float occ_thre, uint8_t* binary,
                         float* mean_out, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser.   Replaces torch.optim.Adam.step (nerf/run_nerf_acc.py:206,305-307) on the flat buffers.
 * step is 1-based.  grad_scale multiplies the gradient first (1/world_size after an all-reduce sum).
 * active (optional, DEVICE float): when non-NULL and *active == 0 the step is skipped -- the sync-free training loop keeps
 * its kept-sample count on the device, and an iteration without samples takes no optimiser step (nerf/run_nerf_acc.py:289).
 */
ANGIO_API int angio_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                    float lr, float beta1, float beta2, float eps, int32_t step, float grad_scale,
                    const float* active, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Ground-truth projector (synthetic data generation).   Replaces ray_tracing (phantomdata/helpers.py:192-224): trilinear
 * lookups (0 outside the grid, like scipy's RegularGridInterpolator(bounds_error=False, fill_value=0)) of volume[nx,ny,nz]
 * (x-major, spanning bounds_host = min xyz, max xyz) at o + d * depths[k]; ct_mode 1: exp(-sum mu_k * dist_k * |d|) with
 * dist = diff(depths) and the reference's 1e10 tail; ct_mode 0 ('sdf'): exp(-sum mu_k).  depths: [n_depths] DEVICE floats
 * shared by all rays.  out: [n_rays].
 */
ANGIO_API int angio_project_volume(const float* volume, int32_t nx, int32_t ny, int32_t nz, const float* bounds_host,
                         const float* rays_o, const float* rays_d, int64_t n_rays, const float* depths,
                         int32_t n_depths, int32_t ct_mode, float* out, void* stream);

/* Data-parallel training (one process per GPU): gradient all-reduce fused INTO the optimiser step over NVLink peer memory.
 * Each rank's gradient lives in a buffer that every peer has mapped (symmetric memory).  After its backward a rank calls
 * angio_signal_peers (system-scope fence + one tag store into every peer's flag array: peer_flags_host[r] is the DEVICE address
 * of rank r's flag array [world] as mapped in this process).  angio_adam_step_allreduce waits until my_flags[r] >= tag for all r,
 * sums the `world` gradient buffers (peer_grads_host[r]: device address of rank r's gradient as mapped here) in rank order --
 * bit-identical on every rank -- and applies angio_adam_step's update in the same pass.  active_index >= 0 names a gradient
 * slot whose all-rank sum gates the step (0 = skip), < 0 disables the gate.  Callers alternate two gradient buffers by step
 * parity.  Replaces the NCCL all_reduce + Adam pair; the pointer arrays are HOST arrays of `world` (<= 16) entries.
 * wait_stats (optional, DEVICE uint64[4]): block 0 accumulates [nanoseconds spent waiting for the peers' tags (globaltimer),
 * number of calls, longest single wait, 0] -- the measurement of how long a rank stalls on its slowest peer.  A peer that
 * does not signal within ANGIO_PEER_TIMEOUT_S seconds (environment, default 600) traps the kernel. */
ANGIO_API int angio_signal_peers(void* const* peer_flags_host, int32_t world, int32_t rank, uint32_t tag, void* stream);
ANGIO_API int angio_adam_step_allreduce(float* params, const void* const* peer_grads_host, int32_t world,
                              const uint32_t* my_flags, uint32_t tag, float* exp_avg, float* exp_avg_sq, int64_t n,
                              float lr, float beta1, float beta2, float eps, int32_t step, float grad_scale,
                              int64_t active_index, unsigned long long* wait_stats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ANGIO_B200_H */
