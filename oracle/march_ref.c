/*
 * oracle/march_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, scalar fp32, one ray at a time) of the third-party
 * `nerfacc` 0.3.x arithmetic the reference calls on its hot path:
 *
 *   nerfacc.ray_marching(...)              <- /root/reference/nerf/nerf_helpers_acc.py:29
 *   nerfacc.OccupancyGrid.query_occ(...)   <- /root/reference/visualization/visualization.py:214
 *   nerfacc render_visibility(...)         <- called inside nerfacc.ray_marching after alpha_fn
 *                                             (/root/reference/nerf/nerf_helpers_acc.py:11-25)
 *   torch_scatter.scatter_mul(...)         <- /root/reference/nerf/nerf_helpers_acc.py:58
 *
 * nerfacc (PyPI, version NOT pinned by the reference; API surface is 0.3.x) and
 * torch_scatter are absent from /root/reference and from this image, and nerfacc
 * has no CPU implementation.  PARITY UNPINNED: this file restates the library's
 * published algorithm (csrc/ray_marching.cu, csrc/intersection.cu,
 * csrc/render_transmittance.cu of nerfacc 0.3.5) from memory; SURVEY.md section 8c
 * makes this restatement the canonical definition that "bit-exact sample indices
 * and segment offsets" are judged against.
 *
 * Arithmetic contract (mirrored op-for-op by the CUDA kernels):
 *   - every operation is an individually rounded IEEE fp32 op; compile with
 *     -ffp-contract=off so gcc never fuses a*b+c;
 *   - divisions are IEEE divisions; min/max are fminf/fmaxf (NaN-dropping);
 *   - float->int conversion truncates toward zero.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

/* nerfacc csrc/intersection.cu: _ray_aabb_intersect (slab test, miss => 1e10/1e10) */
static void ray_aabb(const float *o, const float *d, const float *aabb, float *near_, float *far_) {
    float tmin = (aabb[0] - o[0]) / d[0];
    float tmax = (aabb[3] - o[0]) / d[0];
    if (tmin > tmax) { float t = tmin; tmin = tmax; tmax = t; }
    float tymin = (aabb[1] - o[1]) / d[1];
    float tymax = (aabb[4] - o[1]) / d[1];
    if (tymin > tymax) { float t = tymin; tymin = tymax; tymax = t; }
    if (tmin > tymax || tymin > tmax) { *near_ = 1e10f; *far_ = 1e10f; return; }
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = (aabb[2] - o[2]) / d[2];
    float tzmax = (aabb[5] - o[2]) / d[2];
    if (tzmin > tzmax) { float t = tzmin; tzmin = tzmax; tzmax = t; }
    if (tmin > tzmax || tzmin > tmax) { *near_ = 1e10f; *far_ = 1e10f; return; }
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    *near_ = tmin;
    *far_ = tmax;
}

/* per-ray [t_min, t_max] after nerfacc.ray_marching's clamp to near_plane / far_plane */
void oracle_ray_aabb_intersect(int64_t n_rays, const float *rays_o, const float *rays_d,
                               const float *aabb, float near_plane, float far_plane,
                               float *t_min, float *t_max) {
    for (int64_t i = 0; i < n_rays; ++i) {
        float a, b;
        ray_aabb(rays_o + 3 * i, rays_d + 3 * i, aabb, &a, &b);
        t_min[i] = a < near_plane ? near_plane : a;   /* torch.clamp(t_min, min=near) */
        t_max[i] = b > far_plane ? far_plane : b;     /* torch.clamp(t_max, max=far)  */
    }
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* nerfacc csrc/ray_marching.cu: grid_occupied_at (ContractionType::AABB) */
static int occupied_at(float x, float y, float z, const float *roi, int res, const uint8_t *binary) {
    if (x < roi[0] || x > roi[3] || y < roi[1] || y > roi[4] || z < roi[2] || z > roi[5]) return 0;
    float ux = (x - roi[0]) / (roi[3] - roi[0]);
    float uy = (y - roi[1]) / (roi[4] - roi[1]);
    float uz = (z - roi[2]) / (roi[5] - roi[2]);
    int ix = clampi((int)(ux * (float)res), 0, res - 1);
    int iy = clampi((int)(uy * (float)res), 0, res - 1);
    int iz = clampi((int)(uz * (float)res), 0, res - 1);
    return binary[((int64_t)ix * res + iy) * res + iz] != 0;
}

/* nerfacc csrc/ray_marching.cu: distance_to_next_voxel, one axis */
static float axis_dist(float p, float d, float inv_d, float lo, float hi, float resf) {
    float u = ((p - lo) / (hi - lo)) * resf;
    float s = copysignf(1.0f, d);
    float t = ((floorf(u + 0.5f + 0.5f * s) - u) * inv_d) / resf * (hi - lo);
    return t;
}

/*
 * March ONE ray.  If t_starts == NULL only counts.  Returns the number of samples.
 * nerfacc csrc/ray_marching.cu: ray_marching_kernel with cone_angle = 0 (dt = step_size).
 */
static int march_one(const float *o, const float *d, float tmin, float tmax, const float *roi,
                     int res, const uint8_t *binary, float dt, float *t_starts, float *t_ends) {
    const float inv0 = 1.0f / d[0], inv1 = 1.0f / d[1], inv2 = 1.0f / d[2];
    const float resf = (float)res;
    int j = 0;
    float t0 = tmin;
    float t1 = t0 + dt;
    float tm = (t0 + t1) * 0.5f;
    while (tm < tmax) {
        float x = o[0] + tm * d[0];
        float y = o[1] + tm * d[1];
        float z = o[2] + tm * d[2];
        if (occupied_at(x, y, z, roi, res, binary)) {
            if (t_starts) { t_starts[j] = t0; t_ends[j] = t1; }
            ++j;
            t0 = t1;
            t1 = t0 + dt;
            tm = (t0 + t1) * 0.5f;
        } else {
            float tx = axis_dist(x, d[0], inv0, roi[0], roi[3], resf);
            float ty = axis_dist(y, d[1], inv1, roi[1], roi[4], resf);
            float tz = axis_dist(z, d[2], inv2, roi[2], roi[5], resf);
            float t = fmaxf(fminf(fminf(tx, ty), tz), 0.0f);
            float target = tm + t;
            float _t = tm;
            do { _t += dt; } while (_t < target);
            tm = _t;
            t0 = tm - dt * 0.5f;
            t1 = tm + dt * 0.5f;
        }
    }
    return j;
}

/* pass 1: per-ray sample counts (nerfacc's "first round") */
void oracle_march_count(int64_t n_rays, const float *rays_o, const float *rays_d,
                        const float *t_min, const float *t_max, const float *roi, int res,
                        const uint8_t *binary, float step_size, int32_t *counts) {
    for (int64_t i = 0; i < n_rays; ++i)
        counts[i] = march_one(rays_o + 3 * i, rays_d + 3 * i, t_min[i], t_max[i], roi, res, binary,
                              step_size, NULL, NULL);
}

/* pass 2: write samples at offsets = exclusive cumsum(counts) (nerfacc's "second round") */
void oracle_march_write(int64_t n_rays, const float *rays_o, const float *rays_d,
                        const float *t_min, const float *t_max, const float *roi, int res,
                        const uint8_t *binary, float step_size, const int64_t *offsets,
                        int64_t *ray_indices, float *t_starts, float *t_ends) {
    for (int64_t i = 0; i < n_rays; ++i) {
        int64_t base = offsets[i];
        int n = march_one(rays_o + 3 * i, rays_d + 3 * i, t_min[i], t_max[i], roi, res, binary,
                          step_size, t_starts + base, t_ends + base);
        for (int j = 0; j < n; ++j) ray_indices[base + j] = i;
    }
}

/* nerfacc OccupancyGrid.query_occ -> _C.grid_query: occupancy (0/1) at arbitrary points */
void oracle_grid_query(int64_t n, const float *pts, const float *roi, int res, const uint8_t *binary,
                       float *out) {
    for (int64_t i = 0; i < n; ++i)
        out[i] = (float)occupied_at(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], roi, res, binary);
}

/*
 * nerfacc render_visibility (0.3.5): per ray, T_0 = 1, T_{i+1} = T_i * (1 - alpha_i) accumulated
 * sequentially in fp32 over ALL samples; sample i is kept iff T_i >= early_stop_eps and
 * (alpha_thre <= 0 or alpha_i >= alpha_thre).  offsets has n_rays+1 entries.
 */
void oracle_visibility(int64_t n_rays, const int64_t *offsets, const float *alphas,
                       float early_stop_eps, float alpha_thre, uint8_t *keep) {
    for (int64_t r = 0; r < n_rays; ++r) {
        float T = 1.0f;
        for (int64_t i = offsets[r]; i < offsets[r + 1]; ++i) {
            float a = alphas[i];
            int vis = T >= early_stop_eps;
            if (alpha_thre > 0.0f) vis = vis && (a >= alpha_thre);
            keep[i] = (uint8_t)vis;
            T *= (1.0f - a);
        }
    }
}

/*
 * torch_scatter.scatter_mul(alphas, index, dim=0, out=ones[n_rays]) restated with a fixed
 * (ascending sample) multiplication order -- the library's own order is non-deterministic.
 * /root/reference/nerf/nerf_helpers_acc.py:53-58
 */
void oracle_scatter_mul(int64_t n, const float *src, const int64_t *index, int64_t n_rays, float *out) {
    for (int64_t r = 0; r < n_rays; ++r) out[r] = 1.0f;
    for (int64_t i = 0; i < n; ++i) out[index[i]] *= src[i];
}

#ifdef __cplusplus
}
#endif
