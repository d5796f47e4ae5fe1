// raygen.cu -- cone-beam ray generation from the C-arm projection geometry.
// Replaces get_ray_values (/root/reference/phantomdata/helpers.py:156-175): the reference builds rays in
// float64 on the host, writes them to CSV and casts to float32 at training time
// (/root/reference/nerf/run_nerf_acc.py:88-89).  Here rays are generated on the fly from (view, x, y):
// float64 arithmetic in the reference's operation order (no FMA contraction), one rounding to fp32.
// HBM-bound: 24 B/ray written (+12 B/ray of indices read in gather mode).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) raygen_kernel(const double* __restrict__ cam2world, int view0,
                                                     const int32_t* __restrict__ view_ids,
                                                     const int32_t* __restrict__ px, const int32_t* __restrict__ py,
                                                     int64_t n, int img_w, int img_h, double focal,
                                                     const float* __restrict__ pixels, float* __restrict__ rays_o,
                                                     float* __restrict__ rays_d, float* __restrict__ pix_out) {
  const double half_w = (double)img_w / 2.0, half_h = (double)img_h / 2.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int v, x, y;
    if (view_ids) {
      v = view_ids[i]; x = px[i]; y = py[i];
    } else {
      v = view0; x = (int)(i % img_w); y = (int)(i / img_w);
    }
    const double* M = cam2world + (int64_t)v * 16;
    // directions = [(ii - W/2)/f, -(jj - H/2)/f, -1]
    const double d0 = __ddiv_rn(__dsub_rn((double)x, half_w), focal);
    const double d1 = -__ddiv_rn(__dsub_rn((double)y, half_h), focal);
    const double d2 = -1.0;
    float o[3], d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      // torch.sum(directions[..., None, :] * M[:3, :3], dim=-1): products added left to right
      double s = __dadd_rn(__dadd_rn(__dmul_rn(d0, M[k * 4 + 0]), __dmul_rn(d1, M[k * 4 + 1])), __dmul_rn(d2, M[k * 4 + 2]));
      d[k] = (float)s;
      o[k] = (float)M[k * 4 + 3];
    }
    rays_o[i * 3 + 0] = o[0]; rays_o[i * 3 + 1] = o[1]; rays_o[i * 3 + 2] = o[2];
    rays_d[i * 3 + 0] = d[0]; rays_d[i * 3 + 1] = d[1]; rays_d[i * 3 + 2] = d[2];
    if (pix_out) pix_out[i] = pixels[((int64_t)v * img_h + y) * img_w + x];
  }
}

}  // namespace

extern "C" int angio_raygen(const double* cam2world, int32_t view0, const int32_t* view_ids, const int32_t* px,
                            const int32_t* py, int64_t n, int32_t img_w, int32_t img_h, double focal,
                            const float* pixels, float* rays_o, float* rays_d, float* pix_out, void* stream) {
  ANGIO_REQUIRE(cam2world && rays_o && rays_d, "angio_raygen: null pointer");
  ANGIO_REQUIRE(n >= 0 && img_w > 0 && img_h > 0 && focal != 0.0, "angio_raygen: bad sizes");
  ANGIO_REQUIRE((view_ids == nullptr) == (px == nullptr) && (px == nullptr) == (py == nullptr),
                "angio_raygen: view_ids/px/py must be all NULL (image mode) or all set (gather mode)");
  ANGIO_REQUIRE(!view_ids ? n <= (int64_t)img_w * img_h : true, "angio_raygen: image mode n > W*H");
  ANGIO_REQUIRE((pix_out == nullptr) || (pixels != nullptr), "angio_raygen: pix_out needs pixels");
  if (n == 0) return 0;
  int blocks = angio::blocks_for(n, 256);
  int cap = angio::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  angio::note_launch("raygen_kernel"); raygen_kernel<<<blocks, 256, 0, angio::as_stream(stream)>>>(cam2world, view0, view_ids, px, py, n, img_w, img_h, focal,
                                                             pixels, rays_o, rays_d, pix_out);
  return angio::finish_launch("angio_raygen");
}
