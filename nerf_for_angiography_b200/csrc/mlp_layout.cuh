// mlp_layout.cuh -- flat parameter layout of the CPPN MLP shared by the fp32 and bf16 paths.
// Mirrors the construction order of /root/reference/model/CPPN.py:96-131 (input layer, n_hidden HxH layers,
// output layer) with torch.nn.Linear's [out][in] row-major weights.
#pragma once
#include "common.cuh"

namespace angio {

constexpr int kMaxLinear = 18;  // input layer + up to 16 hidden + output

struct MlpLayout {
  int enc, basis, d_in, H, n_hidden, n_linear;  // n_linear = n_hidden + 2
  int64_t off_coef;                             // -1 if no encoding
  int64_t off_w[kMaxLinear], off_b[kMaxLinear];
  int in_dim[kMaxLinear], out_dim[kMaxLinear];
  int64_t total;
};

inline bool make_layout(const angio_mlp_desc* d, MlpLayout* L) {
  if (!d || d->width <= 0 || d->n_hidden < 0 || d->n_hidden + 2 > kMaxLinear) return false;
  if (d->enc != 0 && d->enc != 1) return false;
  if (d->enc == 1 && d->enc_basis <= 0) return false;
  L->enc = d->enc;
  L->basis = d->enc ? d->enc_basis : 0;
  L->d_in = 3 + 6 * L->basis;
  L->H = d->width;
  L->n_hidden = d->n_hidden;
  L->n_linear = d->n_hidden + 2;
  int64_t off = 0;
  L->off_coef = -1;
  if (L->enc) { L->off_coef = 0; off += 3 * L->basis; }
  for (int l = 0; l < L->n_linear; ++l) {
    L->in_dim[l] = (l == 0) ? L->d_in : L->H;
    L->out_dim[l] = (l == L->n_linear - 1) ? 1 : L->H;
    L->off_w[l] = off; off += (int64_t)L->in_dim[l] * L->out_dim[l];
    L->off_b[l] = off; off += L->out_dim[l];
  }
  L->total = off;
  return true;
}

// midpoint of a ray sample, reference operation order: o + (d * (t0 + t1)) / 2   (run_nerf_acc.py:290-292)
__device__ __forceinline__ void sample_position(const angio_samples& in, int64_t i, float x[3]) {
  if (in.points) {
    x[0] = in.points[i * 3]; x[1] = in.points[i * 3 + 1]; x[2] = in.points[i * 3 + 2];
  } else {
    const int r = in.ray_idx[i];
    const float ts = __fadd_rn(in.t_starts[i], in.t_ends[i]);
#pragma unroll
    for (int k = 0; k < 3; ++k)
      x[k] = __fadd_rn(in.rays_o[(int64_t)r * 3 + k], __fmul_rn(__fmul_rn(in.rays_d[(int64_t)r * 3 + k], ts), 0.5f));
  }
}

}  // namespace angio
