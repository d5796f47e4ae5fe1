// mlp_api.cu -- C-ABI entry points of the CPPN MLP; dispatch between the fp32 check path (mlp_simt.cu) and the
// bf16 tcgen05 path (mlp_tc.cu).  No silent fallback: an unsupported (shape, precision) pair is an error.
#include "mlp_layout.cuh"

namespace angio {
int64_t simt_workspace_bytes(const MlpLayout& L, int64_t n, int training);
int64_t simt_saved_bytes(const MlpLayout& L, int64_t n);
int simt_forward(const MlpLayout& L, const float* params, const angio_samples& in, int out_mode, float* out, void* saved,
                 void* workspace, int64_t workspace_bytes, cudaStream_t st);
int simt_backward(const MlpLayout& L, const float* params, const angio_samples& in, const void* saved, const float* grad_out,
                  float* grad_params, void* workspace, int64_t workspace_bytes, cudaStream_t st);

bool tc_supported(const MlpLayout& L);
int64_t tc_packed_bytes(const MlpLayout& L);
int64_t tc_workspace_bytes(const MlpLayout& L, int64_t n, int training);
int64_t tc_saved_bytes(const MlpLayout& L, int64_t n);
int tc_pack_weights(const MlpLayout& L, const float* params, void* packed, cudaStream_t st);
int tc_forward(const MlpLayout& L, const float* params, const void* packed, const angio_samples& in, int out_mode, float* out,
               void* saved, void* workspace, int64_t workspace_bytes, cudaStream_t st);
int tc_backward(const MlpLayout& L, const float* params, const void* packed, const angio_samples& in, const void* saved,
                const float* grad_out, float* grad_params, void* workspace, int64_t workspace_bytes, cudaStream_t st);
// width 256 (mlp_tc256.cu): weights streamed through shared memory
bool tc256_supported(const MlpLayout& L);
int64_t tc256_packed_bytes(const MlpLayout& L);
int64_t tc256_workspace_bytes(const MlpLayout& L, int64_t n, int training);
int64_t tc256_saved_bytes(const MlpLayout& L, int64_t n);
int tc256_pack_weights(const MlpLayout& L, const float* params, void* packed, cudaStream_t st);
int tc256_forward(const MlpLayout& L, const void* packed, const angio_samples& in, int out_mode, float* out, void* saved, cudaStream_t st);
int tc256_backward(const MlpLayout& L, const void* packed, const angio_samples& in, const void* saved, const float* grad_out,
                   float* grad_params, void* workspace, int64_t workspace_bytes, cudaStream_t st);
}  // namespace angio

using angio::MlpLayout;

static int check_samples(const angio_samples* in, const char* who) {
  if (!in || in->n < 0) { angio::set_error("%s: bad sample descriptor", who); return ANGIO_ERR_INVALID_ARG; }
  if (in->n > 0 && !in->points && !(in->rays_o && in->rays_d && in->ray_idx && in->t_starts && in->t_ends)) {
    angio::set_error("%s: need either points or (rays_o, rays_d, ray_idx, t_starts, t_ends)", who);
    return ANGIO_ERR_INVALID_ARG;
  }
  return 0;
}

extern "C" int64_t angio_mlp_param_count(const angio_mlp_desc* desc) {
  MlpLayout L;
  if (!angio::make_layout(desc, &L)) { angio::set_error("angio_mlp_param_count: invalid descriptor"); return ANGIO_ERR_INVALID_ARG; }
  return L.total;
}
extern "C" int32_t angio_mlp_input_width(const angio_mlp_desc* desc) {
  MlpLayout L;
  if (!angio::make_layout(desc, &L)) { angio::set_error("angio_mlp_input_width: invalid descriptor"); return ANGIO_ERR_INVALID_ARG; }
  return L.d_in;
}
extern "C" int64_t angio_mlp_workspace_bytes(const angio_mlp_desc* desc, int64_t n, int32_t precision, int32_t training) {
  MlpLayout L;
  if (!angio::make_layout(desc, &L) || n < 0) { angio::set_error("angio_mlp_workspace_bytes: invalid arguments"); return ANGIO_ERR_INVALID_ARG; }
  if (precision == ANGIO_PREC_FP32) return angio::simt_workspace_bytes(L, n, training);
  if (precision == ANGIO_PREC_BF16 && angio::tc_supported(L)) return angio::tc_workspace_bytes(L, n, training);
  if (precision == ANGIO_PREC_BF16 && angio::tc256_supported(L)) return angio::tc256_workspace_bytes(L, n, training);
  angio::set_error("angio_mlp_workspace_bytes: unsupported precision/shape");
  return ANGIO_ERR_UNSUPPORTED;
}
extern "C" int64_t angio_mlp_saved_bytes(const angio_mlp_desc* desc, int64_t n, int32_t precision) {
  MlpLayout L;
  if (!angio::make_layout(desc, &L) || n < 0) { angio::set_error("angio_mlp_saved_bytes: invalid arguments"); return ANGIO_ERR_INVALID_ARG; }
  if (precision == ANGIO_PREC_FP32) return angio::simt_saved_bytes(L, n);
  if (precision == ANGIO_PREC_BF16 && angio::tc_supported(L)) return angio::tc_saved_bytes(L, n);
  if (precision == ANGIO_PREC_BF16 && angio::tc256_supported(L)) return angio::tc256_saved_bytes(L, n);
  angio::set_error("angio_mlp_saved_bytes: unsupported precision/shape");
  return ANGIO_ERR_UNSUPPORTED;
}
extern "C" int64_t angio_mlp_packed_bytes(const angio_mlp_desc* desc) {
  MlpLayout L;
  if (!angio::make_layout(desc, &L)) { angio::set_error("angio_mlp_packed_bytes: invalid descriptor"); return ANGIO_ERR_INVALID_ARG; }
  if (angio::tc256_supported(L)) return angio::tc256_packed_bytes(L);
  if (!angio::tc_supported(L)) { angio::set_error("angio_mlp_packed_bytes: shape not supported by the bf16 path"); return ANGIO_ERR_UNSUPPORTED; }
  return angio::tc_packed_bytes(L);
}
extern "C" int angio_mlp_pack_weights(const angio_mlp_desc* desc, const float* params, void* packed, void* stream) {
  MlpLayout L;
  ANGIO_REQUIRE(angio::make_layout(desc, &L) && params && packed, "angio_mlp_pack_weights: bad arguments");
  if (angio::tc256_supported(L)) return angio::tc256_pack_weights(L, params, packed, angio::as_stream(stream));
  if (!angio::tc_supported(L)) { angio::set_error("angio_mlp_pack_weights: shape not supported by the bf16 path"); return ANGIO_ERR_UNSUPPORTED; }
  return angio::tc_pack_weights(L, params, packed, angio::as_stream(stream));
}

extern "C" int angio_mlp_forward(const angio_mlp_desc* desc, const float* params, const void* packed, const angio_samples* in,
                                 int32_t out_mode, int32_t precision, float* out, void* saved, void* workspace,
                                 int64_t workspace_bytes, void* stream) {
  MlpLayout L;
  ANGIO_REQUIRE(angio::make_layout(desc, &L), "angio_mlp_forward: invalid descriptor");
  ANGIO_REQUIRE(params && (out || (in && in->n == 0)), "angio_mlp_forward: null pointer");
  ANGIO_REQUIRE(out_mode >= ANGIO_OUT_LOGIT && out_mode <= ANGIO_OUT_ALPHA, "angio_mlp_forward: bad out_mode");
  if (int rc = check_samples(in, "angio_mlp_forward")) return rc;
  ANGIO_REQUIRE(!(out_mode == ANGIO_OUT_ALPHA && in->n > 0 && !(in->t_starts && in->t_ends)), "angio_mlp_forward: ALPHA output needs t_starts/t_ends");
  if (in->n_dev && precision != ANGIO_PREC_BF16) {
    angio::set_error("angio_mlp_forward: a device-resident sample count (n_dev) is only supported by the bf16 path");
    return ANGIO_ERR_UNSUPPORTED;
  }
  if (in->sample_idx && (precision != ANGIO_PREC_BF16 || saved || in->points)) {
    angio::set_error("angio_mlp_forward: sample_idx is only supported by the bf16 inference forward on ray samples");
    return ANGIO_ERR_UNSUPPORTED;
  }
  if (precision == ANGIO_PREC_FP32)
    return angio::simt_forward(L, params, *in, out_mode, out, saved, workspace, workspace_bytes, angio::as_stream(stream));
  if (precision == ANGIO_PREC_BF16) {
    ANGIO_REQUIRE(packed, "angio_mlp_forward: bf16 path needs the packed weight image (angio_mlp_pack_weights)");
    if (angio::tc256_supported(L)) return angio::tc256_forward(L, packed, *in, out_mode, out, saved, angio::as_stream(stream));
    if (!angio::tc_supported(L)) { angio::set_error("angio_mlp_forward: shape not supported by the bf16 tcgen05 path"); return ANGIO_ERR_UNSUPPORTED; }
    return angio::tc_forward(L, params, packed, *in, out_mode, out, saved, workspace, workspace_bytes, angio::as_stream(stream));
  }
  angio::set_error("angio_mlp_forward: unknown precision %d", precision);
  return ANGIO_ERR_INVALID_ARG;
}

extern "C" int angio_mlp_backward(const angio_mlp_desc* desc, const float* params, const void* packed, const angio_samples* in,
                                  const void* saved, const float* grad_out, int32_t precision, float* grad_params,
                                  void* workspace, int64_t workspace_bytes, void* stream) {
  MlpLayout L;
  ANGIO_REQUIRE(angio::make_layout(desc, &L), "angio_mlp_backward: invalid descriptor");
  ANGIO_REQUIRE(params && grad_params, "angio_mlp_backward: null pointer");
  if (int rc = check_samples(in, "angio_mlp_backward")) return rc;
  ANGIO_REQUIRE(in->n == 0 || (saved && grad_out), "angio_mlp_backward: needs saved activations and grad_out");
  ANGIO_REQUIRE(!in->n_dev || precision == ANGIO_PREC_BF16, "angio_mlp_backward: n_dev is only supported by the bf16 path");
  ANGIO_REQUIRE(!in->sample_idx, "angio_mlp_backward: sample_idx is not supported");
  if (precision == ANGIO_PREC_FP32)
    return angio::simt_backward(L, params, *in, saved, grad_out, grad_params, workspace, workspace_bytes, angio::as_stream(stream));
  if (precision == ANGIO_PREC_BF16) {
    ANGIO_REQUIRE(packed, "angio_mlp_backward: bf16 path needs the packed weight image");
    if (angio::tc256_supported(L))
      return angio::tc256_backward(L, packed, *in, saved, grad_out, grad_params, workspace, workspace_bytes, angio::as_stream(stream));
    if (!angio::tc_supported(L)) { angio::set_error("angio_mlp_backward: shape not supported by the bf16 tcgen05 path"); return ANGIO_ERR_UNSUPPORTED; }
    return angio::tc_backward(L, params, packed, *in, saved, grad_out, grad_params, workspace, workspace_bytes, angio::as_stream(stream));
  }
  angio::set_error("angio_mlp_backward: unknown precision %d", precision);
  return ANGIO_ERR_INVALID_ARG;
}
