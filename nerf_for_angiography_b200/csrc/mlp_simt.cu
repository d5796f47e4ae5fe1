// mlp_simt.cu -- fp32 "check mode" of the CPPN MLP (ANGIO_PREC_FP32): the reference's arithmetic
// (/root/reference/model/CPPN.py:166-222) layer by layer on CUDA cores, used for the <=1e-5 parity mode and
// for shapes the tensor-core kernel does not cover.  Encoding follows the reference op order exactly:
// a = fl(fl(6.2831855f * x) * coeff), accurate sinf/cosf on the unreduced argument.
//
// Structure: encode -> (L+1) x [SGEMM + bias + ReLU] -> output dot product; backward = dgrad SGEMMs with the
// ReLU mask fused, split-K weight-gradient SGEMMs with a fixed-order (deterministic) reduction, and the
// Fourier-coefficient gradient (the coefficients are a learned nn.Parameter, model/CPPN.py:73-75).
// The throughput path is mlp_tc.cu (bf16 tcgen05); this file favours exactness and generality.
#include "mlp_layout.cuh"

namespace {

using angio::MlpLayout;

constexpr float kTwoPi = 6.2831855f;  // float32(2*np.pi)

// ------------------------------------------------------------------------------------------- encoding
__global__ void __launch_bounds__(256) encode_kernel(angio_samples in, int64_t i0, int64_t n, int basis,
                                                     const float* __restrict__ coef, float* __restrict__ X, int ld) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x[3];
  angio::sample_position(in, i0 + i, x);
  float* row = X + i * ld;
  row[0] = x[0]; row[1] = x[1]; row[2] = x[2];
  const int nb = 3 * basis;
  for (int j = 0; j < nb; ++j) {
    const float a = __fmul_rn(__fmul_rn(kTwoPi, x[j % 3]), coef[j]);
    row[3 + j] = sinf(a);
    row[3 + nb + j] = cosf(a);
  }
}

// ------------------------------------------------------------------------------------------- generic tiled SGEMM
// C[M x N] = A[M x K] * B[K x N];  A_T: A stored [K][M] (lda = row stride of the stored matrix), B_T: B stored [N][K].
constexpr int BM = 64, BN = 64, BK = 16;

struct EpiBiasRelu {  // forward: y = relu?(acc + bias[n])
  const float* bias; float* C; int ldc; int relu;
  __device__ __forceinline__ void operator()(int m, int n, float acc, int) const {
    float v = acc + bias[n];
    if (relu) v = fmaxf(v, 0.0f);
    C[(int64_t)m * ldc + n] = v;
  }
};
struct EpiMask {  // dgrad: dx = acc * (act > 0)
  const float* act; int ldact; float* C; int ldc;
  __device__ __forceinline__ void operator()(int m, int n, float acc, int) const {
    C[(int64_t)m * ldc + n] = act[(int64_t)m * ldact + n] > 0.0f ? acc : 0.0f;
  }
};
struct EpiStore {  // plain store
  float* C; int ldc;
  __device__ __forceinline__ void operator()(int m, int n, float acc, int) const { C[(int64_t)m * ldc + n] = acc; }
};
struct EpiPartial {  // split-K partial: partial[z][m][n]
  float* P; int M, N;
  __device__ __forceinline__ void operator()(int m, int n, float acc, int z) const { P[((int64_t)z * M + m) * N + n] = acc; }
};

template <bool A_T, bool B_T, class Epi>
__global__ void __launch_bounds__(256) sgemm_kernel(int M, int N, int64_t K, const float* __restrict__ A, int lda,
                                                    const float* __restrict__ B, int ldb, int64_t k_per_split, Epi epi) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int64_t kb = (int64_t)blockIdx.z * k_per_split;
  const int64_t ke = (kb + k_per_split < K) ? kb + k_per_split : K;
  float acc[4][4] = {};
  for (int64_t k0 = kb; k0 < ke; k0 += BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      int m, k;
      if (A_T) { m = idx % BM; k = idx / BM; } else { k = idx % BK; m = idx / BK; }
      const int gm = m0 + m; const int64_t gk = k0 + k;
      float v = 0.0f;
      if (gm < M && gk < ke) v = A_T ? A[gk * lda + gm] : A[(int64_t)gm * lda + gk];
      As[k][m] = v;
      int n, k2;
      if (B_T) { k2 = idx % BK; n = idx / BK; } else { n = idx % BN; k2 = idx / BN; }
      const int gn = n0 + n; const int64_t gk2 = k0 + k2;
      float w = 0.0f;
      if (gn < N && gk2 < ke) w = B_T ? B[(int64_t)gn * ldb + gk2] : B[gk2 * ldb + gn];
      Bs[k2][n] = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) epi(m, n, acc[i][j], blockIdx.z);
    }
}

template <bool A_T, bool B_T, class Epi>
int launch_sgemm(int M, int N, int64_t K, const float* A, int lda, const float* B, int ldb, int splits, Epi epi, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  const int64_t kps = ((K + splits - 1) / splits + BK - 1) / BK * BK;
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, splits);
  angio::note_launch("sgemm_kernel<A_T, B_T, Epi>"); sgemm_kernel<A_T, B_T, Epi><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, kps > 0 ? kps : BK, epi);
  return angio::finish_launch("sgemm");
}

// ------------------------------------------------------------------------------------------- output layer
// warp per sample: out = transform(dot(a[H], w) + b)
__global__ void __launch_bounds__(256) out_dot_kernel(const float* __restrict__ act, int H, int64_t n, const float* __restrict__ w,
                                                      const float* __restrict__ b, angio_samples in, int64_t i0, int out_mode,
                                                      float* __restrict__ out) {
  const int lane = threadIdx.x % 32;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  if (row >= n) return;
  float acc = 0.0f;
  for (int k = lane; k < H; k += 32) acc = fmaf(act[row * H + k], w[k], acc);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) {
    float v = acc + b[0];
    if (out_mode != ANGIO_OUT_LOGIT) {
      const float s = angio::sigmoidf_ref(v);
      if (out_mode == ANGIO_OUT_SIGMA) v = s;
      else v = 1.0f - expf(-s * (in.t_ends[i0 + row] - in.t_starts[i0 + row]));
    }
    out[i0 + row] = v;
  }
}

// delta of the last hidden activation: dz[s][o] = g[s] * w_out[o] * (a[s][o] > 0)
__global__ void __launch_bounds__(256) dout_kernel(const float* __restrict__ g, const float* __restrict__ w, const float* __restrict__ act,
                                                   int H, int64_t n, float* __restrict__ dz) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * H) return;
  const int64_t s = idx / H; const int o = (int)(idx % H);
  dz[idx] = act[idx] > 0.0f ? g[s] * w[o] : 0.0f;
}

// ------------------------------------------------------------------------------------------- deterministic reductions
// column sums of Y[n x N] (bias gradients; also grad of w_out via Y = g[s]*a[s][:]) as split partials
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ Y, const float* __restrict__ scale, int N,
                                                             int64_t n, int64_t rows_per_split, float* __restrict__ partial) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= N) return;
  const int64_t rb = (int64_t)blockIdx.y * rows_per_split;
  const int64_t re = rb + rows_per_split < n ? rb + rows_per_split : n;
  float acc = 0.0f;
  for (int64_t r = rb; r < re; ++r) acc += scale ? Y[r * N + col] * scale[r] : Y[r * N + col];
  partial[(int64_t)blockIdx.y * N + col] = acc;
}
__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ partial, int splits, int64_t len, float* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= len) return;
  float acc = 0.0f;
  for (int z = 0; z < splits; ++z) acc += partial[(int64_t)z * len + i];
  out[i] = acc;
}
__global__ void __launch_bounds__(256) vecsum_partial_kernel(const float* __restrict__ g, int64_t n, int64_t per_split, float* __restrict__ partial) {
  __shared__ float s[256];
  const int64_t rb = (int64_t)blockIdx.x * per_split;
  const int64_t re = rb + per_split < n ? rb + per_split : n;
  float acc = 0.0f;
  for (int64_t r = rb + threadIdx.x; r < re; r += 256) acc += g[r];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) s[threadIdx.x] += s[threadIdx.x + k]; __syncthreads(); }
  if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}

// Fourier-coefficient gradient partials: dcoef_j = sum_s (dX[s][3+j]*cos(a) - dX[s][3+nb+j]*sin(a)) * fl(2pi*x_c)
__global__ void __launch_bounds__(256) coef_grad_partial_kernel(const float* __restrict__ X0, const float* __restrict__ dX0, int ld,
                                                                int basis, int64_t n, int64_t rows_per_split,
                                                                float* __restrict__ partial) {
  __shared__ float s[256];
  const int nb = 3 * basis;
  const int j = blockIdx.x;  // coefficient index
  const int64_t rb = (int64_t)blockIdx.y * rows_per_split;
  const int64_t re = rb + rows_per_split < n ? rb + rows_per_split : n;
  float acc = 0.0f;
  for (int64_t r = rb + threadIdx.x; r < re; r += 256) {
    const float* x = X0 + r * ld;
    const float* d = dX0 + r * ld;
    // sin(a) and cos(a) are the stored features; d a / d coef = fl(2pi * x_c)
    acc += (d[3 + j] * x[3 + nb + j] - d[3 + nb + j] * x[3 + j]) * __fmul_rn(kTwoPi, x[j % 3]);
  }
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) s[threadIdx.x] += s[threadIdx.x + k]; __syncthreads(); }
  if (threadIdx.x == 0) partial[(int64_t)blockIdx.y * nb + j] = s[0];
}

constexpr int64_t kChunk = 131072;  // inference chunk (the reference's batch_size, run_nerf_acc.py:146)
constexpr int kSplits = 128;        // split-K factor of the weight-gradient reductions

inline int64_t align256(int64_t b) { return (b + 255) / 256 * 256; }

}  // namespace

namespace angio {

int64_t simt_workspace_bytes(const MlpLayout& L, int64_t n, int training) {
  if (!training) {
    const int64_t c = n < kChunk ? n : kChunk;
    return align256(c * L.d_in * 4) + 2 * align256(c * L.H * 4) + 256;
  }
  // backward: two delta buffers [n x H], dX0 [n x d_in], split partials
  const int64_t part = (int64_t)kSplits * ((int64_t)L.H * (L.H > L.d_in ? L.H : L.d_in) + L.H);
  return 2 * align256(n * L.H * 4) + align256(n * L.d_in * 4) + align256(part * 4) + 256;
}

int64_t simt_saved_bytes(const MlpLayout& L, int64_t n) {
  return align256(n * L.d_in * 4) + (int64_t)(L.n_hidden + 1) * align256(n * L.H * 4);
}

// forward over samples [0, n).  saved != NULL: keep X0 and every hidden activation for the backward.
int simt_forward(const MlpLayout& L, const float* params, const angio_samples& in, int out_mode, float* out, void* saved,
                 void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  const int64_t n = in.n;
  if (n == 0) return 0;
  const float* coef = L.enc ? params + L.off_coef : nullptr;
  if (saved) {
    char* base = reinterpret_cast<char*>(saved);
    float* X0 = reinterpret_cast<float*>(base);
    base += align256(n * L.d_in * 4);
    angio::note_launch("encode_kernel"); encode_kernel<<<blocks_for(n, 256), 256, 0, st>>>(in, 0, n, L.basis, coef, X0, L.d_in);
    const float* cur = X0; int cur_ld = L.d_in;
    for (int l = 0; l <= L.n_hidden; ++l) {
      float* Y = reinterpret_cast<float*>(base);
      base += align256(n * L.H * 4);
      EpiBiasRelu epi{params + L.off_b[l], Y, L.H, 1};
      int rc = launch_sgemm<false, true>((int)n, L.H, L.in_dim[l], cur, cur_ld, params + L.off_w[l], L.in_dim[l], 1, epi, st);
      if (rc) return rc;
      cur = Y; cur_ld = L.H;
    }
    const int lo = L.n_linear - 1;
    angio::note_launch("out_dot_kernel"); out_dot_kernel<<<blocks_for(n * 32, 256), 256, 0, st>>>(cur, L.H, n, params + L.off_w[lo], params + L.off_b[lo], in, 0, out_mode, out);
    return finish_launch("simt_forward(train)");
  }
  const int64_t c = n < kChunk ? n : kChunk;
  const int64_t need = align256(c * L.d_in * 4) + 2 * align256(c * L.H * 4);
  if (!workspace || workspace_bytes < need) {
    set_error("angio_mlp_forward(fp32): workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need);
    return ANGIO_ERR_WORKSPACE;
  }
  char* base = reinterpret_cast<char*>(workspace);
  float* X0 = reinterpret_cast<float*>(base);
  float* bufs[2] = {reinterpret_cast<float*>(base + align256(c * L.d_in * 4)),
                    reinterpret_cast<float*>(base + align256(c * L.d_in * 4) + align256(c * L.H * 4))};
  for (int64_t i0 = 0; i0 < n; i0 += c) {
    const int64_t m = (n - i0 < c) ? n - i0 : c;
    angio::note_launch("encode_kernel"); encode_kernel<<<blocks_for(m, 256), 256, 0, st>>>(in, i0, m, L.basis, coef, X0, L.d_in);
    const float* cur = X0; int cur_ld = L.d_in;
    for (int l = 0; l <= L.n_hidden; ++l) {
      float* Y = bufs[l & 1];
      EpiBiasRelu epi{params + L.off_b[l], Y, L.H, 1};
      int rc = launch_sgemm<false, true>((int)m, L.H, L.in_dim[l], cur, cur_ld, params + L.off_w[l], L.in_dim[l], 1, epi, st);
      if (rc) return rc;
      cur = Y; cur_ld = L.H;
    }
    const int lo = L.n_linear - 1;
    angio::note_launch("out_dot_kernel"); out_dot_kernel<<<blocks_for(m * 32, 256), 256, 0, st>>>(cur, L.H, m, params + L.off_w[lo], params + L.off_b[lo], in, i0, out_mode, out);
  }
  return finish_launch("simt_forward");
}

int simt_backward(const MlpLayout& L, const float* params, const angio_samples& in, const void* saved, const float* grad_out,
                  float* grad_params, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  const int64_t n = in.n;
  cudaError_t ce = cudaMemsetAsync(grad_params, 0, L.total * 4, st);
  if (ce != cudaSuccess) { set_error("memset grad_params: %s", cudaGetErrorString(ce)); return (int)ce; }
  if (n == 0) return 0;
  const int64_t need = simt_workspace_bytes(L, n, 1) - 256;
  if (!workspace || workspace_bytes < need) {
    set_error("angio_mlp_backward(fp32): workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need);
    return ANGIO_ERR_WORKSPACE;
  }
  // saved activations
  const char* sb = reinterpret_cast<const char*>(saved);
  const float* X0 = reinterpret_cast<const float*>(sb);
  sb += align256(n * L.d_in * 4);
  const float* act[kMaxLinear];
  for (int l = 0; l <= L.n_hidden; ++l) { act[l] = reinterpret_cast<const float*>(sb); sb += align256(n * L.H * 4); }
  // workspace
  char* wb = reinterpret_cast<char*>(workspace);
  float* dbuf[2] = {reinterpret_cast<float*>(wb), reinterpret_cast<float*>(wb + align256(n * L.H * 4))};
  wb += 2 * align256(n * L.H * 4);
  float* dX0 = reinterpret_cast<float*>(wb);
  wb += align256(n * L.d_in * 4);
  float* partial = reinterpret_cast<float*>(wb);

  const int H = L.H;
  const int lo = L.n_linear - 1;
  const int64_t rows_per_split = (n + kSplits - 1) / kSplits;
  // output layer: d b_out = sum g ; d w_out[o] = sum_s g[s] * a_last[s][o]
  angio::note_launch("vecsum_partial_kernel"); vecsum_partial_kernel<<<kSplits, 256, 0, st>>>(grad_out, n, rows_per_split, partial);
  angio::note_launch("sum_partials_kernel"); sum_partials_kernel<<<1, 256, 0, st>>>(partial, kSplits, 1, grad_params + L.off_b[lo]);
  angio::note_launch("colsum_partial_kernel"); colsum_partial_kernel<<<dim3((H + 255) / 256, kSplits), 256, 0, st>>>(act[L.n_hidden], grad_out, H, n, rows_per_split, partial);
  angio::note_launch("sum_partials_kernel"); sum_partials_kernel<<<blocks_for(H, 256), 256, 0, st>>>(partial, kSplits, H, grad_params + L.off_w[lo]);
  // delta of the last hidden layer
  float* dz = dbuf[0];
  angio::note_launch("dout_kernel"); dout_kernel<<<blocks_for(n * H, 256), 256, 0, st>>>(grad_out, params + L.off_w[lo], act[L.n_hidden], H, n, dz);
  for (int l = L.n_hidden; l >= 0; --l) {
    const float* a_in = (l == 0) ? X0 : act[l - 1];
    const int K_in = L.in_dim[l];
    // weight gradient dW_l[o][i] = sum_s dz[s][o] * a_in[s][i]   (split over samples, fixed-order reduction)
    {
      EpiPartial epi{partial, H, K_in};
      int rc = launch_sgemm<true, false>(H, K_in, n, dz, H, a_in, K_in, kSplits, epi, st);
      if (rc) return rc;
      angio::note_launch("sum_partials_kernel"); sum_partials_kernel<<<blocks_for((int64_t)H * K_in, 256), 256, 0, st>>>(partial, kSplits, (int64_t)H * K_in, grad_params + L.off_w[l]);
    }
    angio::note_launch("colsum_partial_kernel"); colsum_partial_kernel<<<dim3((H + 255) / 256, kSplits), 256, 0, st>>>(dz, nullptr, H, n, rows_per_split, partial);
    angio::note_launch("sum_partials_kernel"); sum_partials_kernel<<<blocks_for(H, 256), 256, 0, st>>>(partial, kSplits, H, grad_params + L.off_b[l]);
    // data gradient
    if (l > 0) {
      float* dnext = dbuf[(dz == dbuf[0]) ? 1 : 0];
      EpiMask epi{act[l - 1], H, dnext, H};
      int rc = launch_sgemm<false, false>((int)n, H, H, dz, H, params + L.off_w[l], H, 1, epi, st);
      if (rc) return rc;
      dz = dnext;
    } else if (L.enc) {
      EpiStore epi{dX0, L.d_in};
      int rc = launch_sgemm<false, false>((int)n, L.d_in, H, dz, H, params + L.off_w[0], L.d_in, 1, epi, st);
      if (rc) return rc;
      const int nb = 3 * L.basis;
      angio::note_launch("coef_grad_partial_kernel"); coef_grad_partial_kernel<<<dim3(nb, kSplits), 256, 0, st>>>(X0, dX0, L.d_in, L.basis, n, rows_per_split, partial);
      angio::note_launch("sum_partials_kernel"); sum_partials_kernel<<<blocks_for(nb, 256), 256, 0, st>>>(partial, kSplits, nb, grad_params + L.off_coef);
    }
  }
  return finish_launch("simt_backward");
}

}  // namespace angio
