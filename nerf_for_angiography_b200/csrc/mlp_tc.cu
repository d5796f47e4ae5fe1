// mlp_tc.cu -- the fused CPPN MLP on Blackwell tensor cores (ANGIO_PREC_BF16): Fourier positional encoding fused
// into the first layer, every layer's weights resident in shared memory, activations never leave the SM.
// Replaces CPPN.forward (/root/reference/model/CPPN.py:166-222), the chunk loop of get_predictions
// (/root/reference/nerf/nerf_helpers.py:24-45), the midpoint gather (/root/reference/nerf/run_nerf_acc.py:290-292)
// and, through the fused output transforms, the alpha_fn / occ_eval_fn closures
// (/root/reference/nerf/nerf_helpers_acc.py:11-25,66-70).
//
// Kernel shape (width 128; one persistent CTA per SM, 288 threads = 9 warps):
//   warp 0      : loads the packed bf16 weight image once with 1-D bulk async copies (TMA unit), then issues every
//                 tcgen05.mma (one elected lane).
//   warps 1-4   : "group 0", warps 5-8: "group 1".  Each group owns one 128-sample tile slot: thread r <-> sample
//                 row r <-> TMEM lane r.  A group computes the encoded features of its tile, stores them as bf16
//                 into TMEM (tcgen05.st), and after each layer's MMA pulls the fp32 accumulators back
//                 (tcgen05.ld), adds the bias, applies ReLU, re-packs to bf16 and feeds them straight back into
//                 TMEM as the next layer's A operand.  The last hidden layer is reduced against w_out on CUDA cores.
//   Two slots ping-pong: while group 0 runs the epilogue of layer l, the tensor core runs layer l of group 1's tile.
//
//   D[128 x 128] (TMEM, fp32) = A[128 x K] (TMEM, bf16, K-major)  x  W_l[128 x K]^T (SMEM, bf16, SWIZZLE_128B)
//
// TMEM map (512 columns allocated): accumulators slot s at columns [128 s, 128 s + 128); A operand slot s at columns
// [256 + 64 s, 256 + 64 s + 64) (two bf16 per 32-bit column).
// First-layer K layout: [x_hi(3) x_lo(3) (sin_j, cos_j) x 3L, 0-pad] -- world coordinates reach +-173 and bf16 keeps 8
// bits, so x is fed as a hi/lo bf16 pair against duplicated weight columns.  The phase a = fl(fl(2 pi x) c) is
// computed exactly as the reference does in fp32, reduced by 2 pi with a two-constant Cody-Waite step, then
// sin/cos use the SFU (abs err ~1e-6, far below bf16 resolution).
#include "mlp_layout.cuh"
#include "tc05.cuh"

namespace {

using angio::MlpLayout;
using namespace tc05;

constexpr int kH = 128;            // hidden width handled by this kernel
constexpr int kTile = 128;         // samples per tile (UMMA M)
constexpr int kThreads = 288;      // warp 0 = control/MMA, warps 1-4 / 5-8 = tile groups
constexpr int kTmemCols = 512;
constexpr float kTwoPi = 6.2831855f;

struct TcPlan {
  int n_hidden;     // number of 128x128 layers
  int basis;        // Fourier basis L (0 = no encoding)
  int k0;           // true first-layer K: 6 + 6 L
  int k0_pad;       // rounded up to 16
  int off_const;    // byte offset of the fp32 constant block inside the packed image
  int n_const;      // floats in the constant block
  int total_bytes;  // packed image size (multiple of 16)
};

__host__ __device__ inline int w_offset(int l) { return l == 0 ? 0 : 16384 + (l - 1) * 32768; }

inline bool make_plan(const MlpLayout& L, TcPlan* p) {
  if (L.H != kH) return false;
  p->n_hidden = L.n_hidden;
  p->basis = L.basis;
  p->k0 = 6 + 6 * L.basis;
  p->k0_pad = (p->k0 + 15) / 16 * 16;
  if (p->k0_pad > 64) return false;
  p->off_const = 16384 + L.n_hidden * 32768;
  // biases (n_hidden+1) x 128 | w_out 128 | b_out (padded to 4) | coef (padded to 4)
  p->n_const = (L.n_hidden + 2) * 128 + 4 + (3 * L.basis + 3) / 4 * 4;
  p->total_bytes = (p->off_const + p->n_const * 4 + 15) / 16 * 16;
  return p->total_bytes + 2048 <= 227 * 1024;
}

// ------------------------------------------------------------------------------------------------ weight packing
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ params, MlpLayout L, TcPlan P, uint8_t* __restrict__ out) {
  const int n_w_elems = (16384 + P.n_hidden * 32768) / 2;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid < n_w_elems) {
    float v = 0.0f;
    uint32_t byte_off;
    if (tid < 8192) {  // layer 0: [128 n][64 k]
      const int n = tid / 64, k = tid % 64;
      if (k < P.k0) {
        // k: [x_hi(3) x_lo(3) | (sin_j, cos_j) pairs]; reference columns: [x(3) | sin(3L) | cos(3L)]
        const int jj = (k - 6) / 2;
        const int src = (k < 3) ? k : (k < 6 ? k - 3 : (((k - 6) & 1) ? 3 + 3 * P.basis + jj : 3 + jj));
        v = params[L.off_w[0] + (int64_t)n * L.d_in + src];
      }
      byte_off = sw128_offset(n, k);
    } else {
      const int e = tid - 8192;
      const int l = 1 + e / 16384, r = e % 16384;
      const int n = r / 128, k = r % 128;
      v = params[L.off_w[l] + (int64_t)n * kH + k];
      byte_off = w_offset(l) + (k / 64) * 16384 + sw128_offset(n, k % 64);
    }
    *reinterpret_cast<__nv_bfloat16*>(out + byte_off) = __float2bfloat16_rn(v);
  }
  if (tid < P.n_const) {
    float v = 0.0f;
    const int nb = (P.n_hidden + 1) * 128;
    if (tid < nb) v = params[L.off_b[tid / 128] + tid % 128];
    else if (tid < nb + 128) v = params[L.off_w[L.n_linear - 1] + (tid - nb)];
    else if (tid == nb + 128) v = params[L.off_b[L.n_linear - 1]];
    else if (tid >= nb + 132 && tid < nb + 132 + 3 * P.basis) v = params[L.off_coef + (tid - nb - 132)];
    reinterpret_cast<float*>(out + P.off_const)[tid] = v;
  }
}

// ------------------------------------------------------------------------------------------------ forward kernel
struct __align__(8) FwdBarriers {
  uint64_t w_ready;
  uint64_t a_ready[2];
  uint64_t acc_ready[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void sincos_reduced(float a, float& s, float& c) {
  // Cody-Waite reduction by 2*pi (two constants), then SFU sin/cos on [-pi, pi]
  const float k = rintf(a * 0.15915494309189535f);
  float r = fmaf(-k, 6.2831854820251465f, a);       // 2*pi rounded to fp32
  r = fmaf(-k, -1.7484555e-07f, r);                  // 2*pi - fl(2*pi)
  s = __sinf(r);
  c = __cosf(r);
}

template <int OUT_MODE>
__device__ __forceinline__ float out_transform(float logit, float dt) {
  if (OUT_MODE == ANGIO_OUT_LOGIT) return logit;
  const float s = 1.0f / (1.0f + __expf(-logit));
  if (OUT_MODE == ANGIO_OUT_SIGMA) return s;
  return 1.0f - __expf(-s * dt);
}

template <int OUT_MODE>
__global__ void __launch_bounds__(kThreads, 1) mlp_fwd_tc_kernel(const uint8_t* __restrict__ packed, TcPlan P, angio_samples in,
                                                                 float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ FwdBarriers bars;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t n = in.n;
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  // tiles of this CTA: blockIdx.x + j * gridDim.x, j = 0, 1, ...; slot = j & 1
  const int64_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (threadIdx.x == 0) {
    mbar_init(&bars.w_ready, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&bars.a_ready[s], 128); mbar_init(&bars.acc_ready[s], 1); }
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&bars.tmem_base, kTmemCols); tmem_relinquish(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = bars.tmem_base;
  const float* consts = reinterpret_cast<const float*>(smem + P.off_const);
  const int L1 = P.n_hidden + 1;  // number of MMA layers

  if (warp == 0) {
    // ===================== control warp: weight load, then MMA issue =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars.w_ready, (uint32_t)P.total_bytes);
      for (int off = 0; off < P.total_bytes; off += 32768) {
        const int bytes = (P.total_bytes - off < 32768) ? P.total_bytes - off : 32768;
        bulk_g2s(smem + off, packed + off, (uint32_t)bytes, &bars.w_ready);
      }
    }
    mbar_wait(&bars.w_ready, 0);
    const uint32_t idesc = make_idesc_bf16(kTile, kH, 0, 0);
    const uint32_t smem_base = smem_u32(smem);
    uint32_t phase[2] = {0, 0};
    for (int64_t j0 = 0; j0 < my_tiles; j0 += 2) {
      const int n_slots = (my_tiles - j0 >= 2) ? 2 : 1;
      for (int l = 0; l < L1; ++l) {
        for (int s = 0; s < n_slots; ++s) {
          mbar_wait(&bars.a_ready[s], phase[s]);
          phase[s] ^= 1;
          fence_after_sync();
          if (lane == 0) {
            const uint32_t d_tmem = tmem + s * 128;
            const uint32_t a_tmem = tmem + 256 + s * 64;
            const int ksteps = (l == 0) ? P.k0_pad / 16 : kH / 16;
            const uint32_t wbase = smem_base + w_offset(l);
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t db = make_smem_desc_sw128(wbase + (k / 4) * 16384 + (k % 4) * 32, 16, 1024);
              mma_ts(d_tmem, a_tmem + k * 8, db, idesc, k > 0);
            }
            mma_commit(&bars.acc_ready[s]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== tile groups: features, epilogues, output =====================
    const int g = (warp - 1) / 4;                 // slot
    const int q = warp % 4;                       // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;                // sample row inside the tile
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t acc_tmem = tmem + g * 128 + lane_off;
    const uint32_t a_tmem = tmem + 256 + g * 64 + lane_off;
    mbar_wait(&bars.w_ready, 0);                  // biases / coefficients live in the packed image
    const float* coef = consts + (P.n_hidden + 2) * 128 + 4;
    const float* w_out = consts + (P.n_hidden + 1) * 128;
    const float b_out = consts[(P.n_hidden + 2) * 128];
    const int nb = 3 * P.basis;
    uint32_t phase = 0;
    for (int64_t j = g; j < my_tiles; j += 2) {
      const int64_t tile = blockIdx.x + j * gridDim.x;
      const int64_t i = tile * kTile + row;
      const bool valid = i < n;
      // ---- sample position (reference op order) and encoded features -> TMEM A operand
      float x[3] = {0.f, 0.f, 0.f};
      float dt = 0.f;
      if (valid) {
        angio::sample_position(in, i, x);
        if (OUT_MODE == ANGIO_OUT_ALPHA) dt = in.t_ends[i] - in.t_starts[i];
      }
      {
        // K layout: [x_hi(3) x_lo(3) | (sin_j, cos_j) pairs | 0-pad]; every index below is static
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) pk[c] = 0u;
        float hi[3], lo[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          hi[c] = __bfloat162float(__float2bfloat16_rn(x[c]));
          lo[c] = x[c] - hi[c];
        }
        pk[0] = pack_bf16x2(hi[0], hi[1]);
        pk[1] = pack_bf16x2(hi[2], lo[0]);
        pk[2] = pack_bf16x2(lo[1], lo[2]);
#pragma unroll
        for (int jf = 0; jf < 29; ++jf) {      // up to 29 (sin, cos) pairs fit k0_pad <= 64
          if (jf < nb) {
            const float a = __fmul_rn(__fmul_rn(kTwoPi, x[jf % 3]), coef[jf]);
            float sn, cs;
            sincos_reduced(a, sn, cs);
            pk[3 + jf] = pack_bf16x2(sn, cs);
          }
        }
        uint32_t v16[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) v16[c] = pk[c];
        tmem_st16(a_tmem, v16);
        if (P.k0_pad > 32) {
#pragma unroll
          for (int c = 0; c < 16; ++c) v16[c] = pk[16 + c];
          tmem_st16(a_tmem + 16, v16);
        }
      }
      wait_st();
      fence_before_sync();
      mbar_arrive(&bars.a_ready[g]);
      // ---- layers
      float dot = 0.0f;
      for (int l = 0; l < L1; ++l) {
        mbar_wait(&bars.acc_ready[g], phase);
        phase ^= 1;
        fence_after_sync();
        const float* bias = consts + l * 128;
        const bool last = (l == L1 - 1);
#pragma unroll 1
        for (int c0 = 0; c0 < kH; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(acc_tmem + c0, r);
          wait_ld();
          if (!last) {
            uint32_t v16[16];
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const float2 b2 = *reinterpret_cast<const float2*>(bias + c0 + 2 * jj);
              v16[jj] = pack_bf16x2_relu(__uint_as_float(r[2 * jj]) + b2.x, __uint_as_float(r[2 * jj + 1]) + b2.y);
            }
            tmem_st16(a_tmem + c0 / 2, v16);
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              dot = fmaf(fmaxf(__uint_as_float(r[jj]) + bias[c0 + jj], 0.0f), w_out[c0 + jj], dot);
          }
        }
        if (!last) {
          wait_st();
          fence_before_sync();
          mbar_arrive(&bars.a_ready[g]);
        }
      }
      if (valid) out[i] = out_transform<OUT_MODE>(dot + b_out, dt);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace

namespace angio {

bool tc_supported(const MlpLayout& L) {
  TcPlan P;
  return make_plan(L, &P);
}
int64_t tc_packed_bytes(const MlpLayout& L) {
  TcPlan P;
  return make_plan(L, &P) ? P.total_bytes : 0;
}
int64_t simt_workspace_bytes(const MlpLayout& L, int64_t n, int training);
int64_t simt_saved_bytes(const MlpLayout& L, int64_t n);
int simt_forward(const MlpLayout& L, const float* params, const angio_samples& in, int out_mode, float* out, void* saved,
                 void* workspace, int64_t workspace_bytes, cudaStream_t st);
int simt_backward(const MlpLayout& L, const float* params, const angio_samples& in, const void* saved, const float* grad_out,
                  float* grad_params, void* workspace, int64_t workspace_bytes, cudaStream_t st);

// Training forward/backward still run the fp32 path until the tcgen05 backward lands; inference (the dominant
// cost: the no-grad visibility pass over every marched sample) runs on tensor cores.
int64_t tc_workspace_bytes(const MlpLayout& L, int64_t n, int training) { return training ? simt_workspace_bytes(L, n, 1) : 256; }
int64_t tc_saved_bytes(const MlpLayout& L, int64_t n) { return simt_saved_bytes(L, n); }

int tc_pack_weights(const MlpLayout& L, const float* params, void* packed, cudaStream_t st) {
  TcPlan P;
  if (!make_plan(L, &P)) { set_error("tc_pack_weights: unsupported shape"); return ANGIO_ERR_UNSUPPORTED; }
  const int n_w_elems = (16384 + P.n_hidden * 32768) / 2;
  const int total = n_w_elems > P.n_const ? n_w_elems : P.n_const;
  angio::note_launch(); pack_kernel<<<blocks_for(total, 256), 256, 0, st>>>(params, L, P, reinterpret_cast<uint8_t*>(packed));
  return finish_launch("tc_pack_weights");
}

template <int MODE>
static int launch_fwd(const TcPlan& P, const void* packed, const angio_samples& in, float* out, cudaStream_t st) {
  const size_t smem = (size_t)P.total_bytes + 1024;  // + slack for the 1024-byte alignment of the dynamic window
  static size_t attr_smem = 0;  // per instantiation; one device per process
  if (attr_smem < smem) {
    cudaError_t e = cudaFuncSetAttribute(mlp_fwd_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      cudaGetLastError();  // clear
      set_error("cudaFuncSetAttribute(%zu bytes): %s", smem, cudaGetErrorString(e));
      return (int)e;
    }
    attr_smem = smem;
  }
  const int64_t n_tiles = (in.n + kTile - 1) / kTile;
  int grid = sm_count();
  if (n_tiles < grid) grid = (int)n_tiles;
  angio::note_launch(); mlp_fwd_tc_kernel<MODE><<<grid, kThreads, smem, st>>>(reinterpret_cast<const uint8_t*>(packed), P, in, out);
  return finish_launch("mlp_fwd_tc_kernel");
}

int tc_forward(const MlpLayout& L, const float* params, const void* packed, const angio_samples& in, int out_mode, float* out,
               void* saved, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  if (saved) return simt_forward(L, params, in, out_mode, out, saved, workspace, workspace_bytes, st);
  TcPlan P;
  if (!make_plan(L, &P)) { set_error("tc_forward: unsupported shape"); return ANGIO_ERR_UNSUPPORTED; }
  if (in.n == 0) return 0;
  if ((reinterpret_cast<uintptr_t>(packed) & 15) != 0) { set_error("tc_forward: packed image must be 16-byte aligned"); return ANGIO_ERR_INVALID_ARG; }
  switch (out_mode) {
    case ANGIO_OUT_LOGIT: return launch_fwd<ANGIO_OUT_LOGIT>(P, packed, in, out, st);
    case ANGIO_OUT_SIGMA: return launch_fwd<ANGIO_OUT_SIGMA>(P, packed, in, out, st);
    default: return launch_fwd<ANGIO_OUT_ALPHA>(P, packed, in, out, st);
  }
}

int tc_backward(const MlpLayout& L, const float* params, const void* packed, const angio_samples& in, const void* saved,
                const float* grad_out, float* grad_params, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  (void)packed;
  return simt_backward(L, params, in, saved, grad_out, grad_params, workspace, workspace_bytes, st);
}

}  // namespace angio
