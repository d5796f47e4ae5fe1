// mlp_tc256.cu -- the fused CPPN MLP at hidden width 256 (BASELINE config 4: 8 x 256) on Blackwell tensor cores: forward,
// data-gradient chain and weight gradients as hand-written tcgen05 kernels with the layer weights STREAMED through shared memory.
// Replaces the same reference code as mlp_tc.cu (CPPN.forward /root/reference/model/CPPN.py:166-222 with the width / depth the
// constructor allows, :96-131; get_predictions nerf/nerf_helpers.py:24-45; the midpoint gather run_nerf_acc.py:290-292; the
// alpha_fn / occ_eval_fn closures nerf_helpers_acc.py:11-25,66-70; autograd through the MLP run_nerf_acc.py:306).
//
// Why a second kernel family.  At width 128 all weights (151 KB of bf16) stay resident in shared memory and two or three sample
// tiles share the tensor core.  At width 256 one layer alone is 128 KB (8 layers: 1 MB) and one 128-sample tile needs 256
// accumulator + 128 operand columns of TMEM, so: ONE tile slot per CTA, N = 256 instructions (the wide instruction streams at
// 87 % of the tensor peak vs 64 % for N = 128, tools/tc_probe.cu), and every layer's weights arrive as four 32 KB K-chunks
// ([256 n][64 k], SWIZZLE_128B, ready-to-MMA bytes in HBM / L2) through a 5-stage cp.async.bulk ring that a producer warp keeps
// full independently of the epilogue.
//
//   D[128 x 256] (TMEM fp32) = A[128 x K] (TMEM bf16: the activations never leave the SM) x W[256 x K]^T (SMEM ring)
//
// Warp roles (320 threads, 1 CTA / SM, persistent over tiles): warp 0 = producer (bulk copies of weight chunks + the fp32
// constant block), warp 1 = MMA issuer, warps 2..9 = epilogue: 4 TMEM lane quadrants x 2 column halves of 128 columns, four
// passes of 32 columns each with the next pass's tcgen05.ld in flight under the current pass's math.
// The data-gradient chain is the same pipeline over a TRANSPOSED weight image (K = out features), so it streams with N = 256
// as well; the weight-gradient kernel gives each CTA (layer, half of the out features, slice of the tiles) 256 + 16 accumulator
// columns that persist over its tiles, both operands MN-major from the saved tile images.
#include <stdlib.h>

#include "mlp_layout.cuh"
#include "tc05.cuh"

namespace {

using angio::MlpLayout;
using namespace tc05;

constexpr int kW = 256;             // hidden width handled by these kernels
constexpr int kTile = 128;          // samples per tile (UMMA M)
constexpr int kChunk = 32768;       // one weight K-chunk: [256 n][64 k] bf16
constexpr int kRing = 5;            // chunks in flight
constexpr int kThreads = 320;       // producer warp, MMA warp, 8 epilogue warps
constexpr int kA0Bytes = 16384;     // a_0 tile image [128 x 64]
constexpr int kActBytes = 65536;    // a_d / delta_d tile image: four [128 x 64] blocks
constexpr int kMaskBytes = 4096;    // ReLU bit mask of one a_d tile: [128 rows][8 x 32 bits]
constexpr int kStageBytes = 8 * 4096;
constexpr float kTwoPi = 6.2831855f;
constexpr uint32_t kColD = 0, kColA = 256;   // TMEM columns of the dgrad pipeline (the forward uses two 256-column regions)

struct Plan256 {
  int n_hidden, basis, k0, k0_pad;
  int n_chunks_fwd;        // 1 + 4 L
  int64_t off_bwd;         // transposed image: layers L..1 (4 x 32 KB each), then layer 0 (4 x 8 KB)
  int64_t off_const;       // fp32 constants: biases (L+1) x 256 | w_out 256 | b_out (4) | coef (padded to 4)
  int n_const;
  int64_t total_bytes;
};

inline bool make_plan(const MlpLayout& L, Plan256* p) {
  if (L.H != kW || L.n_hidden < 1 || L.n_hidden > 16) return false;
  p->n_hidden = L.n_hidden;
  p->basis = L.basis;
  p->k0 = 6 + 6 * L.basis;
  p->k0_pad = (p->k0 + 15) / 16 * 16;
  if (p->k0_pad > 64) return false;
  p->n_chunks_fwd = 1 + 4 * L.n_hidden;
  p->off_bwd = (int64_t)p->n_chunks_fwd * kChunk;
  p->off_const = p->off_bwd + (int64_t)4 * L.n_hidden * kChunk + kChunk;
  p->n_const = (L.n_hidden + 2) * kW + 4 + (3 * L.basis + 3) / 4 * 4;
  p->total_bytes = p->off_const + ((int64_t)p->n_const * 4 + 15) / 16 * 16;
  return true;
}

// ------------------------------------------------------------------------------------------------ weight packing
// first-layer K layout (as in mlp_tc.cu): [x_hi(3) x_lo(3) (sin_j, cos_j) x 3 basis, 0-pad]; reference columns [x | sin | cos]
__device__ __forceinline__ int ref_col0(int k, int basis) {
  const int jj = (k - 6) / 2;
  return (k < 3) ? k : (k < 6 ? k - 3 : (((k - 6) & 1) ? 3 + 3 * basis + jj : 3 + jj));
}

__global__ void __launch_bounds__(256) pack256_kernel(const float* __restrict__ params, MlpLayout L, Plan256 P, uint8_t* __restrict__ out) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t n_l0 = kW * 64, n_h = (int64_t)kW * kW;
  const int64_t n_fwd = n_l0 + P.n_hidden * n_h;
  const int64_t n_all = 2 * n_fwd;
  if (tid < n_all) {
    float v = 0.0f;
    int64_t byte_off;
    if (tid < n_l0) {                                   // forward layer 0: [256 n][64 k]
      const int n = (int)(tid / 64), k = (int)(tid % 64);
      if (k < P.k0) v = params[L.off_w[0] + (int64_t)n * L.d_in + ref_col0(k, P.basis)];
      byte_off = sw128_offset(n, k);
    } else if (tid < n_fwd) {                           // forward layer w: four K-chunks [256 n][64 k]
      const int64_t e = tid - n_l0;
      const int w = 1 + (int)(e / n_h);
      const int r = (int)(e % n_h), n = r / kW, k = r % kW;
      v = params[L.off_w[w] + (int64_t)n * kW + k];
      byte_off = (int64_t)kChunk * (1 + (w - 1) * 4 + k / 64) + sw128_offset(n, k % 64);
    } else if (tid < n_fwd + P.n_hidden * n_h) {        // transposed layer w (order L, L-1, .., 1): rows = in features, K = out features
      const int64_t e = tid - n_fwd;
      const int j = (int)(e / n_h), w = P.n_hidden - j;
      const int r = (int)(e % n_h), i = r / kW, o = r % kW;
      v = params[L.off_w[w] + (int64_t)o * kW + i];
      byte_off = P.off_bwd + (int64_t)kChunk * (j * 4 + o / 64) + sw128_offset(i, o % 64);
    } else {                                            // transposed layer 0: rows = feature columns (64), K = out features: four 8 KB chunks
      const int64_t e = tid - n_fwd - P.n_hidden * n_h;
      const int f = (int)(e / kW), o = (int)(e % kW);
      if (f < P.k0) v = params[L.off_w[0] + (int64_t)o * L.d_in + ref_col0(f, P.basis)];
      byte_off = P.off_bwd + (int64_t)kChunk * 4 * P.n_hidden + 8192 * (o / 64) + sw128_offset(f, o % 64);
    }
    *reinterpret_cast<__nv_bfloat16*>(out + byte_off) = __float2bfloat16_rn(v);
  }
  if (tid < P.n_const) {
    float v = 0.0f;
    const int t = (int)tid, nb = (P.n_hidden + 1) * kW;
    if (t < nb) v = params[L.off_b[t / kW] + t % kW];
    else if (t < nb + kW) v = params[L.off_w[L.n_linear - 1] + (t - nb)];
    else if (t == nb + kW) v = params[L.off_b[L.n_linear - 1]];
    else if (t >= nb + kW + 4 && t < nb + kW + 4 + 3 * P.basis) v = params[L.off_coef + (t - nb - kW - 4)];
    reinterpret_cast<float*>(out + P.off_const)[t] = v;
  }
}

// ------------------------------------------------------------------------------------------------ shared device pieces
struct __align__(8) Bars256 {
  uint64_t full[kRing];
  uint64_t empty[kRing];
  uint64_t c_ready;      // constant block landed
  uint64_t a_ready;      // (dgrad) A operand (or features) of the next stage written by all 8 epilogue warps
  uint64_t a_pass[4];    // (forward) pass p of the epilogue done: 32 + 32 more features of the next A operand are in place
  uint64_t acc_ready;    // accumulator of the current stage complete
  uint32_t tmem_base;
};

__device__ __forceinline__ void sincos_reduced(float a, float& s, float& c) {
  const float k = rintf(a * 0.15915494309189535f);     // Cody-Waite reduction by 2*pi (two constants), then SFU sin/cos
  float r = fmaf(-k, 6.2831854820251465f, a);
  r = fmaf(-k, -1.7484555e-07f, r);
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

// 8 consecutive bf16x2 words (16 K columns) of the encoded features, chunk c8 = words [8 c8, 8 c8 + 8): word 0..2 = x hi/lo,
// word 3 + j = (sin_j, cos_j) -- identical to mlp_tc.cu
__device__ __forceinline__ void encode_feature_chunk(const float x[3], const float* __restrict__ coef, int nb, int c8, uint32_t (&v)[8]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int wd = c8 * 8 + e;
    uint32_t val = 0u;
    if (wd < 3) {
      float hi[3], lo[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) { hi[c] = __bfloat162float(__float2bfloat16_rn(x[c])); lo[c] = x[c] - hi[c]; }
      val = (wd == 0) ? pack_bf16x2(hi[0], hi[1]) : (wd == 1) ? pack_bf16x2(hi[2], lo[0]) : pack_bf16x2(lo[1], lo[2]);
    } else if (wd - 3 < nb) {
      const int jf = wd - 3;
      const float a = __fmul_rn(__fmul_rn(kTwoPi, x[jf % 3]), coef[jf]);
      float sn, cs;
      sincos_reduced(a, sn, cs);
      val = pack_bf16x2(sn, cs);
    }
    v[e] = val;
  }
}

// 16 bf16x2 words -> 32 ReLU bits (bit k = low element of word k, bit 16 + k = high element); values are post-ReLU (>= +0)
__device__ __forceinline__ uint32_t relu_bits16(const uint32_t* w) {
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) m |= ((w[k] + 0x7FFF7FFFu) & 0x80008000u) >> (15 - k);
  return m;
}
__device__ __forceinline__ uint32_t relu_word_mask(uint32_t m, int k) {
  const uint32_t b = (m >> k) & 0x00010001u;
  return (b << 16) - b;
}

__device__ __forceinline__ void signal_ready(uint64_t* bar, int lane) {
  wait_st();
  fence_before_sync();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// One pass (32 columns = 64 bytes = half a swizzled block row) of this lane's row into the warp's 4 KB staging buffer; after the
// second pass of a block the warp sends the 32 rows x 128 B with one bulk copy.
__device__ __forceinline__ void stage_half_row(uint8_t* __restrict__ stage, int lane, int half, const uint32_t (&pk)[16]) {
  uint8_t* base = stage + lane * 128;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(base + (((half * 4 + c) ^ (lane & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
}
__device__ __forceinline__ void flush_stage(uint8_t* __restrict__ stage, uint8_t* __restrict__ gdst, int lane) {
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    bulk_s2g(gdst, stage, 4096);
    bulk_commit();
  }
}
__device__ __forceinline__ void stage_acquire(int lane) {
  if (lane == 0) bulk_wait_read<0>();           // the previous bulk copy out of this buffer has been read
  __syncwarp();
}

__device__ __forceinline__ void pipe_setup(Bars256& bars, int warp) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < kRing; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); }
    mbar_init(&bars.c_ready, 1);
    mbar_init(&bars.a_ready, 8);
    for (int p = 0; p < 4; ++p) mbar_init(&bars.a_pass[p], 8);
    mbar_init(&bars.acc_ready, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&bars.tmem_base, 512); tmem_relinquish(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (bars.tmem_base != 0) __trap();              // one CTA per SM owning all 512 columns: column numbers are compile-time constants
}

// 32 accumulator columns + bias -> 16 packed bf16x2 words (ReLU'd), or (LAST) one sequential fp32 dot-product chain with w_out
template <bool LAST>
__device__ __forceinline__ float bias_relu_32(const uint32_t (&r)[32], const float* __restrict__ bias, const float* __restrict__ w_out,
                                              uint32_t (&pk)[16], float dot) {
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const float4 b0 = *reinterpret_cast<const float4*>(bias + 4 * jj);
    const float2 u0 = __fadd2_rn(make_float2(__uint_as_float(r[4 * jj]), __uint_as_float(r[4 * jj + 1])), make_float2(b0.x, b0.y));
    const float2 u1 = __fadd2_rn(make_float2(__uint_as_float(r[4 * jj + 2]), __uint_as_float(r[4 * jj + 3])), make_float2(b0.z, b0.w));
    if (LAST) {
      const float4 w0 = *reinterpret_cast<const float4*>(w_out + 4 * jj);
      dot = fmaf(fmaxf(u0.x, 0.f), w0.x, dot); dot = fmaf(fmaxf(u0.y, 0.f), w0.y, dot);
      dot = fmaf(fmaxf(u1.x, 0.f), w0.z, dot); dot = fmaf(fmaxf(u1.y, 0.f), w0.w, dot);
    }
    pk[2 * jj] = pack_bf16x2_relu(u0.x, u0.y);
    pk[2 * jj + 1] = pack_bf16x2_relu(u1.x, u1.y);
  }
  return dot;
}

// ------------------------------------------------------------------------------------------------ (1) forward
// saved (TRAIN): [a_0: T x 16 KB][a_1 .. a_{L+1}: (L+1) x T x 64 KB][ReLU bit masks of a_1 .. a_{L+1}: (L+1) x T x 4 KB]
template <int OUT_MODE, bool TRAIN>
__global__ void __launch_bounds__(kThreads, 1) mlp256_fwd_kernel(const uint8_t* __restrict__ packed, Plan256 P, angio_samples in,
                                                                 float* __restrict__ out, uint8_t* __restrict__ saved) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ Bars256 bars;
  __shared__ float s_dot[kTile];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x / 32), 0);
  const int lane = threadIdx.x % 32;
  int64_t n = in.n;
  const int64_t lay_tiles = (n + kTile - 1) / kTile;                     // tile-image layout stride: capacity of the arrays
  if (in.n_dev) { const int64_t nd = *in.n_dev; n = nd < n ? nd : n; }
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  const int64_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  pipe_setup(bars, warp);
  uint8_t* ring = smem;
  const float* consts = reinterpret_cast<const float*>(smem + kRing * kChunk);
  const int L = P.n_hidden;
  const int n_stages = L + 1;

  if (warp == 0) {
    // ===================== producer: constants once, then the weight chunks of every tile, in MMA order =====================
    if (lane == 0) {
      const uint32_t cbytes = (uint32_t)((P.n_const * 4 + 15) / 16 * 16);
      mbar_arrive_expect_tx(&bars.c_ready, cbytes);
      bulk_g2s(smem + kRing * kChunk, packed + P.off_const, cbytes, &bars.c_ready);
      uint32_t it = 0;
      for (int64_t j = 0; j < my_tiles; ++j)
        for (int c = 0; c < P.n_chunks_fwd; ++c, ++it) {
          const int slot = it % kRing;
          // MMA order of a hidden layer's K-chunks: 0, 2, 1, 3 (the two column halves of the epilogue finish chunks 0 and 2 first)
          const int cc = c == 0 ? 0 : 1 + ((c - 1) & ~3) + (((c - 1) & 1) << 1) + (((c - 1) & 2) >> 1);
          mbar_wait(&bars.empty[slot], ((it / kRing) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars.full[slot], kChunk);
          bulk_g2s(ring + slot * kChunk, packed + (int64_t)cc * kChunk, kChunk, &bars.full[slot]);
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // TMEM = two 256-column regions.  Stage g (global count over tiles and layers) accumulates into region g & 1 and reads its A
    // operand from the other one, where the epilogue of stage g - 1 wrote it IN PLACE of the accumulator columns it had drained:
    // hidden units [128 h + 32 p, +32) of pass p land in columns 128 h + 16 p .. +16 of that region.  The issuer follows the epilogue
    // pass by pass (four k-steps = 2 x 32 features per pass), so a layer's MMAs overlap the previous layer's epilogue.
    const uint32_t idesc = make_idesc_bf16(kTile, kW, 0, 0);
    const uint32_t ring_base = smem_u32(ring);
    uint32_t it = 0, g = 0;
    uint32_t phm = 0;                                          // bit p: parity the issuer waits for on a_pass[p]
    for (int64_t j = 0; j < my_tiles; ++j) {
      for (int st = 0; st < n_stages; ++st, ++g) {
        const uint32_t d_col = (g & 1) * 256, a_reg = ((g & 1) ^ 1) * 256;
        if (st == 0) {
          mbar_wait(&bars.a_pass[0], phm & 1u); phm ^= 1u;
          const int slot = it % kRing;
          mbar_wait(&bars.full[slot], (it / kRing) & 1);
          fence_after_sync();
          if (lane == 0) {
            const uint32_t wbase = ring_base + slot * kChunk;
            for (int k = 0; k < P.k0_pad / 16; ++k)               // feature chunk k: columns 128 (k & 1) + 8 (k / 2)
              mma_ts(d_col, a_reg + (k & 1) * 128 + (k >> 1) * 8, make_smem_desc_sw128(wbase + k * 32, 16, 1024), idesc, k != 0);
            mma_commit(&bars.empty[slot]);
            mma_commit(&bars.acc_ready);
          }
          __syncwarp();
          ++it;
          continue;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          mbar_wait(&bars.a_pass[p], (phm >> p) & 1u); phm ^= 1u << p;
          const uint32_t i0 = it + (p >> 1) * 2;                   // ring entries of K-chunks (p / 2) and 2 + (p / 2)
          const int slot0 = i0 % kRing, slot1 = (i0 + 1) % kRing;
          if ((p & 1) == 0) {
            mbar_wait(&bars.full[slot0], (i0 / kRing) & 1);
            mbar_wait(&bars.full[slot1], ((i0 + 1) / kRing) & 1);
          }
          fence_after_sync();
          if (lane == 0) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const uint32_t wbase = ring_base + (hh ? slot1 : slot0) * kChunk;
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const int k = 2 * (p & 1) + kk;                     // k-step inside the 64-feature chunk
                mma_ts(d_col, a_reg + hh * 128 + p * 16 + kk * 8, make_smem_desc_sw128(wbase + k * 32, 16, 1024), idesc,
                       (p | hh | kk) != 0);
              }
            }
            if (p & 1) { mma_commit(&bars.empty[slot0]); mma_commit(&bars.empty[slot1]); }
            if (p == 3) mma_commit(&bars.acc_ready);
          }
          __syncwarp();
        }
        it += 4;
      }
    }
  } else {
    // ===================== epilogue warps: features, bias + ReLU + pack, output =====================
    const int q = warp % 4;                       // TMEM lane quadrant (hardware rule)
    const int h = (warp - 2) / 4;                 // column half: hidden units [128 h, 128 h + 128)
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t half_col = lane_off + h * 128;   // this warp's 128 columns inside either region: accumulator in, next A operand out
    mbar_wait(&bars.c_ready, 0);
    const float* coef = consts + (L + 2) * kW + 4;
    const float b_out = consts[(L + 2) * kW];
    const float* w_out = consts + (L + 1) * kW + h * 128;
    const int pair_bar = 1 + q;                   // named barrier shared by the two column-half warps of this row quadrant
    uint8_t* stage = TRAIN ? smem + kRing * kChunk + 20480 + (warp - 2) * 4096 : nullptr;
    uint8_t* act_base = TRAIN ? saved + lay_tiles * kA0Bytes : nullptr;
    uint8_t* mask_base = TRAIN ? act_base + (int64_t)(L + 1) * lay_tiles * kActBytes : nullptr;
    const int nb = 3 * P.basis;
    uint32_t phase = 0;
    auto fetch = [&](int64_t jt, float (&xx)[3], float& dtt, bool& vv, int64_t& idx) {
      int64_t ii = (blockIdx.x + jt * gridDim.x) * kTile + row;
      vv = (jt < my_tiles) && (ii < n);
      xx[0] = xx[1] = xx[2] = 0.f;
      dtt = 0.f;
      if (vv) {
        if (in.sample_idx) ii = in.sample_idx[ii];
        angio::sample_position(in, ii, xx);
        if (OUT_MODE == ANGIO_OUT_ALPHA) dtt = in.t_ends[ii] - in.t_starts[ii];
      }
      idx = ii;
    };
    // Features of a tile: the two column halves split the 16-column K chunks of a_0 (chunk c8 -> half c8 & 1); this warp's chunks
    // c8 = h and h + 2 are kept in registers and stored to columns [0, 16) of its half of a region once those are free.
    uint32_t feat[2][8];
    auto encode = [&](int64_t jt, const float (&xx)[3]) {
      const int64_t tl = blockIdx.x + jt * gridDim.x;
      uint8_t* a0_row = TRAIN ? saved + tl * kA0Bytes + row * 128 : nullptr;
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int c8 = 2 * c2 + h;
#pragma unroll
        for (int e = 0; e < 8; ++e) feat[c2][e] = 0u;
        if (c8 * 16 < P.k0_pad) encode_feature_chunk(xx, coef, nb, c8, feat[c2]);
        if (TRAIN) {
          *reinterpret_cast<uint4*>(a0_row + (((2 * c8) ^ (row & 7)) << 4)) = make_uint4(feat[c2][0], feat[c2][1], feat[c2][2], feat[c2][3]);
          *reinterpret_cast<uint4*>(a0_row + (((2 * c8 + 1) ^ (row & 7)) << 4)) = make_uint4(feat[c2][4], feat[c2][5], feat[c2][6], feat[c2][7]);
        }
      }
    };
    auto store_features = [&](uint32_t region_col) {
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2)
        if ((2 * c2 + h) * 16 < P.k0_pad) tmem_st8(region_col + half_col + c2 * 8, feat[c2]);
    };
    float xn[3], dtn;
    bool vn;
    int64_t in_;
    fetch(0, xn, dtn, vn, in_);
    if (my_tiles > 0) { encode(0, xn); store_features(256); signal_ready(&bars.a_pass[0], lane); }   // stage 0 accumulates into region 0
    uint32_t g = 0;
    for (int64_t j = 0; j < my_tiles; ++j) {
      const int64_t tile = blockIdx.x + j * gridDim.x;
      const int64_t i = in_;
      const bool valid = vn;
      const float dt = dtn;
      fetch(j + 1, xn, dtn, vn, in_);             // in flight during this tile's layers
      for (int l = 0; l <= L; ++l, ++g) {
        const bool last = l == L;
        const bool more = j + 1 < my_tiles;
        if (last && more) encode(j + 1, xn);      // sin / cos of the next tile while the last layer's MMAs run
        mbar_wait(&bars.acc_ready, phase);
        phase ^= 1;
        fence_after_sync();
        const float* bias = consts + l * kW + h * 128;
        const uint32_t reg_col = (g & 1) * 256 + half_col;      // accumulator of this stage = A operand region of the next
        uint32_t ra[32], rb[32], pk[16];
        uint32_t mbits[4];
        float dot = 0.0f;
        tmem_ld32(reg_col, ra);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          wait_ld();
          if (p < 3) { if (p & 1) tmem_ld32(reg_col + (p + 1) * 32, ra); else tmem_ld32(reg_col + (p + 1) * 32, rb); }   // next pass in flight
          if (last) dot = (p & 1) ? bias_relu_32<true>(rb, bias + p * 32, w_out + p * 32, pk, dot) : bias_relu_32<true>(ra, bias + p * 32, w_out + p * 32, pk, dot);
          else if (p & 1) bias_relu_32<false>(rb, bias + p * 32, nullptr, pk, 0.f);
          else bias_relu_32<false>(ra, bias + p * 32, nullptr, pk, 0.f);
          // in place: the 16 packed columns of pass p replace accumulator columns [16 p, 16 p + 16) of this warp's half, drained in
          // passes <= p; after the last layer the next tile's features take columns [0, 16) instead
          if (!last) { tmem_st16(reg_col + p * 16, pk); signal_ready(&bars.a_pass[p], lane); }
          else if (p == 0 && more) { store_features((g & 1) * 256); signal_ready(&bars.a_pass[0], lane); }
          if (TRAIN) {
            if ((p & 1) == 0) stage_acquire(lane);
            stage_half_row(stage, lane, p & 1, pk);
            mbits[p] = relu_bits16(pk);
            if (p & 1) flush_stage(stage, act_base + ((int64_t)l * lay_tiles + tile) * kActBytes + (2 * h + (p >> 1)) * 16384 + q * 4096, lane);
          }
        }
        if (TRAIN)
          *reinterpret_cast<uint4*>(mask_base + (((int64_t)l * lay_tiles + tile) * kTile + row) * 32 + h * 16) = make_uint4(mbits[0], mbits[1], mbits[2], mbits[3]);
        if (last) {
          // the h = 1 warp hands its half of the dot product to the h = 0 warp of the same row quadrant
          if (h == 1) {
            s_dot[row] = dot;
            asm volatile("bar.arrive %0, 64;" ::"r"(pair_bar) : "memory");
          } else {
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            if (valid) out[i] = out_transform<OUT_MODE>(dot + s_dot[row] + b_out, dt);
          }
        }
      }
    }
    if (TRAIN && lane == 0) bulk_wait<0>();       // our bulk stores are complete before the CTA gives up its shared memory
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(0, 512);
}

// ------------------------------------------------------------------------------------------------ (2) data-gradient chain
// delta images out: [delta_1: T x 64 KB] ... [delta_{L+1}]; coef_partials: [gridDim.x][32] floats
__global__ void __launch_bounds__(kThreads, 1) mlp256_dgrad_kernel(const uint8_t* __restrict__ packed, Plan256 P, angio_samples in,
                                                                   const uint8_t* __restrict__ saved, const float* __restrict__ grad_out,
                                                                   uint8_t* __restrict__ delta, float* __restrict__ coef_partials) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ Bars256 bars;
  __shared__ float s_coef[8][16];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x / 32), 0);
  const int lane = threadIdx.x % 32;
  int64_t n = in.n;
  const int64_t lay_tiles = (n + kTile - 1) / kTile;
  if (in.n_dev) { const int64_t nd = *in.n_dev; n = nd < n ? nd : n; }
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  const int64_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  pipe_setup(bars, warp);
  uint8_t* ring = smem;
  const float* consts = reinterpret_cast<const float*>(smem + kRing * kChunk);
  const int L = P.n_hidden;
  const bool enc = P.basis > 0;
  const int n_stages = L + (enc ? 1 : 0);      // stage s < L multiplies delta_{L+1-s} by W_{L-s}; stage L (enc) by W_0 (feature gradient)
  const int nb = 3 * P.basis;
  float dc[16];
#pragma unroll
  for (int jf = 0; jf < 16; ++jf) dc[jf] = 0.0f;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t cbytes = (uint32_t)((P.n_const * 4 + 15) / 16 * 16);
      mbar_arrive_expect_tx(&bars.c_ready, cbytes);
      bulk_g2s(smem + kRing * kChunk, packed + P.off_const, cbytes, &bars.c_ready);
      uint32_t it = 0;
      for (int64_t j = 0; j < my_tiles; ++j)
        for (int st = 0; st < n_stages; ++st)
          for (int c = 0; c < 4; ++c, ++it) {
            const int slot = it % kRing;
            const uint32_t bytes = st < L ? kChunk : 8192;
            const int64_t src = P.off_bwd + (st < L ? (int64_t)(st * 4 + c) * kChunk : (int64_t)4 * L * kChunk + c * 8192);
            mbar_wait(&bars.empty[slot], ((it / kRing) & 1) ^ 1);
            mbar_arrive_expect_tx(&bars.full[slot], bytes);
            bulk_g2s(ring + slot * kChunk, packed + src, bytes, &bars.full[slot]);
          }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(kTile, kW, 0, 0);
    const uint32_t idesc0 = make_idesc_bf16(kTile, 64, 0, 0);
    const uint32_t ring_base = smem_u32(ring);
    uint32_t it = 0, a_phase = 0;
    for (int64_t j = 0; j < my_tiles; ++j) {
      for (int st = 0; st < n_stages; ++st) {
        mbar_wait(&bars.a_ready, a_phase);
        a_phase ^= 1;
        for (int c = 0; c < 4; ++c, ++it) {
          const int slot = it % kRing;
          mbar_wait(&bars.full[slot], (it / kRing) & 1);
          fence_after_sync();
          if (lane == 0) {
            const uint32_t wbase = ring_base + slot * kChunk;
            for (int k = 0; k < 4; ++k)
              mma_ts(kColD, kColA + (c * 4 + k) * 8, make_smem_desc_sw128(wbase + k * 32, 16, 1024), st < L ? idesc : idesc0, (c | k) != 0);
            mma_commit(&bars.empty[slot]);
            if (c == 3) mma_commit(&bars.acc_ready);
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int q = warp % 4;
    const int h = (warp - 2) / 4;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t acc_col = kColD + lane_off + h * 128;
    const uint32_t a_col = kColA + lane_off + h * 64;
    mbar_wait(&bars.c_ready, 0);
    const float* coef = consts + (L + 2) * kW + 4;
    const float* w_out = consts + (L + 1) * kW + h * 128;
    uint8_t* stage = smem + kRing * kChunk + 20480 + (warp - 2) * 4096;
    // ReLU bits of a_d (d >= 1), row r, column half h: mask_base + (((d-1) * lay_tiles + tile) * 128 + r) * 32 + 16 h
    const uint8_t* mask_base = saved + lay_tiles * kA0Bytes + (int64_t)(L + 1) * lay_tiles * kActBytes + row * 32 + h * 16;
    uint32_t phase = 0;
    for (int64_t j = 0; j < my_tiles; ++j) {
      const int64_t tile = blockIdx.x + j * gridDim.x;
      const int64_t i = tile * kTile + row;
      const bool valid = i < n;
      const float gr = valid ? grad_out[i] : 0.0f;
      // ---- delta_{L+1} = g * w_out * relu'(a_{L+1})
      uint4 mk = __ldg(reinterpret_cast<const uint4*>(mask_base + ((int64_t)L * lay_tiles + tile) * (kTile * 32)));
      {
        const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          uint32_t pk[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float2 w2 = *reinterpret_cast<const float2*>(w_out + p * 32 + 2 * c);
            pk[c] = pack_bf16x2(gr * w2.x, gr * w2.y) & relu_word_mask(mw[p], c);
          }
          tmem_st16(a_col + p * 16, pk);
          if ((p & 1) == 0) stage_acquire(lane);
          stage_half_row(stage, lane, p & 1, pk);
          if (p & 1) flush_stage(stage, delta + ((int64_t)L * lay_tiles + tile) * kActBytes + (2 * h + (p >> 1)) * 16384 + q * 4096, lane);
        }
        if (n_stages > 0) signal_ready(&bars.a_ready, lane);
      }
      // ---- hidden chain: delta_{d-1} = (delta_d W_{d-1}) * relu'(a_{d-1}),  d = L+1 .. 2
      for (int st = 0; st < L; ++st) {
        const int d = L + 1 - st;
        mk = __ldg(reinterpret_cast<const uint4*>(mask_base + ((int64_t)(d - 2) * lay_tiles + tile) * (kTile * 32)));   // in flight during the MMA
        mbar_wait(&bars.acc_ready, phase);
        phase ^= 1;
        fence_after_sync();
        const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
        uint32_t ra[32], rb[32];
        tmem_ld32(acc_col, ra);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          wait_ld();
          if (p < 3) { if (p & 1) tmem_ld32(acc_col + (p + 1) * 32, ra); else tmem_ld32(acc_col + (p + 1) * 32, rb); }
          uint32_t pk[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float v0 = __uint_as_float((p & 1) ? rb[2 * c] : ra[2 * c]), v1 = __uint_as_float((p & 1) ? rb[2 * c + 1] : ra[2 * c + 1]);
            pk[c] = pack_bf16x2(v0, v1) & relu_word_mask(mw[p], c);
          }
          tmem_st16(a_col + p * 16, pk);
          if ((p & 1) == 0) stage_acquire(lane);
          stage_half_row(stage, lane, p & 1, pk);
          if (p & 1) flush_stage(stage, delta + ((int64_t)(d - 2) * lay_tiles + tile) * kActBytes + (2 * h + (p >> 1)) * 16384 + q * 4096, lane);
        }
        if (st + 1 < n_stages) signal_ready(&bars.a_ready, lane);
      }
      // ---- feature gradient (delta_1 W_0) -> Fourier-coefficient gradient; half h reads feature columns [32h, 32h+32)
      if (enc) {
        mbar_wait(&bars.acc_ready, phase);
        phase ^= 1;
        fence_after_sync();
        uint32_t r[32];
        tmem_ld32(kColD + lane_off + h * 32, r);
        wait_ld();
        if (valid) {
          float x[3];
          angio::sample_position(in, i, x);
          // feature columns 6+2j (sin) and 7+2j (cos); d sin/d coef = cos * 2 pi x, d cos/d coef = -sin * 2 pi x
          if (h == 0) {
#pragma unroll
            for (int jj = 0; jj < 13; ++jj) {
              if (jj < nb) {
                const float tp = __fmul_rn(kTwoPi, x[jj % 3]);
                float sn, cs;
                sincos_reduced(__fmul_rn(tp, coef[jj]), sn, cs);
                dc[jj] = fmaf(__uint_as_float(r[6 + 2 * jj]) * cs - __uint_as_float(r[7 + 2 * jj]) * sn, tp, dc[jj]);
              }
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const int jf = 13 + jj;
              if (jf < nb) {
                const float tp = __fmul_rn(kTwoPi, x[jf % 3]);
                float sn, cs;
                sincos_reduced(__fmul_rn(tp, coef[jf]), sn, cs);
                dc[jj] = fmaf(__uint_as_float(r[2 * jj]) * cs - __uint_as_float(r[2 * jj + 1]) * sn, tp, dc[jj]);
              }
            }
          }
        }
      }
    }
    if (lane == 0) bulk_wait<0>();
    if (enc) {
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        float v = dc[jj];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (lane == 0) s_coef[warp - 2][jj] = v;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tmem_dealloc(0, 512);
    if (enc && lane < nb) {
      const int hh = lane < 13 ? 0 : 1, jj = lane - 13 * hh;
      float v = 0.0f;
      for (int wv = 0; wv < 8; ++wv)
        if (wv / 4 == hh) v += s_coef[wv][jj];
      coef_partials[blockIdx.x * 32 + lane] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ (3) weight gradients
// grid = (L+1) * 2 * G CTAs; CTA ((d-1) * 2 + mh, j) accumulates rows [128 mh, 128 mh + 128) of dW_{d-1} = delta_d^T a_{d-1} and of
// db_{d-1} over tiles j, j+G, ...; partials: [gridDim.x][128 * 256 + 128] floats.  The two CTAs of a layer slice read the same
// a_{d-1} tile images; they run side by side, so the second read is served by L2.
constexpr int kWgRing = 2;
constexpr int kWgStage = 98304;              // 32 KB delta half image + 64 KB activation image
constexpr int kWgThreads = 128;
constexpr int kWgPartial = 128 * 256 + 128;

struct __align__(8) WgBars {
  uint64_t full[kWgRing];
  uint64_t empty[kWgRing];
  uint64_t done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kWgThreads, 1) mlp256_wgrad_kernel(const uint8_t* __restrict__ saved, const uint8_t* __restrict__ delta,
                                                                     int64_t lay_tiles, const int32_t* __restrict__ n_dev, int L, int G,
                                                                     float* __restrict__ partials) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ WgBars bars;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x / 32), 0);
  const int lane = threadIdx.x % 32;
  const int unit = blockIdx.x / G;                    // (d-1) * 2 + mh
  const int d = unit / 2 + 1, mh = unit % 2;
  const int j0 = blockIdx.x % G;
  const int n_in = (d == 1) ? 64 : kW;                // columns of a_{d-1}
  const uint32_t b_bytes = (d == 1) ? kA0Bytes : kActBytes;
  uint8_t* s_ones = smem;                             // 4 KB of bf16 1.0
  uint8_t* s_stage = smem + 4096;
  for (int t = threadIdx.x; t < 1024; t += kWgThreads) reinterpret_cast<uint32_t*>(s_ones)[t] = 0x3F803F80u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgRing; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); }
    mbar_init(&bars.done, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&bars.tmem_base, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = bars.tmem_base;
  int64_t n_tiles = lay_tiles;
  if (n_dev) { const int64_t t = ((int64_t)*n_dev + kTile - 1) / kTile; n_tiles = t < n_tiles ? t : n_tiles; }
  const int64_t my_tiles = (n_tiles > j0) ? (n_tiles - j0 + G - 1) / G : 0;
  const uint8_t* d_base = delta + (int64_t)(d - 1) * lay_tiles * kActBytes + mh * 32768;
  const uint8_t* a_base = (d == 1) ? saved : saved + lay_tiles * kA0Bytes + (int64_t)(d - 2) * lay_tiles * kActBytes;

  if (warp == 0 && lane == 0) {
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int s = (int)(t % kWgRing);
      mbar_wait(&bars.empty[s], (uint32_t)(((t / kWgRing) & 1) ^ 1));
      const int64_t tile = j0 + t * G;
      mbar_arrive_expect_tx(&bars.full[s], 32768 + b_bytes);
      bulk_g2s(s_stage + s * kWgStage, d_base + tile * kActBytes, 32768, &bars.full[s]);
      bulk_g2s(s_stage + s * kWgStage + 32768, a_base + tile * (int64_t)b_bytes, b_bytes, &bars.full[s]);
    }
  } else if (warp == 1 && lane == 0) {
    // D[out 128 x in] += delta^T a (both MN-major), Db[out 128 x 16] += delta^T ones
    const uint32_t idesc_w = make_idesc_bf16(128, n_in, 1, 1);
    const uint32_t idesc_b = make_idesc_bf16(128, 16, 1, 0);
    const uint32_t ones = smem_u32(s_ones);
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int s = (int)(t % kWgRing);
      mbar_wait(&bars.full[s], (uint32_t)((t / kWgRing) & 1));
      fence_after_sync();
      const uint32_t da = smem_u32(s_stage + s * kWgStage), ba = da + 32768;
      for (int k = 0; k < kTile / 16; ++k) {
        const uint64_t desc_a = make_smem_desc_sw128(da + k * 2048, 16384, 1024);
        mma_ss(tmem, desc_a, make_smem_desc_sw128(ba + k * 2048, 16384, 1024), idesc_w, (t | k) != 0);
        mma_ss(tmem + 256, desc_a, make_smem_desc_sw128(ones + (k / 4) * 2048 + (k % 4) * 32, 16, 1024), idesc_b, (t | k) != 0);
      }
      mma_commit(&bars.empty[s]);
    }
    mma_commit(&bars.done);
  }
  __syncwarp();
  mbar_wait(&bars.done, 0);
  fence_after_sync();
  {
    const int row = warp * 32 + lane;                 // out feature 128 mh + row
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    float* Pp = partials + (int64_t)blockIdx.x * kWgPartial;
    if (my_tiles > 0) {
      for (int c0 = 0; c0 < n_in; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem + lane_off + c0, r);
        wait_ld();
#pragma unroll
        for (int c = 0; c < 32; c += 4)
          *reinterpret_cast<float4*>(Pp + row * 256 + c0 + c) =
              make_float4(__uint_as_float(r[c]), __uint_as_float(r[c + 1]), __uint_as_float(r[c + 2]), __uint_as_float(r[c + 3]));
      }
      uint32_t b;
      tmem_ld1(tmem + lane_off + 256, b);
      wait_ld();
      Pp[128 * 256 + row] = __uint_as_float(b);
    } else {
      for (int c = 0; c < n_in; ++c) Pp[row * 256 + c] = 0.0f;
      Pp[128 * 256 + row] = 0.0f;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// fixed-order reduction over the G slices, scattered into the reference parameter layout
__global__ void __launch_bounds__(256) wgrad256_reduce_kernel(const float* __restrict__ partials, int G, MlpLayout Lay, Plan256 P,
                                                              float* __restrict__ grad) {
  const int w = blockIdx.y;                                 // linear layer 0..L
  const int e = blockIdx.x * blockDim.x + threadIdx.x;      // element of [256 x 256] (+256 bias)
  if (e >= kW * kW + kW) return;
  const bool is_bias = e >= kW * kW;
  const int o = is_bias ? e - kW * kW : e / kW, c = is_bias ? 0 : e % kW;
  if (!is_bias && w == 0 && c >= 64) return;                // layer 0 accumulates only 64 feature columns
  const int mh = o / 128, r = o % 128;
  const float* base = partials + ((int64_t)(w * 2 + mh) * G) * kWgPartial + (is_bias ? 128 * 256 + r : r * 256 + c);
  float acc = 0.0f;
  for (int g = 0; g < G; ++g) acc += base[(int64_t)g * kWgPartial];
  if (is_bias) { grad[Lay.off_b[w] + o] = acc; return; }
  if (w > 0) { grad[Lay.off_w[w] + (int64_t)o * kW + c] = acc; return; }
  if (c >= P.k0) return;
  if (c < 3) {                                              // x_hi and x_lo both belong to reference column c
    float lo = 0.0f;
    for (int g = 0; g < G; ++g) lo += base[(int64_t)g * kWgPartial + 3];
    grad[Lay.off_w[0] + (int64_t)o * Lay.d_in + c] = acc + lo;
  } else if (c >= 6) {
    grad[Lay.off_w[0] + (int64_t)o * Lay.d_in + ref_col0(c, P.basis)] = acc;
  }
}

// output layer: dw_out[o] = sum_s g[s] a_{L+1}[s][o], db_out = sum_s g[s]; streams the a_{L+1} tile images (512 B/sample)
__global__ void __launch_bounds__(256) outgrad256_partial_kernel(const uint8_t* __restrict__ a_last, const float* __restrict__ g, int64_t n,
                                                                 const int32_t* __restrict__ n_dev, float* __restrict__ partials /*[gridDim.x][260]*/) {
  __shared__ float s_acc[8][260];
  if (n_dev) { const int64_t nd = *n_dev; n = nd < n ? nd : n; }
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  const int cidx = threadIdx.x % 32;            // logical 16-byte chunk: columns [8 cidx, 8 cidx + 8)
  const int rg = threadIdx.x / 32;              // rows rg, rg + 8, ..., rg + 120
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float gsum = 0.0f;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint8_t* img = a_last + tile * kActBytes + (cidx / 8) * 16384;
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
      const int r = rg + 8 * k;
      const int64_t i = tile * kTile + r;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(img + r * 128 + (((cidx % 8) ^ (r & 7)) << 4)));
      const float gv = (i < n) ? g[i] : 0.0f;
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
        acc[2 * e] = fmaf(gv, f.x, acc[2 * e]);
        acc[2 * e + 1] = fmaf(gv, f.y, acc[2 * e + 1]);
      }
      gsum += gv;
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) s_acc[rg][cidx * 8 + e] = acc[e];
  if (cidx == 0) s_acc[rg][256] = gsum;
  __syncthreads();
  for (int c = threadIdx.x; c < 257; c += 256) {
    float v = 0.0f;
    for (int r = 0; r < 8; ++r) v += s_acc[r][c];
    partials[(int64_t)blockIdx.x * 260 + c] = v;
  }
}

// out[e] = sum_p partials[p][e] for e < len: one CTA per element, fixed-order tree over the partials
__global__ void __launch_bounds__(128) small_reduce256_kernel(const float* __restrict__ partials, int n_part, int stride, int len,
                                                              float* __restrict__ out0, int split, float* __restrict__ out1) {
  __shared__ float s_v[128];
  const int e = blockIdx.x;
  if (e >= len) return;
  float v = 0.0f;
  for (int p = threadIdx.x; p < n_part; p += 128) v += partials[(int64_t)p * stride + e];
  s_v[threadIdx.x] = v;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_v[threadIdx.x] += s_v[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) { if (e < split) out0[e] = s_v[0]; else out1[e - split] = s_v[0]; }
}

inline int64_t align256(int64_t b) { return (b + 255) / 256 * 256; }
constexpr int kOutgradBlocks = 592;

template <class K>
int ensure_smem(K kernel, size_t smem, size_t* cached) {
  if (*cached >= smem) return 0;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    cudaGetLastError();
    angio::set_error("cudaFuncSetAttribute(%zu bytes): %s", smem, cudaGetErrorString(e));
    return (int)e;
  }
  *cached = smem;
  return 0;
}

constexpr size_t kSmemFwd = (size_t)kRing * kChunk + 20480 + 1024;                      // ring | constants (<= 20 KB)
constexpr size_t kSmemTrain = (size_t)kRing * kChunk + 20480 + kStageBytes + 1024;      // ... | 8 x 4 KB tile-image staging

template <int MODE, bool TRAIN>
int launch_fwd(const Plan256& P, const void* packed, const angio_samples& in, float* out, void* saved, cudaStream_t st) {
  const size_t smem = TRAIN ? kSmemTrain : kSmemFwd;
  static size_t cached = 0;
  if (int rc = ensure_smem(mlp256_fwd_kernel<MODE, TRAIN>, smem, &cached)) return rc;
  const int64_t n_tiles = (in.n + kTile - 1) / kTile;
  int grid = angio::sm_count();
  if (n_tiles < grid) grid = (int)n_tiles;
  angio::note_launch(TRAIN ? "mlp256_fwd_kernel<LOGIT,train>" : MODE == ANGIO_OUT_ALPHA ? "mlp256_fwd_kernel<ALPHA>"
                     : MODE == ANGIO_OUT_SIGMA ? "mlp256_fwd_kernel<SIGMA>" : "mlp256_fwd_kernel<LOGIT>");
  mlp256_fwd_kernel<MODE, TRAIN><<<grid, kThreads, smem, st>>>(reinterpret_cast<const uint8_t*>(packed), P, in, out, reinterpret_cast<uint8_t*>(saved));
  return angio::finish_launch("mlp256_fwd_kernel");
}

int wgrad_groups(const MlpLayout& L) {
  const int g = angio::sm_count() / (2 * (L.n_hidden + 1));
  return g > 0 ? g : 1;
}

}  // namespace

namespace angio {

bool tc256_supported(const MlpLayout& L) {
  Plan256 P;
  return make_plan(L, &P) && (P.n_const * 4 + 15) / 16 * 16 <= 20480;
}
int64_t tc256_packed_bytes(const MlpLayout& L) {
  Plan256 P;
  return make_plan(L, &P) ? P.total_bytes : 0;
}
int64_t tc256_workspace_bytes(const MlpLayout& L, int64_t n, int training) {
  if (!training) return 256;
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  const int G = wgrad_groups(L);
  return align256((int64_t)(L.n_hidden + 1) * n_tiles * kActBytes)               // delta images
         + align256((int64_t)(L.n_hidden + 1) * 2 * G * kWgPartial * 4)           // wgrad partials
         + align256((int64_t)sm_count() * 32 * 4)                                // coefficient-gradient partials
         + align256((int64_t)kOutgradBlocks * 260 * 4) + 256;                    // output-layer partials
}
int64_t tc256_saved_bytes(const MlpLayout& L, int64_t n) {
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  return n_tiles * (kA0Bytes + (int64_t)(L.n_hidden + 1) * (kActBytes + kMaskBytes)) + 256;
}

int tc256_pack_weights(const MlpLayout& L, const float* params, void* packed, cudaStream_t st) {
  Plan256 P;
  if (!make_plan(L, &P)) { set_error("tc256_pack_weights: unsupported shape"); return ANGIO_ERR_UNSUPPORTED; }
  const int64_t total = 2 * ((int64_t)kW * 64 + (int64_t)P.n_hidden * kW * kW);
  note_launch("pack256_kernel");
  pack256_kernel<<<blocks_for(total, 256), 256, 0, st>>>(params, L, P, reinterpret_cast<uint8_t*>(packed));
  return finish_launch("tc256_pack_weights");
}

int tc256_forward(const MlpLayout& L, const void* packed, const angio_samples& in, int out_mode, float* out, void* saved, cudaStream_t st) {
  Plan256 P;
  if (!make_plan(L, &P)) { set_error("tc256_forward: unsupported shape"); return ANGIO_ERR_UNSUPPORTED; }
  if (in.n == 0) return 0;
  if ((reinterpret_cast<uintptr_t>(packed) & 15) != 0) { set_error("tc256_forward: packed image must be 16-byte aligned"); return ANGIO_ERR_INVALID_ARG; }
  if (saved) {
    if ((reinterpret_cast<uintptr_t>(saved) & 15) != 0) { set_error("tc256_forward: saved buffer must be 16-byte aligned"); return ANGIO_ERR_INVALID_ARG; }
    if (out_mode != ANGIO_OUT_LOGIT) { set_error("tc256_forward: training forward returns logits only"); return ANGIO_ERR_INVALID_ARG; }
    return launch_fwd<ANGIO_OUT_LOGIT, true>(P, packed, in, out, saved, st);
  }
  switch (out_mode) {
    case ANGIO_OUT_LOGIT: return launch_fwd<ANGIO_OUT_LOGIT, false>(P, packed, in, out, nullptr, st);
    case ANGIO_OUT_SIGMA: return launch_fwd<ANGIO_OUT_SIGMA, false>(P, packed, in, out, nullptr, st);
    default: return launch_fwd<ANGIO_OUT_ALPHA, false>(P, packed, in, out, nullptr, st);
  }
}

int tc256_backward(const MlpLayout& L, const void* packed, const angio_samples& in, const void* saved, const float* grad_out,
                   float* grad_params, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  Plan256 P;
  if (!make_plan(L, &P)) { set_error("tc256_backward: unsupported shape"); return ANGIO_ERR_UNSUPPORTED; }
  cudaError_t ce = cudaMemsetAsync(grad_params, 0, L.total * 4, st);
  if (ce != cudaSuccess) { set_error("memset grad_params: %s", cudaGetErrorString(ce)); return (int)ce; }
  const int64_t n = in.n;
  if (n == 0) return 0;
  const int64_t need = tc256_workspace_bytes(L, n, 1) - 256;
  if (!workspace || workspace_bytes < need) {
    set_error("angio_mlp_backward(bf16, width 256): workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need);
    return ANGIO_ERR_WORKSPACE;
  }
  if ((reinterpret_cast<uintptr_t>(workspace) & 15) != 0 || (reinterpret_cast<uintptr_t>(saved) & 15) != 0) {
    set_error("tc256_backward: saved / workspace must be 16-byte aligned");
    return ANGIO_ERR_INVALID_ARG;
  }
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  const int G = wgrad_groups(L);
  const int nl = L.n_hidden + 1;
  char* wb = reinterpret_cast<char*>(workspace);
  uint8_t* delta = reinterpret_cast<uint8_t*>(wb); wb += align256((int64_t)nl * n_tiles * kActBytes);
  float* wpart = reinterpret_cast<float*>(wb); wb += align256((int64_t)nl * 2 * G * kWgPartial * 4);
  float* cpart = reinterpret_cast<float*>(wb); wb += align256((int64_t)sm_count() * 32 * 4);
  float* opart = reinterpret_cast<float*>(wb);
  const uint8_t* sv = reinterpret_cast<const uint8_t*>(saved);
  {
    static size_t cached = 0;
    if (int rc = ensure_smem(mlp256_dgrad_kernel, kSmemTrain, &cached)) return rc;
    int grid = sm_count();
    if (n_tiles < grid) grid = (int)n_tiles;
    note_launch("mlp256_dgrad_kernel");
    mlp256_dgrad_kernel<<<grid, kThreads, kSmemTrain, st>>>(reinterpret_cast<const uint8_t*>(packed), P, in, sv, grad_out, delta, cpart);
    if (int rc = finish_launch("mlp256_dgrad_kernel")) return rc;
    if (L.enc) {
      note_launch("small_reduce256_kernel");
      small_reduce256_kernel<<<3 * L.basis, 128, 0, st>>>(cpart, grid, 32, 3 * L.basis, grad_params + L.off_coef, 3 * L.basis, nullptr);
    }
  }
  {
    const size_t smem = 4096 + (size_t)kWgRing * kWgStage + 1024;
    static size_t cached = 0;
    if (int rc = ensure_smem(mlp256_wgrad_kernel, smem, &cached)) return rc;
    note_launch("mlp256_wgrad_kernel");
    mlp256_wgrad_kernel<<<nl * 2 * G, kWgThreads, smem, st>>>(sv, delta, n_tiles, in.n_dev, L.n_hidden, G, wpart);
    if (int rc = finish_launch("mlp256_wgrad_kernel")) return rc;
    note_launch("wgrad256_reduce_kernel");
    wgrad256_reduce_kernel<<<dim3((kW * kW + kW + 255) / 256, nl), 256, 0, st>>>(wpart, G, L, P, grad_params);
  }
  {
    const uint8_t* a_last = sv + n_tiles * kA0Bytes + (int64_t)L.n_hidden * n_tiles * kActBytes;
    note_launch("outgrad256_partial_kernel");
    outgrad256_partial_kernel<<<kOutgradBlocks, 256, 0, st>>>(a_last, grad_out, n, in.n_dev, opart);
    const int lo = L.n_linear - 1;
    note_launch("small_reduce256_kernel");
    small_reduce256_kernel<<<257, 128, 0, st>>>(opart, kOutgradBlocks, 260, 257, grad_params + L.off_w[lo], 256, grad_params + L.off_b[lo]);
  }
  return finish_launch("tc256_backward");
}

}  // namespace angio
