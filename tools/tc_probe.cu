// tools/tc_probe.cu -- bring-up probe for the tcgen05 building blocks used by csrc/mlp_tc.cu.
// Each variant checks one layout/descriptor assumption against an exact integer-valued host reference
// (or measures issue throughput).  Run one variant per process so a trap cannot poison later variants:
//   ./tc_probe <variant>
//     0  SS  A K-major,  B K-major      (forward GEMM view)
//     1  TS  A in TMEM,  B K-major      (forward, activations fed back through TMEM)
//     2  SS  A K-major,  B MN-major     (dgrad view of the same weight tile)
//     3  SS  A MN-major, B MN-major     (wgrad view: reduction over the sample rows)
//     4..7 throughput: 4 SS N=128, 5 TS N=128, 6 SS N=256, 7 TS N=256
//     8  SS  bf16 inputs, D format F16 (is it legal? halves the accumulator drain)
//     9  LDTM (tcgen05.ld) bandwidth: 4 / 8 warps reading 128 columns repeatedly
//     10..13 CTA-pair throughput (cta_group::2, M = 256): 10 SS N=128, 11 TS N=128, 12 SS N=256, 13 TS N=256
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../nerf_for_angiography_b200/csrc/tc05.cuh"

using namespace tc05;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

static constexpr int M = 128, N = 128, K = 64;

// a_img / b_img: 16 KB smem images (already swizzled).  D out: [128][128] fp32.
__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t* a_img, const uint8_t* b_img, float* D, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;               // 16 KB
  uint8_t* sB = smem + 16384;       // 16 KB
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  if (threadIdx.x == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 256); tmem_relinquish(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_d = tmem, tmem_a = tmem + 128;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_load, 32768);
    bulk_g2s(sA, a_img, 16384, &bar_load);
    bulk_g2s(sB, b_img, 16384, &bar_load);
  }
  mbar_wait(&bar_load, 0);

  if (variant == 1) {
    // A K-major image in smem -> registers -> TMEM (row = lane, packed bf16x2 per column)
    const int row = warp * 32 + lane;
    uint32_t v[16];
    for (int half = 0; half < 2; ++half) {
      for (int c = 0; c < 16; ++c) {
        int col = (half * 16 + c) * 2;
        v[c] = *reinterpret_cast<const uint32_t*>(sA + sw128_offset(row, col));
      }
      tmem_st16(tmem_a + ((uint32_t)(warp * 32) << 16) + half * 16, v);
    }
    wait_st();
    fence_before_sync();
  }
  __syncthreads();
  fence_after_sync();

  if (warp == 0 && lane == 0) {
    const uint32_t a_mn = (variant == 3), b_mn = (variant == 2 || variant == 3);
    uint32_t idesc = make_idesc_bf16(M, N, a_mn, b_mn);
    if (variant == 8) idesc &= ~(3u << 4);   // D format 0 = F16
    for (int k = 0; k < K / 16; ++k) {
      // K-major: +32 B per k-step inside the 128-B row; MN-major: +16 rows * 128 B per k-step
      uint64_t da = a_mn ? make_smem_desc_sw128(smem_u32(sA) + k * 2048, 8192, 1024)
                         : make_smem_desc_sw128(smem_u32(sA) + k * 32, 16, 1024);
      uint64_t db = b_mn ? make_smem_desc_sw128(smem_u32(sB) + k * 2048, 8192, 1024)
                         : make_smem_desc_sw128(smem_u32(sB) + k * 32, 16, 1024);
      if (variant == 1) mma_ts(tmem_d, tmem_a + k * 8, db, idesc, k > 0);
      else mma_ss(tmem_d, da, db, idesc, k > 0);
    }
    mma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  fence_after_sync();

  const int row = warp * 32 + lane;
  if (variant == 8) {
    for (int c0 = 0; c0 < N / 2; c0 += 32) {   // packed f16x2: column c holds elements (2c, 2c+1)?
      uint32_t r[32];
      tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, r);
      wait_ld();
      for (int j = 0; j < 32; ++j) {
        __half2 h = *reinterpret_cast<__half2*>(&r[j]);
        D[row * N + 2 * (c0 + j)] = __low2float(h);
        D[row * N + 2 * (c0 + j) + 1] = __high2float(h);
      }
    }
  } else
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, r);
    wait_ld();
    for (int j = 0; j < 32; ++j) D[row * N + c0 + j] = __uint_as_float(r[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// throughput: each CTA issues iters * 8 MMAs (K = 128 per accumulation group) and waits once at the end
template <int NN>
__global__ void __launch_bounds__(128, 1) perf_kernel(int iters, int ts_mode, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                 // 2 x 16 KB (K = 128)
  uint8_t* sB = smem + 32768;         // NN rows x 128 K: 2 blocks of NN*128 B
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x / 32;
  for (int i = threadIdx.x; i < (32768 + NN * 256) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar_mma, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, NN, 0, 0);
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < 8; ++k) {
        uint32_t a_off = (k / 4) * 16384 + (k % 4) * 32;
        uint32_t b_off = (k / 4) * (NN * 128) + (k % 4) * 32;
        uint64_t da = make_smem_desc_sw128(smem_u32(sA) + a_off, 16, 1024);
        uint64_t db = make_smem_desc_sw128(smem_u32(sB) + b_off, 16, 1024);
        if (ts_mode) mma_ts(tmem, tmem + 256 + k * 8, db, idesc, k > 0);
        else mma_ss(tmem, da, db, idesc, k > 0);
      }
    }
    mma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  fence_after_sync();
  if (threadIdx.x == 0 && sink) sink[blockIdx.x] = 1.0f;
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}


// CTA-pair throughput (cta_group::2): M = 256 over two SMs (128 rows each), each CTA holds NN/2 rows of B; only the leader issues.
// Per SM the FLOPs per instruction equal the single-CTA M = 128 case, but every SM reads half of the B operand.
template <int NN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) perf_pair_kernel(int iters, int ts_mode, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                 // 2 x 16 KB (K = 128), this CTA's 128 rows
  uint8_t* sB = smem + 32768;         // NN/2 rows x 128 K: 2 blocks of (NN/2)*128 B
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x / 32;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < (32768 + NN * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar_mma, 1); fence_mbar_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  fence_before_sync();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(256, NN, 0, 0);
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < 8; ++k) {
        uint32_t a_off = (k / 4) * 16384 + (k % 4) * 32;
        uint32_t b_off = (k / 4) * (NN / 2 * 128) + (k % 4) * 32;
        uint64_t da = make_smem_desc_sw128(smem_u32(sA) + a_off, 16, 1024);
        uint64_t db = make_smem_desc_sw128(smem_u32(sB) + b_off, 16, 1024);
        const uint32_t acc = k > 0;
        if (ts_mode)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(tmem + 256 + k * 8), "l"(db),
                       "r"(idesc), "r"(acc) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
                       : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(&bar_mma)), "h"((uint16_t)3) : "memory");
  }
  mbar_wait(&bar_mma, 0);
  fence_after_sync();
  if (threadIdx.x == 0 && sink) sink[blockIdx.x] = 1.0f;
  fence_before_sync();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// issue-cadence experiments (TS mode, K = 128 per accumulation group):
//   mode 0: one issuer, one accumulator (baseline = variant 5)      mode 1: one issuer alternating two accumulators per k-step
//   mode 2: two issuer warps, one accumulator each                  mode 3: one issuer, accumulate flag always 0 (no RAW chain)
//   mode 4: one issuer, N = 64 instructions                          mode 5: two issuers, N = 256, one accumulator each
__global__ void __launch_bounds__(128, 1) cadence_kernel(int iters, int mode, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem;                 // 256 rows x 128 K
  __shared__ uint64_t bar_mma[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int i = threadIdx.x; i < (256 * 256) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar_mma[0], 1); mbar_init(&bar_mma[1], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const int NN = (mode == 4) ? 64 : (mode == 5 ? 256 : 128);
  const uint32_t idesc = make_idesc_bf16(128, NN, 0, 0);
  const bool issuer = (lane == 0) && (warp == 0 || ((mode == 2 || mode == 5) && warp == 1));
  long long t0 = clock64();
  if (issuer) {
    // accumulators: mode 5 uses [0,256) and... only 512 columns exist, A lives in the accumulator of the other warp's
    // region tail; values are irrelevant for timing
    const uint32_t d0 = tmem + (mode == 5 ? 0 : warp * 128);
    const uint32_t a0 = tmem + 384 + warp * 64;
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < 8; ++k) {
        uint32_t b_off = (k / 4) * (NN * 128) + (k % 4) * 32;
        uint64_t db = make_smem_desc_sw128(smem_u32(sB) + b_off, 16, 1024);
        uint32_t d = d0;
        if (mode == 1) d = tmem + (k & 1) * 128;
        uint32_t acc = (mode == 3) ? 0u : (uint32_t)(mode == 1 ? k > 1 : k > 0);
        mma_ts(d, a0 + k * 8, db, idesc, acc);
      }
    }
    mma_commit(&bar_mma[warp]);
  }
  if (warp == 0) mbar_wait(&bar_mma[0], 0);
  if (warp == 1 && (mode == 2 || mode == 5)) mbar_wait(&bar_mma[1], 0);
  long long t1 = clock64();
  fence_after_sync();
  if (lane == 0 && warp < 2 && cycles) cycles[blockIdx.x * 2 + warp] = t1 - t0;
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}


// drain/restart experiment: groups of G MMAs, commit, wait for completion, repeat.  Reports cycles per group measured by the
// issuing thread: [issue of the G MMAs + commit] and [commit -> mbarrier phase observed].
__global__ void __launch_bounds__(128, 1) drain_kernel(int iters, int G, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem;
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x / 32;
  for (int i = threadIdx.x; i < (128 * 256) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar_mma, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
  if (threadIdx.x == 0) {
    long long t_issue = 0, t_wait = 0;
    for (int it = 0; it < iters; ++it) {
      long long a = clock64();
      for (int k = 0; k < G; ++k) {
        uint32_t b_off = (k / 4) * (128 * 128) + (k % 4) * 32;
        mma_ts(tmem, tmem + 384 + k * 8, make_smem_desc_sw128(smem_u32(sB) + b_off, 16, 1024), idesc, k > 0);
      }
      mma_commit(&bar_mma);
      long long b = clock64();
      mbar_wait(&bar_mma, it & 1);
      fence_after_sync();
      long long c = clock64();
      t_issue += b - a; t_wait += c - b;
    }
    cycles[0] = t_issue; cycles[1] = t_wait;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

__global__ void __launch_bounds__(256, 1) ldtm_bw_kernel(int iters, unsigned* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x / 32;
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s + ((uint32_t)((warp % 4) * 32) << 16) + (warp / 4) * 128;
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint32_t r0[32], r1[32], r2[32], r3[32];
    tmem_ld32(tmem, r0); tmem_ld32(tmem + 32, r1); tmem_ld32(tmem + 64, r2); tmem_ld32(tmem + 96, r3);
    wait_ld();
    for (int j = 0; j < 32; ++j) acc += r0[j] ^ r1[j] ^ r2[j] ^ r3[j];
  }
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }  // exact for small ints

int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs=%d variant=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, variant);

  if (variant >= 10 && variant <= 13) {
    const int NN = (variant >= 12) ? 256 : 128, ts = (variant & 1);
    const int iters = 4096;
    size_t smem = 32768 + NN * 128 + 1024;
    const int blocks = prop.multiProcessorCount & ~1;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      if (NN == 128) {
        CK(cudaFuncSetAttribute(perf_pair_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        perf_pair_kernel<128><<<blocks, 128, smem>>>(iters, ts, nullptr);
      } else {
        CK(cudaFuncSetAttribute(perf_pair_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        perf_pair_kernel<256><<<blocks, 128, smem>>>(iters, ts, nullptr);
      }
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
      CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    double flops = 2.0 * 128 * NN * 16 * 8.0 * iters * blocks;
    printf("PERF-PAIR variant=%d mode=%s M=256 N=%d: %.3f ms  %.1f TFLOP/s\n", variant, ts ? "TS" : "SS", NN, ms, flops / ms * 1e-9);
    return 0;
  }

  if (variant == 9) {
    for (int nw = 4; nw <= 8; nw += 4) {
      const int iters = 20000;
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        ldtm_bw_kernel<<<prop.multiProcessorCount, nw * 32>>>(iters, nullptr);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        CK(cudaEventElapsedTime(&ms, e0, e1));
      }
      double bytes_per_sm = (double)iters * nw * 32 * 128 * 4;
      int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
      printf("LDTM warps=%d: %.3f ms, %.1f GB/s per SM, %.1f B/clk/SM at %.0f MHz (nominal max clock)\n", nw, ms, bytes_per_sm / ms * 1e-6,
             bytes_per_sm / (ms * 1e-3 * khz * 1e3), khz * 1e-3);
    }
    return 0;
  }

  if (variant >= 10 && variant <= 15) {
    const int mode = variant - 10, iters = 2048;
    size_t smem = 65536 + 1024;
    CK(cudaFuncSetAttribute(cadence_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long* dcy; CK(cudaMalloc(&dcy, sizeof(long long) * 2 * prop.multiProcessorCount));
    for (int nblk : {1, prop.multiProcessorCount}) {
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        cadence_kernel<<<nblk, 128, smem>>>(iters, mode, dcy);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        CK(cudaEventElapsedTime(&ms, e0, e1));
      }
      long long cy[2]; CK(cudaMemcpy(cy, dcy, sizeof(cy), cudaMemcpyDeviceToHost));
      const int n_issuers = (mode == 2 || mode == 5) ? 2 : 1;
      const double n_mma = (double)iters * 8 * n_issuers;
      printf("CADENCE mode=%d blocks=%d: %.3f ms, %.1f cycles per MMA (CTA 0, all issuers), kernel-level %.1f ns per MMA\n", mode, nblk, ms,
             (double)cy[0] / n_mma, ms * 1e6 / n_mma);
    }
    return 0;
  }

  if (variant == 16) {
    size_t smem = 32768 + 1024;
    CK(cudaFuncSetAttribute(drain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long* dcy; CK(cudaMalloc(&dcy, sizeof(long long) * 2));
    for (int G : {1, 2, 3, 4, 8}) {
      const int iters = 1000;
      drain_kernel<<<1, 128, smem>>>(iters, G, dcy);
      CK(cudaDeviceSynchronize());
      long long cy[2]; CK(cudaMemcpy(cy, dcy, sizeof(cy), cudaMemcpyDeviceToHost));
      printf("DRAIN G=%d: issue+commit %.1f cycles, commit->observed %.1f cycles, total %.1f per group\n", G, (double)cy[0] / iters,
             (double)cy[1] / iters, (double)(cy[0] + cy[1]) / iters);
    }
    return 0;
  }
  if (variant >= 4 && variant < 8) {
    const int NN = (variant >= 6) ? 256 : 128, ts = (variant & 1);
    const int iters = 4096;
    size_t smem = 32768 + NN * 256 + 1024;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      if (NN == 128) {
        CK(cudaFuncSetAttribute(perf_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        perf_kernel<128><<<prop.multiProcessorCount, 128, smem>>>(iters, ts, nullptr);
      } else {
        CK(cudaFuncSetAttribute(perf_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        perf_kernel<256><<<prop.multiProcessorCount, 128, smem>>>(iters, ts, nullptr);
      }
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
      CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    double flops = 2.0 * 128 * NN * 16 * 8.0 * iters * prop.multiProcessorCount;
    printf("PERF variant=%d mode=%s N=%d: %.3f ms  %.1f TFLOP/s\n", variant, ts ? "TS" : "SS", NN, ms, flops / ms * 1e-9);
    return 0;
  }

  // logical operands: A[m][k], B[n][k]; small integers => exact in bf16 and fp32
  std::vector<float> A(M * K), B(N * K), Dref(M * N, 0.f);
  srand(123 + variant);
  for (auto& v : A) v = (float)(rand() % 5 - 2);
  for (auto& v : B) v = (float)(rand() % 5 - 2);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k]; Dref[m * N + n] = s; }

  std::vector<uint8_t> a_img(16384, 0), b_img(16384, 0);
  auto put = [](std::vector<uint8_t>& img, uint32_t off, float v) { uint16_t h = f2bf(v); memcpy(&img[off], &h, 2); };
  const bool a_mn = (variant == 3), b_mn = (variant == 2 || variant == 3);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
    uint32_t off = a_mn ? (uint32_t)(m / 64) * 8192 + sw128_offset(k, m % 64) : sw128_offset(m, k);
    put(a_img, off, A[m * K + k]);
  }
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
    uint32_t off = b_mn ? (uint32_t)(n / 64) * 8192 + sw128_offset(k, n % 64) : sw128_offset(n, k);
    put(b_img, off, B[n * K + k]);
  }
  uint8_t *da, *db; float* dD;
  CK(cudaMalloc(&da, 16384)); CK(cudaMalloc(&db, 16384)); CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMemcpy(da, a_img.data(), 16384, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b_img.data(), 16384, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, M * N * 4));
  size_t smem = 32768 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(da, db, dD, variant);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D(M * N);
  CK(cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
  int bad = 0; double maxerr = 0;
  for (int i = 0; i < M * N; ++i) { double e = fabs((double)D[i] - Dref[i]); if (!(e <= 0)) { if (bad < 5) printf("  mismatch at m=%d n=%d got %f want %f\n", i / N, i % N, D[i], Dref[i]); ++bad; } if (e > maxerr) maxerr = e; }
  printf("PROBE variant=%d %s  mismatches=%d maxerr=%g\n", variant, bad ? "FAIL" : "PASS", bad, maxerr);
  return bad ? 1 : 0;
}
