// common.cuh -- shared host/device helpers for libangio_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/angio_b200.h"

namespace angio {

// thread-local last-error text (angio_last_error_string)
void set_error(const char* fmt, ...);
int sm_count();
void note_launch(const char* kernel_name);  // counts kernel launches (angio_launch_count) and feeds the per-launch timeline (angio_profile_*)

inline int finish_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

#define ANGIO_REQUIRE(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      angio::set_error(__VA_ARGS__);    \
      return ANGIO_ERR_INVALID_ARG;     \
    }                                   \
  } while (0)

#define ANGIO_CUDA(call)                                                         \
  do {                                                                           \
    cudaError_t e_ = (call);                                                     \
    if (e_ != cudaSuccess) {                                                     \
      angio::set_error("%s failed: %s", #call, cudaGetErrorString(e_));          \
      return (int)e_;                                                            \
    }                                                                            \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline int blocks_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

struct Roi {
  float lo[3];
  float hi[3];
};
inline Roi make_roi(const float* h) {
  Roi r;
  for (int i = 0; i < 3; ++i) { r.lo[i] = h[i]; r.hi[i] = h[3 + i]; }
  return r;
}

__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace angio
