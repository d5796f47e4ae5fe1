// core.cu -- version, error reporting, device queries
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

#include <atomic>
namespace angio {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}
}  // namespace angio

extern "C" int angio_version(void) { return ANGIO_B200_VERSION; }
extern "C" const char* angio_last_error_string(void) { return angio::g_err; }
extern "C" int angio_sm_count(void) { return angio::sm_count(); }
extern "C" int64_t angio_launch_count(void) { return (int64_t)angio::g_launches.load(); }
