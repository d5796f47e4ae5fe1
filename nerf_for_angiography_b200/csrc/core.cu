// core.cu -- version, error reporting, device queries
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

#include <atomic>
#include <vector>
namespace angio {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

// Per-launch timeline (bench.py's per-kernel rooflines): while a profile is active every kernel launch of this library first
// records a CUDA event on the profiled stream, so the time between two consecutive events is the device time of the earlier
// launch (plus whatever the caller enqueued in between).  Events are pooled; nothing is recorded when no profile is active.
struct Profile {
  bool on = false;
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> events;      // pool, grows on demand
  std::vector<const char*> names;       // one per recorded launch (string literals)
  size_t used = 0;                      // events recorded so far (launches + the closing event)
};
static Profile g_prof;

static bool prof_record() {
  if (g_prof.used == g_prof.events.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return false;
    g_prof.events.push_back(e);
  }
  return cudaEventRecord(g_prof.events[g_prof.used++], g_prof.stream) == cudaSuccess;
}

void note_launch(const char* name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (g_prof.on && prof_record()) g_prof.names.push_back(name);
}

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

extern "C" int angio_profile_start(void* stream) {
  angio::g_prof.on = true;
  angio::g_prof.stream = angio::as_stream(stream);
  angio::g_prof.used = 0;
  angio::g_prof.names.clear();
  return 0;
}
extern "C" int64_t angio_profile_stop(void) {
  if (!angio::g_prof.on) { angio::set_error("angio_profile_stop: no profile is active"); return ANGIO_ERR_INVALID_ARG; }
  angio::g_prof.on = false;
  if (angio::g_prof.names.empty()) return 0;
  if (!angio::prof_record()) { angio::set_error("angio_profile_stop: cannot record the closing event"); return ANGIO_ERR_INVALID_ARG; }
  cudaError_t e = cudaEventSynchronize(angio::g_prof.events[angio::g_prof.used - 1]);
  if (e != cudaSuccess) { angio::set_error("angio_profile_stop: %s", cudaGetErrorString(e)); return (int)e; }
  return (int64_t)angio::g_prof.names.size();
}
extern "C" int angio_profile_entry(int64_t i, char* name_out, int32_t name_cap, float* ms_out) {
  if (angio::g_prof.on || i < 0 || (size_t)i >= angio::g_prof.names.size() || (size_t)i + 1 >= angio::g_prof.used || !name_out || name_cap <= 0 || !ms_out) {
    angio::set_error("angio_profile_entry: bad index or no finished profile");
    return ANGIO_ERR_INVALID_ARG;
  }
  snprintf(name_out, (size_t)name_cap, "%s", angio::g_prof.names[i]);
  cudaError_t e = cudaEventElapsedTime(ms_out, angio::g_prof.events[i], angio::g_prof.events[i + 1]);
  if (e != cudaSuccess) { angio::set_error("angio_profile_entry: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}
