// Library-wide runtime state: last-error string, SM count, launch counter.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <vector>

namespace fervit {

static thread_local char g_error[1024] = "";
unsigned long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_error; }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("FERVIT_PDL"); on = (e && atoi(e) == 0) ? 0 : 1; }
  return on == 1;
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}


// ---------------------------------------------------------------------------------------------
// Optional per-kernel-class timing with CUDA events on the launching stream (bench.py's roofline leg).
// Off by default; never enabled inside a timed region or a graph capture.
// ---------------------------------------------------------------------------------------------
struct ProfRec { cudaEvent_t a, b; int cls; double work; };
static bool g_prof_on = false;      // record per-launch events
static bool g_prof_serial = false;  // keep the backward on one stream (no side branch) without recording anything
static std::vector<ProfRec> g_prof;

// Events come from a pool created when profiling is switched on: creating two events per launch on the hot host path
// made the profiled pass host-bound, the GPU idled between launches and the idle time landed inside the intervals.
static std::vector<cudaEvent_t> g_pool;
static size_t g_pool_next = 0;
constexpr size_t PROF_POOL = 8192;

bool prof_enabled() { return g_prof_on; }
bool prof_serial() { return g_prof_on || g_prof_serial; }
int prof_open(int cls, double work, cudaStream_t st) {
  if (g_pool_next + 2 > g_pool.size()) return -1;   // pool exhausted: stop recording rather than stall the host
  ProfRec r;
  r.cls = cls; r.work = work;
  r.a = g_pool[g_pool_next++];
  r.b = g_pool[g_pool_next++];
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
  return (int)g_prof.size() - 1;
}
void prof_close(int id, cudaStream_t st) {
  if (id >= 0 && id < (int)g_prof.size()) cudaEventRecord(g_prof[id].b, st);
}
void prof_enable(int mode) {   // 0 off, 1 per-launch events (serial), 2 serial only
  const bool on = mode == 1;
  g_prof_serial = mode == 2;
  g_prof.clear();
  g_pool_next = 0;
  if (on && g_pool.empty()) {
    g_pool.resize(PROF_POOL);
    for (auto& e : g_pool) {
      if (cudaEventCreate(&e) != cudaSuccess) { g_pool.clear(); break; }
    }
  }
  g_prof_on = on && !g_pool.empty();
}
int prof_read(int cls, double* ms, double* work, long long* count) {
  double tm = 0, w = 0; long long n = 0;
  for (auto& r : g_prof) {
    if (r.cls != cls) continue;
    if (cudaEventSynchronize(r.b) != cudaSuccess) return 1;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) return 1;
    tm += t; w += r.work; ++n;
  }
  *ms = tm; *work = w; *count = n;
  return 0;
}

}  // namespace fervit
