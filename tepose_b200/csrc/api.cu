// Library-level entry points: version, error text, device query.
#include "common.cuh"
#include <string.h>
#include <atomic>

namespace tp {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launches() { return g_launches.load(std::memory_order_relaxed); }

static long long* g_trace = nullptr;
long long* trace_ptr() { return g_trace; }
void set_trace_ptr(long long* p) { g_trace = p; }

static std::atomic<int> g_pdl{1};
bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; }
void set_pdl(int on) { g_pdl.store(on ? 1 : 0, std::memory_order_relaxed); }

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
}  // namespace tp

extern "C" int tp_version(void) { return 100; }

namespace tp { void set_pdl(int on); }
extern "C" int tp_set_pdl(int enable) { const int was = tp::pdl_enabled() ? 1 : 0; tp::set_pdl(enable); return was; }

namespace tp { unsigned long long launches(); }
extern "C" unsigned long long tp_launch_count(void) { return tp::launches(); }

extern "C" const char* tp_last_error(void) { return tp::g_err; }

extern "C" int tp_device_info(int device, int* sm_count, int* cc_major, int* cc_minor,
                              size_t* smem_per_block_optin) {
  cudaDeviceProp p;
  TP_CUDA(cudaGetDeviceProperties(&p, device));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (smem_per_block_optin) *smem_per_block_optin = p.sharedMemPerBlockOptin;
  return TP_OK;
}
