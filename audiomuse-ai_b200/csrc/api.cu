// Library lifecycle, error reporting, launch accounting (C ABI: include/audiomuse_b200.h).
#include "common.cuh"

#include <mutex>

namespace am {

static thread_local std::string t_error;
std::atomic<uint64_t> g_launches{0};
static std::mutex g_init_mu;
static std::atomic<int> g_inited{0};
static int g_sms = 0, g_cc = 0;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  t_error = buf;
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  const bool oom = (e == cudaErrorMemoryAllocation);
  // "out of memory" keeps tasks/memory_utils.py's OOM detection (string match) working
  set_error("%s%s failed: %s (%s:%d)", oom ? "out of memory: " : "", what, cudaGetErrorString(e), file,
            line);
  cudaGetLastError();  // clear the sticky-free error state
  return oom ? AM_ERR_OOM : AM_ERR_CUDA;
}

int ensure_init() {
  if (g_inited.load(std::memory_order_acquire)) return AM_OK;
  return am_init(-1);
}
int sm_count() { return g_sms; }
int device_cc() { return g_cc; }

}  // namespace am

using namespace am;

extern "C" int am_init(int device_ordinal) {
  std::lock_guard<std::mutex> lk(g_init_mu);
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device visible (%s); libaudiomuse_b200 has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    cudaGetLastError();
    return AM_ERR_NO_DEVICE;
  }
  if (device_ordinal >= 0) {
    AM_CHECK(device_ordinal < n, "am_init: device %d out of range (%d visible)", device_ordinal, n);
    AM_CUDA(cudaSetDevice(device_ordinal));
  }
  int dev = 0;
  AM_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  AM_CUDA(cudaGetDeviceProperties(&p, dev));
  if (p.major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, p.major, p.minor);
    return AM_ERR_NO_DEVICE;
  }
  g_sms = p.multiProcessorCount;
  g_cc = p.major * 10 + p.minor;
  AM_CUDA(cudaFree(0));
  g_inited.store(1, std::memory_order_release);
  return AM_OK;
}

extern "C" void am_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (g_inited.load()) cudaDeviceSynchronize();
  g_inited.store(0);
}

extern "C" const char* am_last_error(void) { return t_error.c_str(); }
extern "C" int am_version(void) { return 100; }
extern "C" uint64_t am_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
