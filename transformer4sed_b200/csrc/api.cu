// Library-level entry points: version, error string, device check.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace t4s {
long long g_launches = 0;
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return T4S_ERR_CUDA;
}

void ensure_context() {
  static thread_local bool done = false;
  if (!done) {
    cudaFree(nullptr);
    done = true;
  }
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached > 0 ? cached : 148;
}
}  // namespace t4s

extern "C" {

int t4s_version(void) { return 100; }

const char* t4s_last_error(void) { return t4s::g_err; }

int t4s_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  T4S_CUDA(cudaGetDevice(&dev));
  T4S_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  T4S_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    t4s::set_error("libt4s is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return T4S_ERR_DEVICE;
  }
  return T4S_OK;
}

int t4s_sm_count(void) { return t4s::sm_count(); }

long long t4s_launch_count(void) { return t4s::g_launches; }

}  // extern "C"
