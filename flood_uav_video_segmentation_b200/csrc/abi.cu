// abi.cu — process-wide plumbing behind the C ABI (include/fuvs.h): error
// string, launch counter, device capability check.  No kernels here.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "fuvs_common.cuh"

namespace fuvs {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// After a <<<>>> launch.  The error is reported but NOT cleared: if it was already pending when the entry was called (a
// failed call of another library on this thread, e.g. PyTorch) its owner must still see it; a launch failure of our own
// is equally visible to the caller's next CUDA check.  (Launches through cudaLaunchKernelEx check its return value.)
int check_launch(const char* what) {
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess)
    return set_error(FUVS_ECUDA, "%s: %s (pending CUDA error, not cleared by libfuvs)", what, cudaGetErrorString(e));
  count_launch();
  return FUVS_OK;
}

struct DevInfo {
  int ok = -100;   // not probed
  int sms = 0;
};
static DevInfo g_dev[64];

static DevInfo* probe() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    cudaGetLastError();
    return nullptr;
  }
  DevInfo* d = &g_dev[dev];
  if (d->ok == -100) {
    int major = 0, sms = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    d->sms = sms;
    d->ok = (major == 10) ? FUVS_OK : FUVS_ENODEV;
  }
  return d;
}

int device_ok() {
  DevInfo* d = probe();
  if (!d) return set_error(FUVS_ENODEV, "no CUDA device is current (libfuvs has no CPU path)");
  if (d->ok != FUVS_OK)
    return set_error(FUVS_ENODEV, "libfuvs is built for sm_100a only; current device is not compute capability 10.x");
  return FUVS_OK;
}

int sm_count() {
  DevInfo* d = probe();
  return (d && d->sms > 0) ? d->sms : 148;
}

}  // namespace fuvs

extern "C" {

int fuvs_abi_version(void) { return FUVS_ABI_VERSION; }
const char* fuvs_last_error(void) { return fuvs::g_err; }
long long fuvs_launch_count(void) { return fuvs::g_launches.load(std::memory_order_relaxed); }
int fuvs_device_ok(void) { return fuvs::device_ok(); }

}  // extern "C"
