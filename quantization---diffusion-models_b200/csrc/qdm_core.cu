// libqdm.so core: thread-local error text, launch accounting, device gate.
#include "qdm_common.cuh"
#include <mutex>

namespace {
thread_local char g_err[768] = "";
thread_local int64_t g_launches = 0;
std::mutex g_dev_mu;
int g_dev_ok[64];  // 0 = unknown, 1 = cc 10.0, -1 = anything else
}  // namespace

void qdm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void qdm_count_launch(int n) { g_launches += n; }

// Every compute entry calls this first: the library only carries sm_100a SASS and has no
// other path, so anything that is not a cc-10.0 device is refused up front.
int qdm_require_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    qdm_set_error("no CUDA device: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return QDM_ERR_DEVICE;
  }
  if (dev >= 0 && dev < 64 && g_dev_ok[dev] == 1) return QDM_OK;
  return qdm_device_check(dev);
}

extern "C" int qdm_version(void) { return 100; }

extern "C" const char* qdm_last_error(void) { return g_err; }

extern "C" int qdm_device_check(int device) {
  int major = 0, minor = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (e != cudaSuccess) {
    qdm_set_error("qdm_device_check(%d): %s", device, cudaGetErrorString(e));
    cudaGetLastError();
    return QDM_ERR_DEVICE;
  }
  const bool ok = (major == 10 && minor == 0);
  if (device >= 0 && device < 64) {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    g_dev_ok[device] = ok ? 1 : -1;
  }
  if (!ok) {
    qdm_set_error("device %d has compute capability %d.%d; libqdm is built for sm_100a (B200) only",
                  device, major, minor);
    return QDM_ERR_DEVICE;
  }
  return QDM_OK;
}

extern "C" int64_t qdm_launch_count(int reset) {
  int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}
