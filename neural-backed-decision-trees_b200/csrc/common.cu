#include "common.h"

#include <cstring>

namespace nbdt {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int device_info(int device, DeviceInfo* out) {
  static DeviceInfo cache[64];
  static bool have[64] = {false};
  if (device < 0 || device >= 64) return set_error(NBDT_EINVAL, "bad device index %d", device);
  if (!have[device]) {
    NBDT_CUDA_OK(cudaDeviceGetAttribute(&cache[device].sm_count, cudaDevAttrMultiProcessorCount, device));
    NBDT_CUDA_OK(cudaDeviceGetAttribute(&cache[device].max_smem_optin,
                                        cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    have[device] = true;
  }
  *out = cache[device];
  return NBDT_OK;
}

}  // namespace nbdt

extern "C" {

int nbdt_version(void) { return NBDT_B200_VERSION; }

const char* nbdt_last_error(void) { return nbdt::error_buffer(); }

int nbdt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

}  // extern "C"
