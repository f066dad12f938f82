// Shared host-side helpers for libnbdt_b200 (error reporting across the C ABI, device queries).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdint>

#include "../../include/nbdt_b200.h"

namespace nbdt {

// thread-local message returned by nbdt_last_error(); never throws across the ABI
char* error_buffer();
int set_error(int code, const char* fmt, ...);

#define NBDT_CUDA_OK(expr)                                                                    \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return ::nbdt::set_error(NBDT_ECUDA, "%s failed: %s (%s:%d)", #expr,                    \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                   \
  } while (0)

#define NBDT_REQUIRE(cond, ...)                                      \
  do {                                                               \
    if (!(cond)) return ::nbdt::set_error(NBDT_EINVAL, __VA_ARGS__); \
  } while (0)

struct DeviceInfo {
  int sm_count;
  int max_smem_optin;
};
int device_info(int device, DeviceInfo* out);

}  // namespace nbdt
