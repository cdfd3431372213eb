// Shared host-side helpers of the lxg C-ABI implementation.
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace lxg {
// Records `msg` as the calling thread's last error and returns `code`.
int set_error(int code, const std::string& msg);
bool is_device_ptr(const void* p);
int num_sms();
}  // namespace lxg

#define LXG_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t lxg_e_ = (call);                                                                \
    if (lxg_e_ != cudaSuccess) {                                                                \
      cudaGetLastError();                                                                       \
      return ::lxg::set_error(LXG_ECUDA, std::string(#call) + ": " + cudaGetErrorString(lxg_e_)); \
    }                                                                                           \
  } while (0)
