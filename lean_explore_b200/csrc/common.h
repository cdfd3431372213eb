// Shared host-side helpers of the lxg C-ABI implementation.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <string>

namespace lxg {
// Records `msg` as the calling thread's last error and returns `code`.
int set_error(int code, const std::string& msg);
bool is_device_ptr(const void* p);
int num_sms();
// cuTensorMapEncodeTiled resolved through the runtime by lxg_init (no link-time libcuda dependency).
bool encode_tensor_map_ready();
CUresult encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dt, cuuint32_t rank, void* base,
                           const cuuint64_t* gdim, const cuuint64_t* gstride, const cuuint32_t* box,
                           const cuuint32_t* estr, CUtensorMapInterleave il, CUtensorMapSwizzle sw,
                           CUtensorMapL2promotion l2, CUtensorMapFloatOOBfill oob);
}  // namespace lxg

// Kernel launch with optional programmatic dependent launch (see ptx::pdl_wait).
template <typename... KArgs, typename... Args>
inline cudaError_t lxg_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define LXG_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t lxg_e_ = (call);                                                                \
    if (lxg_e_ != cudaSuccess) {                                                                \
      cudaGetLastError();                                                                       \
      return ::lxg::set_error(LXG_ECUDA, std::string(#call) + ": " + cudaGetErrorString(lxg_e_)); \
    }                                                                                           \
  } while (0)
