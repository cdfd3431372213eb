// Shared host-side helpers of the lxg C-ABI implementation.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <nvtx3/nvToolsExt.h>

#include <string>

namespace lxg {
// NVTX range around an entry point / phase (header-only NVTX 3: a no-op unless a profiler is attached).
// nsys / ncu --nvtx show the request path as lxg_search { prep | scan | merge | exact }, lxg_encode,
// lxg_decoder_embed / lxg_decoder_rerank, lxg_merge_topk.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
// Records `msg` as the calling thread's last error and returns `code`.
int set_error(int code, const std::string& msg);
bool is_device_ptr(const void* p);
// SM count of the calling thread's current device (filled by lxg_init for every device it brought up).
int num_sms();
// Ordinal of the device that owns device pointer `p`, or -1 (host / unknown memory).
int device_of_ptr(const void* p);
// Raises the dynamic shared-memory limit of kernel `func` ON THE CURRENT DEVICE to at least `bytes`.
// The attribute is per device and the cache behind this is per (device, kernel) and thread safe -
// handles on different GPUs of one process and concurrent host threads do not trip over each other.
cudaError_t ensure_dyn_smem(const void* func, size_t bytes);

// Every entry point that takes a handle makes the handle's device current for the duration of the
// call (worker threads of an executor start on device 0) and restores the caller's device on exit.
struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  explicit DeviceGuard(int device) {
    if (device >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != device) {
      if (cudaSetDevice(device) == cudaSuccess) changed = true;
    }
  }
  ~DeviceGuard() {
    if (changed) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
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
