// Host-side helpers shared by the encoder / decoder entry points: tensor maps of the GEMM
// operands and the launch of gemm_tc_kernel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <string>

#include "../../include/lxg.h"
#include "common.h"
#include "encoder_kernels.cuh"

namespace lxg {

inline int make_map(CUtensorMap* m, const void* base, int rows, int cols, int box_rows = kGemmBM) {
  // row-major fp16 [rows, cols]; box = 64 columns x box_rows (128) rows, 128-byte swizzle
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * sizeof(__half)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kGemmBK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = lxg::encode_tensor_map(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                                      box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(LXG_ECUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(r));
  return LXG_OK;
}

// Row-major fp32 [rows, cols] as 32 x 32 boxes, 128-byte swizzle: the destination of the TMA reduce
// epilogue (gemm_pair_kernel<kEpiAccF32, BN, true>).
inline int make_map_f32_acc(CUtensorMap* m, void* base, int rows, int cols) {
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * sizeof(float)};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = lxg::encode_tensor_map(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(LXG_ECUDA, "cuTensorMapEncodeTiled (fp32 accumulate map) failed with CUresult " + std::to_string(r));
  return LXG_OK;
}

// LXG_GEMM_TMA_ACC=0: per-thread read-modify-write epilogue for the fp32 accumulate GEMMs (A/B measurements)
inline bool gemm_tma_acc_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("LXG_GEMM_TMA_ACC");
    return !(e && e[0] == '0');
  }();
  return on;
}

// LXG_GEMM_SINGLE=1 keeps every GEMM on the single-CTA kernel (A/B measurements)
inline bool gemm_pairs_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("LXG_GEMM_SINGLE");
    return !(e && e[0] == '1');
  }();
  return on;
}

// LXG_GEMM_NARROW=1 runs N = 1024-class projections whose 256 x 256 tiles would not even give every cluster
// one tile on 256 x 128 tiles (w64 maps).  Off by default since the fp32 accumulate epilogue went to the TMA
// engine: the narrow tiles bought overlap of a slow epilogue with the next main loop (29.9 -> 29.4 us per GEMM);
// with the fast epilogue one 256 x 256 tile per cluster wins (rerank 16 x 256: 3.57 vs 3.77 ms).
inline bool gemm_narrow_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("LXG_GEMM_NARROW");
    return e && e[0] == '1';
  }();
  return on;
}

// w64: optional second map of W with 64-row boxes (make_map(..., 64)) for the narrow tiles above.
// out_map: optional fp32 map of the output (make_map_f32_acc) IN DEVICE MEMORY: kEpiAccF32 then accumulates
// through TMA reduces.
template <int EPI>
inline cudaError_t launch_gemm(const CUtensorMap& a, const CUtensorMap& w, const GemmParams& gp, cudaStream_t st, bool pdl = false,
                               const CUtensorMap* w64 = nullptr, const CUtensorMap* out_map = nullptr) {
  // more than one row tile and N a multiple of 256: 256 x 256 tiles on CTA pairs
  if (gp.m > kGemmBM && gp.n % kPairBN == 0 && gemm_pairs_enabled()) {
    const int max_clusters = std::max(1, lxg::num_sms() / 2);
    const int tiles = ((gp.m + 2 * kGemmBM - 1) / (2 * kGemmBM)) * (gp.n / kPairBN);
    const bool narrow = EPI != kEpiSwiGLU && w64 != nullptr && tiles <= max_clusters && 2 * tiles > max_clusters && gemm_narrow_enabled();
    if constexpr (EPI == kEpiAccF32) {
      if (out_map != nullptr && gemm_tma_acc_enabled()) {
        if (narrow) {
          cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(&gemm_pair_kernel<EPI, 128, true>), kPairSmemAcc);
          if (e != cudaSuccess) return e;
          return lxg_launch(gemm_pair_kernel<EPI, 128, true>, dim3(2 * std::min(2 * tiles, max_clusters)), dim3(kPairThreads), kPairSmemAcc,
                            st, pdl, a, *w64, out_map, gp);
        }
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(&gemm_pair_kernel<EPI, kPairBN, true>), kPairSmemAcc);
        if (e != cudaSuccess) return e;
        return lxg_launch(gemm_pair_kernel<EPI, kPairBN, true>, dim3(2 * std::min(tiles, max_clusters)), dim3(kPairThreads), kPairSmemAcc, st,
                          pdl, a, w, out_map, gp);
      }
    }
    if constexpr (EPI != kEpiSwiGLU) {
      if (narrow) {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(&gemm_pair_kernel<EPI, 128>), kPairSmem);
        if (e != cudaSuccess) return e;
        return lxg_launch(gemm_pair_kernel<EPI, 128>, dim3(2 * std::min(2 * tiles, max_clusters)), dim3(kPairThreads), kPairSmem, st,
                          pdl, a, *w64, static_cast<const CUtensorMap*>(nullptr), gp);
      }
    }
    {
      cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(&gemm_pair_kernel<EPI>), kPairSmem);
      if (e != cudaSuccess) return e;
    }
    const int clusters = std::min(tiles, max_clusters);
    return lxg_launch(gemm_pair_kernel<EPI>, dim3(2 * clusters), dim3(kPairThreads), kPairSmem, st, pdl, a, w,
                      static_cast<const CUtensorMap*>(nullptr), gp);
  }
  {
    cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(&gemm_tc_kernel<EPI>), kGemmSmem);
    if (e != cudaSuccess) return e;
  }
  const int tiles = ((gp.m + kGemmBM - 1) / kGemmBM) * (gp.n / kGemmBN) * std::max(1, gp.ksplit);
  return lxg_launch(gemm_tc_kernel<EPI>, dim3(std::min(tiles, std::max(1, lxg::num_sms()))), dim3(kGemmThreads), kGemmSmem, st, pdl, a, w, gp);
}

// Query path (at most 32 rows): narrow output tiles, A staged as one 32-row box per k-block.
// a32 = map of A with 32-row boxes, wbn = map of W with BN-row boxes.
template <int EPI, int BN>
inline cudaError_t launch_gemm_query(const CUtensorMap& a32, const CUtensorMap& wbn, const GemmParams& gp, cudaStream_t st, bool pdl) {
  {
    cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(&gemm_tc_kernel<EPI, BN, 32>), kGemmSmem);
    if (e != cudaSuccess) return e;
  }
  const int tiles = (gp.n / BN) * std::max(1, gp.ksplit);
  return lxg_launch(gemm_tc_kernel<EPI, BN, 32>, dim3(std::min(tiles, std::max(1, lxg::num_sms()))), dim3(kGemmThreads), kGemmSmem, st, pdl,
                    a32, wbn, gp);
}

}  // namespace lxg
