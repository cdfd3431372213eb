// Kernels of the BERT-class sentence encoder (replaces SentenceTransformer.encode inside
// EmbeddingClient.embed, reference src/lean_explore/util/embedding_client.py:88-101):
//   embeddings + LayerNorm -> L x [ QKV GEMM, attention, out-proj GEMM (+residual), LayerNorm,
//   FFN GEMM + GELU, FFN GEMM (+residual), LayerNorm ] -> pooling (mean / CLS) -> L2 normalise.
// Post-LN BERT exactly as transformers.BertModel computes it (erf GELU, learned absolute
// positions, token type 0, additive -inf key mask).  Activations are fp16 between kernels,
// every accumulation / LayerNorm / softmax is fp32.
//
// The GEMMs run on tcgen05 tensor cores: TMA (128B swizzle) stages A[128 x 64] and W[BN x 64]
// tiles into a shared-memory ring, one elected thread issues tcgen05.mma (M=128, N=BN, K=16) into
// a TMEM accumulator, four epilogue warps read it back with tcgen05.ld and apply
// bias / GELU / residual in registers.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include "ptx.cuh"

namespace lxg {

// ------------------------------------------------------------------ embeddings + LayerNorm
// One warp per token.  out = LN(word[id] + pos[s] + type[0]) * g + b   (fp16)
static __global__ void __launch_bounds__(256)
embed_ln_kernel(const int* __restrict__ ids, int tokens, int seq, int hidden, int vocab,
                const __half* __restrict__ word, const __half* __restrict__ pos,
                const __half* __restrict__ type0, const float* __restrict__ g,
                const float* __restrict__ b, float eps, __half* __restrict__ out) {
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= tokens) return;
  int id = ids[t];
  id = min(max(id, 0), vocab - 1);
  const int s = t % seq;
  const __half2* w2 = reinterpret_cast<const __half2*>(word + static_cast<size_t>(id) * hidden);
  const __half2* p2 = reinterpret_cast<const __half2*>(pos + static_cast<size_t>(s) * hidden);
  const __half2* t2 = reinterpret_cast<const __half2*>(type0);
  constexpr int kMax = 16;  // hidden <= 1024
  float2 v[kMax];
  float sum = 0.f;
  const int n2 = hidden >> 1;
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int j = i * 32 + lane;
    v[i] = make_float2(0.f, 0.f);
    if (j < n2) {
      const float2 a = __half22float2(w2[j]), c = __half22float2(p2[j]), e = __half22float2(t2[j]);
      v[i] = make_float2(a.x + c.x + e.x, a.y + c.y + e.y);
      sum += v[i].x + v[i].y;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / hidden;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int j = i * 32 + lane;
    if (j < n2) {
      const float dx = v[i].x - mean, dy = v[i].y - mean;
      var += dx * dx + dy * dy;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / hidden + eps);
  __half2* o2 = reinterpret_cast<__half2*>(out + static_cast<size_t>(t) * hidden);
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int j = i * 32 + lane;
    if (j < n2) {
      const float2 gg = reinterpret_cast<const float2*>(g)[j], bb = reinterpret_cast<const float2*>(b)[j];
      o2[j] = __floats2half2_rn((v[i].x - mean) * rstd * gg.x + bb.x, (v[i].y - mean) * rstd * gg.y + bb.y);
    }
  }
}

// ------------------------------------------------------------------ LayerNorm (fp32 in, fp16 out)
static __global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int tokens, int hidden, const float* __restrict__ g,
                 const float* __restrict__ b, float eps, __half* __restrict__ out) {
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= tokens) return;
  const float2* x2 = reinterpret_cast<const float2*>(x + static_cast<size_t>(t) * hidden);
  constexpr int kMax = 16;
  float2 v[kMax];
  float sum = 0.f;
  const int n2 = hidden >> 1;
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int j = i * 32 + lane;
    v[i] = make_float2(0.f, 0.f);
    if (j < n2) {
      v[i] = x2[j];
      sum += v[i].x + v[i].y;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / hidden;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int j = i * 32 + lane;
    if (j < n2) {
      const float dx = v[i].x - mean, dy = v[i].y - mean;
      var += dx * dx + dy * dy;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / hidden + eps);
  __half2* o2 = reinterpret_cast<__half2*>(out + static_cast<size_t>(t) * hidden);
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int j = i * 32 + lane;
    if (j < n2) {
      const float2 gg = reinterpret_cast<const float2*>(g)[j], bb = reinterpret_cast<const float2*>(b)[j];
      o2[j] = __floats2half2_rn((v[i].x - mean) * rstd * gg.x + bb.x, (v[i].y - mean) * rstd * gg.y + bb.y);
    }
  }
}

__device__ __forceinline__ uint32_t pack_half2(float x, float y) {
  const __half2 h = __floats2half2_rn(x, y);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ------------------------------------------------------------------ tcgen05 GEMM
// C[M, N] = epi(A[M, K] . W[N, K]^T + bias[N])
//   kEpiStore : fp16 out                     (QKV projection)
//   kEpiGelu  : erf-GELU, fp16 out           (intermediate.dense)
//   kEpiResid : + residual fp16 [M, N], fp32 out   (attention.output.dense / output.dense, pre-LN)
//   kEpiAccF32: out fp32 [M, N] += result    (decoder o_proj / down_proj onto the fp32 residual stream)
//   kEpiSwiGLU: silu(gate) * up, fp16 out [M, N/2]; W rows interleaved per 128-row tile as
//               gate[32] | up[32] | gate[32] | up[32] so one thread holds both halves of a column
// bias may be NULL (decoder projections have none).
//   kEpiPartial: fp32 out [ksplit][M, N] - split-K partial sums of a skinny GEMM (M <= 128): the
//               consumer (RMSNorm / head kernel) adds the slabs to the residual stream in a fixed order
enum GemmEpilogue { kEpiStore = 0, kEpiGelu = 1, kEpiResid = 2, kEpiAccF32 = 3, kEpiSwiGLU = 4, kEpiPartial = 5 };

constexpr int kGemmBM = 128;
constexpr int kGemmBN = 128;
constexpr int kGemmBK = 64;
constexpr int kGemmStages = 6;
constexpr int kGemmStageBytes = (kGemmBM + kGemmBN) * kGemmBK * 2;  // 32 KB
constexpr int kGemmEpiWarps = 8;                                    // 2 per TMEM lane quarter (column halves)
constexpr int kGemmThreads = (kGemmEpiWarps + 2) * 32;              // + TMA warp + MMA warp
constexpr int kGemmSmem = kGemmStages * kGemmStageBytes + 1024;

struct GemmParams {
  const float* bias;       // [N]
  const __half* residual;  // [M, N] (kEpiResid)
  void* out;               // fp16 [M, N] or fp32 [M, N]
  int m, n, k;
  int ksplit;              // single-CTA kernel only: K is cut into ksplit ranges, one tile each (0 / 1 = off)
  size_t split_stride;     // kEpiPartial: elements between the partial slabs
  unsigned long long* trace;  // optional device timeline of gemm_pair_kernel: [CTA][16] %globaltimer stamps (LXG_GEMM_TRACE), else NULL
};

__device__ __forceinline__ void gemm_trace(const GemmParams& p, int slot) {
  if (p.trace != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[static_cast<size_t>(blockIdx.x) * 16 + slot] = t;
  }
}

// Epilogue of one thread (= one output row) over `nchunks` 32-column chunks of an accumulator,
// starting at chunk c0: taddr = TMEM address of the accumulator (lane field set), n0 = first
// output column of the tile.  Shared by the single-CTA and the CTA-pair kernels.
template <int EPI>
__device__ __forceinline__ void gemm_epilogue_row(uint32_t taddr, int c0, int nchunks, int row, int n0, const GemmParams& p) {
  if constexpr (EPI == kEpiSwiGLU) {
#pragma unroll 1
    for (int c = c0; c < c0 + nchunks; c += 2) {  // chunk c = gate, chunk c + 1 = up of the same 32 outputs
      uint32_t rg[32], ru[32];
      ptx::tmem_ld_32x32b_x32(taddr + c * 32, rg);
      ptx::tmem_ld_32x32b_x32(taddr + (c + 1) * 32, ru);
      ptx::tc_wait_ld();
      if (row < p.m) {
        const int ocol = (n0 >> 1) + (c >> 1) * 32;
        uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + static_cast<size_t>(row) * (p.n >> 1) + ocol);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float g = __uint_as_float(rg[j]), u = __uint_as_float(ru[j]);
          v[j] = fminf(fmaxf(__fdividef(g, 1.0f + __expf(-g)) * u, -65504.f), 65504.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          o.x = pack_half2(v[8 * j], v[8 * j + 1]);
          o.y = pack_half2(v[8 * j + 2], v[8 * j + 3]);
          o.z = pack_half2(v[8 * j + 4], v[8 * j + 5]);
          o.w = pack_half2(v[8 * j + 6], v[8 * j + 7]);
          o4[j] = o;
        }
      }
    }
  } else {
#pragma unroll 1
    for (int c = c0; c < c0 + nchunks; ++c) {
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
      ptx::tc_wait_ld();
      const int col0 = n0 + c * 32;
      if (row < p.m && col0 < p.n) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias != nullptr) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = __ldg(b4 + j);
            v[4 * j + 0] += bb.x;
            v[4 * j + 1] += bb.y;
            v[4 * j + 2] += bb.z;
            v[4 * j + 3] += bb.w;
          }
        }
        if constexpr (EPI == kEpiGelu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.5f * v[j] * (1.0f + erff(v[j] * 0.70710678118654752f));
        }
        if constexpr (EPI == kEpiResid) {
          const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + static_cast<size_t>(row) * p.n + col0);
          float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.n + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 rr = __ldg(r4 + j);
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&rr.x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&rr.y));
            const float2 cc = __half22float2(*reinterpret_cast<const __half2*>(&rr.z));
            const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&rr.w));
            o4[2 * j] = make_float4(v[8 * j] + a.x, v[8 * j + 1] + a.y, v[8 * j + 2] + b.x, v[8 * j + 3] + b.y);
            o4[2 * j + 1] = make_float4(v[8 * j + 4] + cc.x, v[8 * j + 5] + cc.y, v[8 * j + 6] + d.x, v[8 * j + 7] + d.y);
          }
        } else if constexpr (EPI == kEpiPartial) {
          float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.n + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else if constexpr (EPI == kEpiAccF32) {
          float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.n + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = o4[j];
            o.x += v[4 * j + 0];
            o.y += v[4 * j + 1];
            o.z += v[4 * j + 2];
            o.w += v[4 * j + 3];
            o4[j] = o;
          }
        } else {
          uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + static_cast<size_t>(row) * p.n + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_half2(v[8 * j], v[8 * j + 1]);
            o.y = pack_half2(v[8 * j + 2], v[8 * j + 3]);
            o.z = pack_half2(v[8 * j + 4], v[8 * j + 5]);
            o.w = pack_half2(v[8 * j + 6], v[8 * j + 7]);
            o4[j] = o;
          }
        }
      }
    }
  }
}

// Persistent, warp-specialised: grid = min(#tiles, #SMs), every CTA walks tiles
// t = blockIdx.x, blockIdx.x + gridDim.x, ... (m fastest, so concurrently running CTAs share a W
// tile in L2).  Warp 8 streams A / W k-blocks through a 6-stage TMA ring, warp 9 issues
// tcgen05.mma (SS, M=128 N=128 K=16) into one of TWO TMEM accumulators, warps 0-7 drain the other
// one (thread = output row, warp >> 2 = column half): the epilogue of tile i (bias, erf-GELU,
// residual, stores) overlaps the main loop of tile i+1.
// BN / AROWS: the query path (a few dozen rows: one row tile, weights streamed once) runs narrower
// output tiles - N = 32 or 64 columns, so that 100+ CTAs share the weight stream instead of N / 128 -
// and stages only the AROWS = 32 rows of A that exist (the MMA still reads 128 rows of shared
// memory; what it makes of the unwritten ones lands in accumulator rows nobody stores).
template <int EPI, int BN = kGemmBN, int AROWS = kGemmBM>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
               const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kGemmStages];
  __shared__ __align__(8) uint64_t empty_bar[kGemmStages];
  __shared__ __align__(8) uint64_t acc_full_bar[2];
  __shared__ __align__(8) uint64_t acc_empty_bar[2];
  __shared__ uint32_t tmem_base_holder;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (p.m + kGemmBM - 1) / kGemmBM;
  static_assert(BN == 128 || BN == 64 || (BN == 32 && EPI != kEpiSwiGLU), "tile widths: 128, 64, or 32 for the plain epilogues");
  constexpr int kStageTx = (AROWS + BN) * kGemmBK * 2;  // bytes TMA delivers per stage
  const int tiles_mn = tiles_m * (p.n / BN);
  const int ksplit = p.ksplit > 1 ? p.ksplit : 1;
  const int tiles = tiles_mn * ksplit;  // tile t: k range t / tiles_mn, output tile t % tiles_mn
  const int num_kb = (p.k + kGemmBK - 1) / kGemmBK;
  const uint32_t ring_u32 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t kIdesc = ptx::make_idesc_f16(kGemmBM, BN);
  constexpr int kTmaWarp = kGemmEpiWarps, kMmaWarp = kGemmEpiWarps + 1;

  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < kGemmStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&acc_full_bar[a], 1);
      ptx::mbar_init(&acc_empty_bar[a], kGemmEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == kTmaWarp) {
    if (lane == 0) {
      ptx::prefetch_tensormap(&tmap_a);
      ptx::prefetch_tensormap(&tmap_w);
    }
    ptx::tmem_alloc(&tmem_base_holder, 2 * BN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;
  // Everything above overlapped the previous kernel's tail (programmatic dependent launch).  So do the
  // first kGemmStages WEIGHT tiles: W does not depend on the previous kernel, so the TMA warp arms the
  // ring and issues those loads before it waits for the dependency - on the query path (a few k-blocks
  // per CTA) that is the CTA's whole weight stream (Qwen3 query embedding 0.975 -> 0.918 ms).  Every
  // role waits before it touches activations.

  if (warp == kTmaWarp) {
    const uint32_t full0 = ptx::opaque(ptx::smem_u32(&full_bar[0]));
    const uint32_t empty0 = ptx::opaque(ptx::smem_u32(&empty_bar[0]));
    const uint32_t ring0 = ptx::opaque(ring_u32);
    int pre = 0;  // ring slots whose W tile is already in flight
    for (int t = blockIdx.x; t < tiles && pre < kGemmStages; t += gridDim.x) {
      const int ks = t / tiles_mn, tt = t % tiles_mn;
      const int n0 = (tt / tiles_m) * BN;
      const int kb0 = ks * num_kb / ksplit, kb1 = (ks + 1) * num_kb / ksplit;
      for (int kb = kb0; kb < kb1 && pre < kGemmStages; ++kb, ++pre) {
        if (ptx::elect_one()) {
          const uint32_t fb = full0 + pre * 8;
          ptx::mbar_arrive_expect_tx_a(fb, kStageTx);
          ptx::tma_load_2d_a(ring0 + pre * kGemmStageBytes + kGemmBM * kGemmBK * 2, &tmap_w, kb * kGemmBK, n0, fb, ptx::kEvictLast);
        }
        __syncwarp();
      }
    }
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    uint32_t stage = 0, phase = 0;
    int step = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int ks = t / tiles_mn, tt = t % tiles_mn;
      const int m0 = (tt % tiles_m) * kGemmBM, n0 = (tt / tiles_m) * BN;
      const int kb0 = ks * num_kb / ksplit, kb1 = (ks + 1) * num_kb / ksplit;
      for (int kb = kb0; kb < kb1; ++kb, ++step) {
        ptx::mbar_wait_a(empty0 + stage * 8, phase ^ 1u);
        if (ptx::elect_one()) {
          const uint32_t fb = full0 + stage * 8;
          const uint32_t dst = ring0 + stage * kGemmStageBytes;
          if (step >= pre) ptx::mbar_arrive_expect_tx_a(fb, kStageTx);
          ptx::tma_load_2d_a(dst, &tmap_a, kb * kGemmBK, m0, fb, ptx::kEvictNormal);
          if (step >= pre) ptx::tma_load_2d_a(dst + kGemmBM * kGemmBK * 2, &tmap_w, kb * kGemmBK, n0, fb, ptx::kEvictLast);
        }
        __syncwarp();
        if (++stage == kGemmStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == kMmaWarp) {
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    const uint32_t full0 = ptx::opaque(ptx::smem_u32(&full_bar[0]));
    const uint32_t empty0 = ptx::opaque(ptx::smem_u32(&empty_bar[0]));
    const uint32_t afull0 = ptx::opaque(ptx::smem_u32(&acc_full_bar[0]));
    const uint32_t aempty0 = ptx::opaque(ptx::smem_u32(&acc_empty_bar[0]));
    const uint32_t desc_lo0 = ptx::opaque(((ring_u32 & 0x3FFFFu) >> 4) | (1u << 16));
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    uint32_t stage = 0, phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      ptx::mbar_wait_a(aempty0 + acc * 8, ((it >> 1) & 1) ^ 1u);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      const int ks = t / tiles_mn;
      const int kb0 = ks * num_kb / ksplit, kb1 = (ks + 1) * num_kb / ksplit;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait_a(full0 + stage * 8, phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t lo = desc_lo0 + stage * (kGemmStageBytes >> 4);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t adesc = (static_cast<uint64_t>(kDescHi) << 32) | (lo + k4 * 2);
            const uint64_t bdesc = (static_cast<uint64_t>(kDescHi) << 32) | (lo + ((kGemmBM * kGemmBK * 2) >> 4) + k4 * 2);
            ptx::mma_f16_ss(d_tmem, adesc, bdesc, kIdesc, ((kb - kb0) | k4) != 0 ? 1u : 0u);
          }
          ptx::tc_commit_a(empty0 + stage * 8);
          if (kb == kb1 - 1) ptx::tc_commit_a(afull0 + acc * 8);
        }
        __syncwarp();
        if (++stage == kGemmStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    // epilogue: thread = output row of the tile, warp >> 2 = which half of the columns (a SwiGLU
    // tile of 64 and a tile of 32 are one thread's work: the other half only releases the accumulator)
    const int half = warp >> 2;
    constexpr bool kOneHalf = BN == 32 || (BN == 64 && EPI == kEpiSwiGLU);
    constexpr int kChunksPerHalf = kOneHalf ? BN / 32 : BN / 64;
    const int c0 = kOneHalf ? 0 : half * kChunksPerHalf, nc = (kOneHalf && half == 1) ? 0 : kChunksPerHalf;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t afull0 = ptx::opaque(ptx::smem_u32(&acc_full_bar[0]));
    int it = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      const int ks = t / tiles_mn, tt = t % tiles_mn;
      const int m0 = (tt % tiles_m) * kGemmBM, n0 = (tt / tiles_m) * BN;
      const int row = m0 + (warp & 3) * 32 + lane;
      ptx::mbar_wait_a(afull0 + acc * 8, (it >> 1) & 1);
      ptx::tc_fence_after();
      if constexpr (EPI == kEpiPartial) {
        GemmParams ps = p;  // this k range's slab
        ps.out = reinterpret_cast<float*>(p.out) + static_cast<size_t>(ks) * p.split_stride;
        gemm_epilogue_row<EPI>(tmem_base + lane_base + acc * BN, c0, nc, row, n0, ps);
      } else {
        gemm_epilogue_row<EPI>(tmem_base + lane_base + acc * BN, c0, nc, row, n0, p);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty_bar[acc]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kTmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------ tcgen05 GEMM, CTA pairs
// Same contract as gemm_tc_kernel for N % 256 == 0, on 256 x 256 output tiles computed by a
// cluster of two CTAs with ONE tcgen05.mma.cta_group::2 stream (M = 256, N = 256, K = 16) issued
// by the even CTA.  Each CTA stages its own 128 rows of A and 128 of the tile's 256 rows of W per
// k-block (32 KB per stage, six stages) and drains its own 128 accumulator lanes.  Why: an SS MMA
// of a 128 x 128 tile reads 8 KB of operands from shared memory per 64 tensor-core clocks while TMA
// writes the next 8 KB - twice what the 128 B/clk shared-memory port delivers, which is where the
// single-CTA kernel saturates (~50 % of the tensor peak).  The pair tile reads and writes 8 KB per
// 128 clocks per SM.
constexpr int kPairBN = 256;
constexpr int kPairRing = 196608;  // bytes of shared memory used as operand pipeline
constexpr int kPairSmem = kPairRing + 1024;
// BN = 128 (`gemm_pair_kernel<EPI, 128>`, W staged as 64-row boxes): for N = 1024 projections over a
// few thousand rows, where 256 x 256 tiles give every cluster at most ONE tile - no epilogue / main
// loop overlap at all - and 256 x 128 tiles give most clusters two.  The narrower tile pays ~1.5x
// the shared-memory traffic per flop (24 KB per 256 tensor clocks and SM, read and written).
constexpr int kPairEpiWarps = 16;  // 4 per TMEM lane quarter (64 columns each): the epilogue is latency bound, more warps hide it
constexpr int kPairThreads = (kPairEpiWarps + 2) * 32;
// kTmaAcc (kEpiAccF32 only): out += result goes through the TMA engine - every epilogue warp stages its
// 32 x 32 fp32 block in shared memory (128B-swizzled, 4 KB per warp) and issues ONE
// cp.reduce.async.bulk.tensor (.add, f32) for it.  The read-modify-write of the fp32 residual stream
// then happens in the L2, fully coalesced, and no epilogue thread ever waits for a global load: the
// per-thread version (8 float4 loads of a 4 KB-strided row, add, 8 stores) kept a tile's accumulator
// busy for 5.2-5.6 us of a 27 us GEMM (device timeline, LXG_GEMM_TRACE).  The staging blocks take 64 KB;
// the operand ring keeps 5 (BN = 256) / 6 (BN = 128) stages instead of 6 / 8.
constexpr int kPairAccStage = kPairEpiWarps * 4096;
constexpr int kPairSmemAcc = 5 * 32768 + kPairAccStage + 1024;

template <int EPI, int BN = kPairBN, bool kTmaAcc = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 const CUtensorMap* __restrict__ tmap_out, const GemmParams p) {
  // (tmap_out lives in device memory and travels as a pointer: a third 128-byte __grid_constant__ map
  // pushed the parameter block past what the launch path handles cheaply - every pair-GEMM launch of
  // the build that carried it took ~4.5 us longer, 0.5 ms per reranker forward)
  static_assert(BN == 256 || (BN == 128 && EPI != kEpiSwiGLU), "tile widths: 256, or 128 for the plain epilogues");
  static_assert(!kTmaAcc || EPI == kEpiAccF32, "the TMA reduce epilogue is the fp32 accumulate one");
  constexpr int kPairStageBytes = (kGemmBM + BN / 2) * kGemmBK * 2;  // 32 KB (24 KB) per CTA and stage
  constexpr int kPairStages = (kTmaAcc ? 5 * 32768 : kPairRing) / kPairStageBytes;  // 6 (8); with staging blocks 5 (6)
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kPairStages];
  __shared__ __align__(8) uint64_t empty_bar[kPairStages];
  __shared__ __align__(8) uint64_t acc_full_bar[2];
  __shared__ __align__(8) uint64_t acc_empty_bar[2];
  __shared__ uint32_t tmem_base_holder;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) gemm_trace(p, 0);  // slot 0: CTA start
  const uint32_t rank = ptx::cluster_ctarank();
  const int cluster = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int tiles_m = (p.m + 2 * kGemmBM - 1) / (2 * kGemmBM);
  const int tiles = tiles_m * (p.n / BN);
  const int num_kb = (p.k + kGemmBK - 1) / kGemmBK;
  const uint32_t ring_u32 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t acc_stage_u32 = ring_u32 + 5 * 32768;  // kTmaAcc: 16 x 4 KB behind the ring (1024-aligned)
  constexpr uint32_t kIdesc = ptx::make_idesc_f16(2 * kGemmBM, BN);
  constexpr int kTmaWarp = kPairEpiWarps, kMmaWarp = kPairEpiWarps + 1;

  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < kPairStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&acc_full_bar[a], 1);
      ptx::mbar_init(&acc_empty_bar[a], 2 * kPairEpiWarps);  // both CTAs' epilogue warps
    }
    ptx::fence_barrier_init();
  }
  if (warp == kTmaWarp) {
    if (lane == 0) {
      ptx::prefetch_tensormap(&tmap_a);
      ptx::prefetch_tensormap(&tmap_w);
      if constexpr (kTmaAcc) ptx::prefetch_tensormap(tmap_out);
    }
    ptx::tmem_alloc_pair(&tmem_base_holder, 2 * BN);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;
  if (threadIdx.x == 0) gemm_trace(p, 1);  // slot 1: prologue done (barriers, TMEM, cluster sync)

  if (warp == kTmaWarp) {
    // both CTAs load their halves; all bytes are counted on the even CTA's full barrier
    const uint32_t full0 = ptx::opaque(ptx::smem_u32(&full_bar[0]));
    const uint32_t empty0 = ptx::opaque(ptx::smem_u32(&empty_bar[0]));
    const uint32_t ring0 = ptx::opaque(ring_u32);
    // the first ring-full of WEIGHT tiles goes out before the dependency is awaited (see gemm_tc_kernel)
    int pre = 0;
    for (int t = cluster; t < tiles && pre < kPairStages; t += nclusters) {
      const int n0 = (t / tiles_m) * BN + static_cast<int>(rank) * (BN / 2);
      for (int kb = 0; kb < num_kb && pre < kPairStages; ++kb, ++pre) {
        if (ptx::elect_one()) {
          const uint32_t fb = full0 + pre * 8;
          if (rank == 0) ptx::mbar_arrive_expect_tx_a(fb, 2 * kPairStageBytes);
          ptx::tma_load_2d_pair_a(ring0 + pre * kPairStageBytes + kGemmBM * kGemmBK * 2, &tmap_w, kb * kGemmBK, n0, fb, ptx::kEvictLast);
        }
        __syncwarp();
      }
    }
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    if (lane == 0) gemm_trace(p, 2);  // slot 2: dependency resolved
    uint32_t stage = 0, phase = 0;
    int step = 0;
    for (int t = cluster; t < tiles; t += nclusters) {
      const int m0 = (t % tiles_m) * 2 * kGemmBM + static_cast<int>(rank) * kGemmBM;
      const int n0 = (t / tiles_m) * BN + static_cast<int>(rank) * (BN / 2);
      for (int kb = 0; kb < num_kb; ++kb, ++step) {
        ptx::mbar_wait_a(empty0 + stage * 8, phase ^ 1u);
        if (ptx::elect_one()) {
          const uint32_t fb = full0 + stage * 8;
          if (rank == 0 && step >= pre) ptx::mbar_arrive_expect_tx_a(fb, 2 * kPairStageBytes);
          const uint32_t dst = ring0 + stage * kPairStageBytes;
          ptx::tma_load_2d_pair_a(dst, &tmap_a, kb * kGemmBK, m0, fb, ptx::kEvictNormal);
          if (step >= pre) ptx::tma_load_2d_pair_a(dst + kGemmBM * kGemmBK * 2, &tmap_w, kb * kGemmBK, n0, fb, ptx::kEvictLast);
        }
        __syncwarp();
        if (++stage == kPairStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == kMmaWarp) {
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    if (rank == 0) {
      const uint32_t full0 = ptx::opaque(ptx::smem_u32(&full_bar[0]));
      const uint32_t empty0 = ptx::opaque(ptx::smem_u32(&empty_bar[0]));
      const uint32_t afull0 = ptx::opaque(ptx::smem_u32(&acc_full_bar[0]));
      const uint32_t aempty0 = ptx::opaque(ptx::smem_u32(&acc_empty_bar[0]));
      const uint32_t desc_lo0 = ptx::opaque(((ring_u32 & 0x3FFFFu) >> 4) | (1u << 16));
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int t = cluster; t < tiles; t += nclusters, ++it) {
        const uint32_t acc = it & 1;
        ptx::mbar_wait_a(aempty0 + acc * 8, ((it >> 1) & 1) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait_a(full0 + stage * 8, phase);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t lo = desc_lo0 + stage * (kPairStageBytes >> 4);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint64_t adesc = (static_cast<uint64_t>(kDescHi) << 32) | (lo + k4 * 2);
              const uint64_t bdesc = (static_cast<uint64_t>(kDescHi) << 32) | (lo + ((kGemmBM * kGemmBK * 2) >> 4) + k4 * 2);
              ptx::mma_f16_ss_pair(d_tmem, adesc, bdesc, kIdesc, (kb | k4) != 0 ? 1u : 0u);
            }
            ptx::tc_commit_pair_a(empty0 + stage * 8, 3);
            if (kb == num_kb - 1) ptx::tc_commit_pair_a(afull0 + acc * 8, 3);
          }
          __syncwarp();
          if (++stage == kPairStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    // epilogue: thread = output row of this CTA's half of the tile, warp >> 2 = which quarter of the columns
    const int quarter = warp >> 2;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t afull0 = ptx::opaque(ptx::smem_u32(&acc_full_bar[0]));
    int it = 0;
    for (int t = cluster; t < tiles; t += nclusters, ++it) {
      const uint32_t acc = it & 1;
      const int m0 = (t % tiles_m) * 2 * kGemmBM + static_cast<int>(rank) * kGemmBM, n0 = (t / tiles_m) * BN;
      const int row = m0 + (warp & 3) * 32 + lane;
      ptx::mbar_wait_a(afull0 + acc * 8, (it >> 1) & 1);
      ptx::tc_fence_after();
      if (threadIdx.x == 0 && it < 3) gemm_trace(p, 6 + it);  // slots 6-8: accumulator of tile 0 / 1 / 2 complete
      if constexpr (kTmaAcc) {
        const uint32_t stg = acc_stage_u32 + warp * 4096;
#pragma unroll 1
        for (int c = 0; c < BN / 128; ++c) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(tmem_base + lane_base + acc * BN + (quarter * (BN / 128) + c) * 32, r);
          ptx::tc_wait_ld();
          if (c == BN / 128 - 1) {  // the accumulator is in registers: hand it back before the stores
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(&acc_empty_bar[acc], 0);
          }
          // the previous reduce of this warp has read the staging block
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
          const bool live = row < p.m;
#pragma unroll
          for (int j = 0; j < 8; ++j) {  // row `lane` of the block, 16-byte unit j at j ^ (lane & 7) (SWIZZLE_128B)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 128 + ((j ^ (lane & 7)) << 4)),
                         "r"(live ? r[4 * j] : 0u), "r"(live ? r[4 * j + 1] : 0u), "r"(live ? r[4 * j + 2] : 0u), "r"(live ? r[4 * j + 3] : 0u)
                         : "memory");
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            const int col0 = n0 + (quarter * (BN / 128) + c) * 32;
            asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                             reinterpret_cast<uint64_t>(tmap_out)),
                         "r"(col0), "r"(m0 + (warp & 3) * 32), "r"(stg)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (threadIdx.x == 0 && it < 3) gemm_trace(p, 9 + it);
      } else {
        gemm_epilogue_row<EPI>(tmem_base + lane_base + acc * BN, quarter * (BN / 128), BN / 128, row, n0, p);
        ptx::tc_fence_before();
        __syncwarp();
        if (threadIdx.x == 0 && it < 3) gemm_trace(p, 9 + it);  // slots 9-11: warp 0 drained its part of the tile
        if (lane == 0) ptx::mbar_arrive_cluster(&acc_empty_bar[acc], 0);
      }
    }
  }
  if constexpr (kTmaAcc) {
    if (warp < kPairEpiWarps && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the reduces have landed
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (threadIdx.x == 0) gemm_trace(p, 12);  // slot 12: every warp of the pair is done
  if (warp == kTmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------ attention (tensor cores)
// One CTA per (sequence, head), one warp per 16 query rows.  K (row-major) and V (transposed) of
// the head live in shared memory as fp16; scores and context run on mma.sync m16n8k16 (fp16 in,
// fp32 accumulate) with an online softmax over blocks of 64 keys, so a sequence of any length up
// to the position table needs the same registers.  The attention tiles of this model family
// (S <= 512, head size 32 / 64) are far too small for a 128-row tcgen05 tile; it is < 2 % of the
// encoder's FLOPs.  Additive -inf key mask and fp32 softmax as transformers.BertModel.
__device__ __forceinline__ void mma_m16n8k16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int DH>
__global__ void __launch_bounds__(256)
attention_mma_kernel(const __half* __restrict__ qkv, const int* __restrict__ mask, int seq, int hidden,
                     int heads, __half* __restrict__ ctx) {
  constexpr int kKSteps = DH / 16;   // k-steps of Q.K^T
  constexpr int kOTiles = DH / 8;    // n-tiles of the context
  constexpr int kKPitch = DH + 8;    // halves; (DH+8)/2 words = 4 * odd: conflict-free fragment loads
  extern __shared__ __align__(16) uint8_t asm_raw[];
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int seq_pad = (seq + 15) & ~15;
  const int vpitch = seq_pad + 8;
  __half* ks = reinterpret_cast<__half*>(asm_raw);             // [seq_pad][kKPitch]
  __half* vt = ks + static_cast<size_t>(seq_pad) * kKPitch;    // [DH][vpitch]  (V transposed)
  float* bias = reinterpret_cast<float*>(vt + static_cast<size_t>(DH) * vpitch);  // [seq_pad], log2 domain
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const size_t row_stride = static_cast<size_t>(3) * hidden;
  const __half* base = qkv + static_cast<size_t>(b) * seq * row_stride + h * DH;

  for (int i = threadIdx.x; i < seq_pad * (DH / 2); i += blockDim.x) {
    const int j = i / (DH / 2), c = i % (DH / 2);
    __half2 kk = __floats2half2_rn(0.f, 0.f), vv = kk;
    if (j < seq) {
      kk = *reinterpret_cast<const __half2*>(base + j * row_stride + hidden + 2 * c);
      vv = *reinterpret_cast<const __half2*>(base + j * row_stride + 2 * hidden + 2 * c);
    }
    *reinterpret_cast<__half2*>(ks + j * kKPitch + 2 * c) = kk;
    vt[(2 * c) * vpitch + j] = __low2half(vv);
    vt[(2 * c + 1) * vpitch + j] = __high2half(vv);
  }
  for (int j = threadIdx.x; j < seq_pad; j += blockDim.x)
    bias[j] = (j < seq && mask[b * seq + j] != 0) ? 0.f : -CUDART_INF_F;
  __syncthreads();

  const float scale = rsqrtf(static_cast<float>(DH)) * 1.4426950408889634f;  // softmax in base 2
  for (int rt = warp; rt * 16 < seq_pad; rt += nwarps) {
    const int r0 = rt * 16 + g, r1 = r0 + 8;
    // Q fragments (A operand), straight from global memory
    uint32_t qa[kKSteps][4];
#pragma unroll
    for (int kk = 0; kk < kKSteps; ++kk) {
      const int c = kk * 16 + 2 * t;
      qa[kk][0] = r0 < seq ? *reinterpret_cast<const uint32_t*>(base + r0 * row_stride + c) : 0u;
      qa[kk][1] = r1 < seq ? *reinterpret_cast<const uint32_t*>(base + r1 * row_stride + c) : 0u;
      qa[kk][2] = r0 < seq ? *reinterpret_cast<const uint32_t*>(base + r0 * row_stride + c + 8) : 0u;
      qa[kk][3] = r1 < seq ? *reinterpret_cast<const uint32_t*>(base + r1 * row_stride + c + 8) : 0u;
    }
    float o[kOTiles][4];
#pragma unroll
    for (int n = 0; n < kOTiles; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, l0 = 0.f, l1 = 0.f;

    for (int kb0 = 0; kb0 < seq_pad; kb0 += 64) {
      const int ntiles = min(8, (seq_pad - kb0) >> 3);  // warp-uniform, even
      float sc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
        if (j < ntiles) {
          const __half* kr = ks + (kb0 + j * 8 + g) * kKPitch + 2 * t;
#pragma unroll
          for (int kk = 0; kk < kKSteps; ++kk)
            mma_m16n8k16(sc[j], qa[kk], *reinterpret_cast<const uint32_t*>(kr + kk * 16),
                         *reinterpret_cast<const uint32_t*>(kr + kk * 16 + 8));
        }
      }
      // scale + key mask, block row maxima
      float bm0 = -CUDART_INF_F, bm1 = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < ntiles) {
          const float b0 = bias[kb0 + j * 8 + 2 * t], b1 = bias[kb0 + j * 8 + 2 * t + 1];
          sc[j][0] = sc[j][0] * scale + b0;
          sc[j][1] = sc[j][1] * scale + b1;
          sc[j][2] = sc[j][2] * scale + b0;
          sc[j][3] = sc[j][3] * scale + b1;
          bm0 = fmaxf(bm0, fmaxf(sc[j][0], sc[j][1]));
          bm1 = fmaxf(bm1, fmaxf(sc[j][2], sc[j][3]));
        }
      }
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
      const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);
      // every key masked so far: keep exponents finite (the row then sums to 0 and yields 0)
      const float mu0 = mn0 == -CUDART_INF_F ? 0.f : mn0, mu1 = mn1 == -CUDART_INF_F ? 0.f : mn1;
      const float corr0 = exp2f(m0 - mu0), corr1 = exp2f(m1 - mu1);
      m0 = mn0;
      m1 = mn1;
      float s0 = 0.f, s1 = 0.f;
      uint32_t pa[4][4];  // P as A fragments: k-step kk covers keys kb0 + 16 kk .. + 15
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
        if (j < ntiles) {
          p0 = exp2f(sc[j][0] - mu0);
          p1 = exp2f(sc[j][1] - mu0);
          p2 = exp2f(sc[j][2] - mu1);
          p3 = exp2f(sc[j][3] - mu1);
        }
        s0 += p0 + p1;
        s1 += p2 + p3;
        pa[j >> 1][(j & 1) * 2] = pack_half2(p0, p1);
        pa[j >> 1][(j & 1) * 2 + 1] = pack_half2(p2, p3);
      }
      l0 = l0 * corr0 + s0;
      l1 = l1 * corr1 + s1;
#pragma unroll
      for (int n = 0; n < kOTiles; ++n) {
        o[n][0] *= corr0;
        o[n][1] *= corr0;
        o[n][2] *= corr1;
        o[n][3] *= corr1;
      }
      // context += P . V
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (kk * 2 < ntiles) {
#pragma unroll
          for (int n = 0; n < kOTiles; ++n) {
            const __half* vr = vt + (n * 8 + g) * vpitch + kb0 + kk * 16 + 2 * t;
            mma_m16n8k16(o[n], pa[kk], *reinterpret_cast<const uint32_t*>(vr), *reinterpret_cast<const uint32_t*>(vr + 8));
          }
        }
      }
    }
    // row sums live in the quad
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
    __half* out0 = ctx + (static_cast<size_t>(b) * seq + r0) * hidden + h * DH + 2 * t;
    __half* out1 = ctx + (static_cast<size_t>(b) * seq + r1) * hidden + h * DH + 2 * t;
#pragma unroll
    for (int n = 0; n < kOTiles; ++n) {
      if (r0 < seq) *reinterpret_cast<__half2*>(out0 + n * 8) = __floats2half2_rn(o[n][0] * inv0, o[n][1] * inv0);
      if (r1 < seq) *reinterpret_cast<__half2*>(out1 + n * 8) = __floats2half2_rn(o[n][2] * inv1, o[n][3] * inv1);
    }
  }
}

// ------------------------------------------------------------------ attention (scalar fallback)
// One CTA per (sequence, head); K and V of the head live in shared memory (fp16), one warp per
// query row at a time: scores over keys (lane = key), fp32 softmax with the additive -inf key
// mask of BertModel, then context (lane = feature).  CUDA-core kernel, used only when the head size is
// not a multiple of 16 (no shipped model).
constexpr int kAttnThreads = 256;

static __global__ void __launch_bounds__(kAttnThreads)
attention_scalar_kernel(const __half* __restrict__ qkv, const int* __restrict__ mask, int seq, int hidden,
                 int heads, __half* __restrict__ ctx) {
  extern __shared__ __align__(16) uint8_t asm_raw[];
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  const int dh = hidden / heads;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int kpitch = dh + 2;  // fp16 elements; odd word pitch -> conflict-free key-strided reads
  __half* ks = reinterpret_cast<__half*>(asm_raw);
  __half* vs = ks + static_cast<size_t>(seq) * kpitch;
  float* bias = reinterpret_cast<float*>(vs + static_cast<size_t>(seq) * kpitch);
  float* probs = bias + seq;  // [warps, seq]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const size_t row_stride = static_cast<size_t>(3) * hidden;
  const __half* base = qkv + static_cast<size_t>(b) * seq * row_stride;
  for (int i = threadIdx.x; i < seq * (dh / 2); i += blockDim.x) {
    const int j = i / (dh / 2), c = i % (dh / 2);
    const __half2 kk = *reinterpret_cast<const __half2*>(base + j * row_stride + hidden + h * dh + 2 * c);
    const __half2 vv = *reinterpret_cast<const __half2*>(base + j * row_stride + 2 * hidden + h * dh + 2 * c);
    *reinterpret_cast<__half2*>(ks + j * kpitch + 2 * c) = kk;
    *reinterpret_cast<__half2*>(vs + j * kpitch + 2 * c) = vv;
  }
  for (int j = threadIdx.x; j < seq; j += blockDim.x)
    bias[j] = mask[b * seq + j] != 0 ? 0.f : -CUDART_INF_F;
  __syncthreads();
  const float scale = rsqrtf(static_cast<float>(dh));
  float* pw = probs + warp * seq;
  for (int i = warp; i < seq; i += nwarps) {
    // q row in registers: lane holds features 2*lane, 2*lane+1 (dh <= 64)
    const __half* qp = base + i * row_stride + h * dh;
    float2 qv = make_float2(0.f, 0.f);
    if (2 * lane < dh) qv = __half22float2(*reinterpret_cast<const __half2*>(qp + 2 * lane));
    float mx = -CUDART_INF_F;
    for (int j0 = 0; j0 < seq; j0 += 32) {
      const int j = j0 + lane;
      float s = 0.f;
      for (int c = 0; c < dh / 2; ++c) {
        const float qx = __shfl_sync(0xffffffffu, qv.x, c), qy = __shfl_sync(0xffffffffu, qv.y, c);
        if (j < seq) {
          const float2 kk = __half22float2(*reinterpret_cast<const __half2*>(ks + j * kpitch + 2 * c));
          s = fmaf(qx, kk.x, s);
          s = fmaf(qy, kk.y, s);
        }
      }
      if (j < seq) {
        s = s * scale + bias[j];
        pw[j] = s;
        mx = fmaxf(mx, s);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (mx == -CUDART_INF_F) mx = 0.f;  // every key masked: BertModel yields a uniform row; value unused
    float sum = 0.f;
    __syncwarp();
    for (int j = lane; j < seq; j += 32) {
      const float e = __expf(pw[j] - mx);
      pw[j] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    __syncwarp();
    float2 acc = make_float2(0.f, 0.f);
    if (2 * lane < dh) {
      for (int j = 0; j < seq; ++j) {
        const float pj = pw[j];
        const float2 vv = __half22float2(*reinterpret_cast<const __half2*>(vs + j * kpitch + 2 * lane));
        acc.x = fmaf(pj, vv.x, acc.x);
        acc.y = fmaf(pj, vv.y, acc.y);
      }
      *reinterpret_cast<__half2*>(ctx + (static_cast<size_t>(b) * seq + i) * hidden + h * dh + 2 * lane) =
          __floats2half2_rn(acc.x * inv, acc.y * inv);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ pooling + L2 normalise
// sentence-transformers Pooling (mean over unmasked tokens, clamp(sum_mask, 1e-9) / CLS) followed
// by Normalize (x / max(||x||_2, 1e-12)).  One CTA per sequence, fp32 output.
static __global__ void __launch_bounds__(256)
pool_normalize_kernel(const __half* __restrict__ hs, const int* __restrict__ mask, int seq, int hidden,
                      int pool_cls, float* __restrict__ out) {
  ptx::pdl_wait();
  const int b = blockIdx.x;
  __shared__ float red[8];
  __shared__ float s_cnt;
  extern __shared__ float pooled[];
  if (threadIdx.x == 0) {
    float c = 0.f;
    for (int j = 0; j < seq; ++j) c += mask[b * seq + j] != 0 ? 1.f : 0.f;
    s_cnt = fmaxf(c, 1e-9f);
  }
  __syncthreads();
  float ss = 0.f;
  for (int c = threadIdx.x; c < hidden; c += blockDim.x) {
    float v;
    if (pool_cls) {
      v = __half2float(hs[static_cast<size_t>(b) * seq * hidden + c]);
    } else {
      float acc = 0.f;
      for (int j = 0; j < seq; ++j)
        if (mask[b * seq + j] != 0) acc += __half2float(hs[(static_cast<size_t>(b) * seq + j) * hidden + c]);
      v = acc / s_cnt;
    }
    pooled[c] = v;
    ss += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) tot += red[w];
  const float inv = 1.f / fmaxf(sqrtf(tot), 1e-12f);
  for (int c = threadIdx.x; c < hidden; c += blockDim.x) out[static_cast<size_t>(b) * hidden + c] = pooled[c] * inv;
}

}  // namespace lxg
