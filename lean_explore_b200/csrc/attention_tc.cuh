// Causal GQA attention of the Qwen3 backbone on the 5th-generation tensor cores (head size 128):
// the contraction pair of `Qwen3Attention.forward` that the reference runs through
// sentence-transformers / transformers (src/lean_explore/util/embedding_client.py:97-99,
// src/lean_explore/util/reranker_client.py:137-168).
//
// One CTA (8 warps) per (128 query rows, q head, sequence); threads t and t + 128 share query row t =
// TMEM lane t: each owns 64 of a chunk's 128 keys (and 64 of the 128 context dims).
//   * TMA: the Q tile and, per chunk of 128 keys, the K and V tiles of the head's KV group come
//     straight out of the packed [token][q heads | k heads | v heads] activation matrix as two
//     128B-swizzled boxes of 128 rows x 64 dims each (one tensor map serves all three).
//   * S = Q.K^T: tcgen05.mma (SS, M = 128, N = 128, 8 steps of K = 16), fp32 scores in TENSOR MEMORY.
//   * Online softmax in fp32 by the row's two threads (tcgen05.ld, base-2 exponentials, maxima and
//     sums exchanged through shared memory); the probabilities
//     go back to tensor memory as fp16 - over the columns of S the thread has already consumed - and
//     are the A operand of the second contraction.
//   * O += P.V: tcgen05.mma (TS: A from tensor memory, B = the V tile read MN-MAJOR - V lies
//     [key][dim] in memory, exactly the image TMA delivers - b_major bit of the instruction
//     descriptor), fp32 context accumulator in tensor memory (128 columns), rescaled in place when a
//     row's running maximum moves.
// A CTA runs its steps one after the other; two CTAs share an SM (96 KB of shared memory and 256
// tensor-memory columns each), so one CTA's softmax overlaps the other's contractions.
// Key j is visible to query i iff j <= i and mask[j] != 0 (HF create_causal_mask with a padding
// mask); rows with no visible key (left padding) yield 0 and are never read downstream.  Packed
// batches (cu != NULL, no padding tokens at all) give sequence b the tokens [cu[b], cu[b+1]).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "ptx.cuh"

namespace lxg {

constexpr int kTcAttnRows = 128;  // query rows per CTA == UMMA M
constexpr int kTcAttnKeys = 128;  // keys per chunk == UMMA N of S == K extent of P.V
constexpr int kTcAttnDH = 128;    // head size
constexpr int kTcAttnBox = kTcAttnRows * 128;                 // one TMA box: 128 rows x 64 fp16
constexpr int kTcAttnSmem = 6 * kTcAttnBox + 1024;            // Q, K, V: two boxes each (+ alignment)
constexpr uint32_t kTcAttnTmemCols = 256;                     // S / P at column 0, O at column 128

namespace ptx {
// registers -> TMEM: this warp's 32 lanes x 32 consecutive 32-bit columns, then wait for the stores
__device__ __forceinline__ void tmem_st_x32_wait(uint32_t taddr, const uint32_t (&r)[32]) {
  tmem_st_32x32b_x32(taddr, r);
  tc_wait_st();
}
// Shared-memory matrix descriptor of an MN-major operand staged by TMA with SWIZZLE_128B as boxes of
// [rows = K index][64 x 16-bit = 128 bytes of MN index]: canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO))
// in 16-byte units - 8-row groups of 1024 bytes along K (SBO), `box_bytes` between the 64-element
// groups along MN (LBO).
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t box_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((box_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
}  // namespace ptx

__device__ __forceinline__ uint32_t pack_half2_rn(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

constexpr int kTcAttnThreads = 2 * kTcAttnRows;  // two threads per query row: each owns 64 of a chunk's 128 keys

__global__ void __launch_bounds__(kTcAttnThreads, 2)
attention_causal_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const int* __restrict__ mask,
                           const int* __restrict__ cu, int seq, int heads, int kv_heads, __half* __restrict__ ctx) {
  extern __shared__ uint8_t tc_attn_smem[];
  __shared__ __align__(8) uint64_t bar_q, bar_k, bar_v, bar_mma;
  __shared__ uint32_t tmem_holder;
  __shared__ float bias_sm[kTcAttnKeys];
  __shared__ float xch[2][kTcAttnRows];  // the two halves of a row exchange their maxima / sums

  const int tid = threadIdx.x, warp = tid >> 5;
  const int half = tid >> 7;               // which 64 keys of every chunk (and which 64 context dims) this thread owns
  const int rt = tid & (kTcAttnRows - 1);  // query row of the tile == TMEM lane
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const uint32_t smem0 = (ptx::smem_u32(tc_attn_smem) + 1023u) & ~1023u;
  const uint32_t smem_q = smem0, smem_k = smem0 + 2 * kTcAttnBox, smem_v = smem0 + 4 * kTcAttnBox;

  if (tid == 0) {
    ptx::mbar_init(&bar_q, 1);
    ptx::mbar_init(&bar_k, 1);
    ptx::mbar_init(&bar_v, 1);
    ptx::mbar_init(&bar_mma, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&map_qkv);
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_holder, kTcAttnTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;  // a warp reaches the TMEM lanes of its quarter
  const uint32_t tm_s = tmem_base + lane_base + half * 64;         // this thread's 64 score columns (fp32)
  const uint32_t tm_p = tmem_base + lane_base + half * 32;         // ... become 32 columns of fp16 pairs, over consumed scores
  const uint32_t tm_o = tmem_base + lane_base + 128u + half * 64;  // this thread's 64 context columns (fp32)

  ptx::pdl_wait();  // the activations are the previous kernel's output
  ptx::pdl_launch_dependents();
  // packed batch (cu != NULL): sequence b is tokens [cu[b], cu[b+1]), every key is real
  const int tok0 = cu != nullptr ? cu[b] : b * seq;
  if (cu != nullptr) seq = cu[b + 1] - tok0;
  const bool cta_live = qb * kTcAttnRows < seq;  // warp-uniform (whole CTA)
  const int kvh = h / (heads / kv_heads);
  const int col_q = h * kTcAttnDH, col_k = (heads + kvh) * kTcAttnDH, col_v = (heads + kv_heads + kvh) * kTcAttnDH;
  const int row = qb * kTcAttnRows + rt;  // this thread's query row within the sequence
  const int key_end = min(seq, (qb + 1) * kTcAttnRows);  // causal: no key beyond the block's last row
  const int nchunks = cta_live ? (key_end + kTcAttnKeys - 1) / kTcAttnKeys : 0;

  constexpr uint32_t kIdescQK = ptx::make_idesc_f16(kTcAttnRows, kTcAttnKeys);
  constexpr uint32_t kIdescPV = ptx::make_idesc_f16(kTcAttnRows, kTcAttnDH) | (1u << 16);  // B (= V) is MN-major
  constexpr uint32_t kDescHiK = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO, version, SWIZZLE_128B (see ptx.cuh)

  auto load_tile = [&](uint32_t dst, int col, int r0, uint64_t* bar) {  // thread 0 only
    ptx::mbar_arrive_expect_tx(bar, 2 * kTcAttnBox);
    ptx::tma_load_2d_a(dst, &map_qkv, col, r0, ptx::smem_u32(bar), ptx::kEvictNormal);
    ptx::tma_load_2d_a(dst + kTcAttnBox, &map_qkv, col + 64, r0, ptx::smem_u32(bar), ptx::kEvictNormal);
  };
  if (tid == 0 && cta_live) {
    load_tile(smem_q, col_q, tok0 + qb * kTcAttnRows, &bar_q);
    load_tile(smem_k, col_k, tok0, &bar_k);
    load_tile(smem_v, col_v, tok0, &bar_v);
  }

  const float scale = rsqrtf(static_cast<float>(kTcAttnDH)) * 1.4426950408889634f;  // softmax in base 2
  float m_run = -CUDART_INF_F;  // the row's running maximum (both halves hold the same value)
  float l_half = 0.f;           // this half's share of the row's running sum
  uint32_t mma_phase = 0;
  for (int c = 0; c < nchunks; ++c) {
    const int kb0 = c * kTcAttnKeys;
    // key bias of the chunk: 0 for a visible key, -inf for padding / beyond the sequence
    if (tid < kTcAttnKeys) {
      const int j = kb0 + tid;
      bias_sm[tid] = (j < seq && (cu != nullptr || mask[tok0 + j] != 0)) ? 0.f : -CUDART_INF_F;
    }
    // ---- S = Q . K^T
    if (warp == 0) {
      if (c == 0) ptx::mbar_wait(&bar_q, 0);
      ptx::mbar_wait(&bar_k, c & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t qlo = ((smem_q & 0x3FFFFu) >> 4) | (1u << 16), klo = ((smem_k & 0x3FFFFu) >> 4) | (1u << 16);
#pragma unroll
        for (int bx = 0; bx < 2; ++bx) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t ad = (static_cast<uint64_t>(kDescHiK) << 32) | (qlo + bx * (kTcAttnBox >> 4) + k4 * 2);
            const uint64_t bd = (static_cast<uint64_t>(kDescHiK) << 32) | (klo + bx * (kTcAttnBox >> 4) + k4 * 2);
            ptx::mma_f16_ss(tmem_base, ad, bd, kIdescQK, (bx | k4) != 0 ? 1u : 0u);
          }
        }
        ptx::tc_commit(&bar_mma);
      }
      __syncwarp();
    }
    __syncthreads();  // bias_sm visible
    ptx::mbar_wait(&bar_mma, mma_phase);
    mma_phase ^= 1u;
    ptx::tc_fence_after();
    // K tile is free: prefetch the next chunk's keys under the softmax
    if (tid == 0 && c + 1 < nchunks) load_tile(smem_k, col_k, tok0 + kb0 + kTcAttnKeys, &bar_k);

    // ---- online softmax: this thread's 64 scores of the row, both loads in flight at once
    uint32_t r0[32], r1[32];
    ptx::tmem_ld_32x32b_x32(tm_s, r0);
    ptx::tmem_ld_32x32b_x32(tm_s + 32, r1);
    ptx::tc_wait_ld();
    float sc0[32], sc1[32];
    float cmax = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int i0 = half * 64 + j, i1 = i0 + 32;
      sc0[j] = kb0 + i0 <= row ? __uint_as_float(r0[j]) * scale + bias_sm[i0] : -CUDART_INF_F;
      sc1[j] = kb0 + i1 <= row ? __uint_as_float(r1[j]) * scale + bias_sm[i1] : -CUDART_INF_F;
      cmax = fmaxf(cmax, fmaxf(sc0[j], sc1[j]));
    }
    xch[half][rt] = cmax;
    __syncthreads();  // (also: both halves hold their scores in registers - the P stores below may overwrite them)
    const float m_new = fmaxf(m_run, fmaxf(xch[0][rt], xch[1][rt]));
    const float mu = m_new == -CUDART_INF_F ? 0.f : m_new;
    const float corr = exp2f(m_run - mu);  // m_run == -inf: 0 (nothing accumulated yet)
    float psum = 0.f;
    uint32_t pk[32];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float a = exp2f(sc0[j] - mu), bb = exp2f(sc0[j + 1] - mu);
      const float cc = exp2f(sc1[j] - mu), dd = exp2f(sc1[j + 1] - mu);
      psum += (a + bb) + (cc + dd);
      pk[j >> 1] = pack_half2_rn(a, bb);
      pk[16 + (j >> 1)] = pack_half2_rn(cc, dd);
    }
    ptx::tmem_st_32x32b_x32(tm_p, pk);  // P columns [32 half, 32 half + 32) <- keys [64 half, 64 half + 64)
    l_half = l_half * corr + psum;
    // rescale the context accumulated so far when any row of the warp moved its maximum
    if (c > 0 && __any_sync(0xffffffffu, m_new > m_run)) {
#pragma unroll 1
      for (int q2 = 0; q2 < 2; ++q2) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tm_o + q2 * 32, r);
        ptx::tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * corr);
        ptx::tmem_st_32x32b_x32(tm_o + q2 * 32, r);
      }
    }
    ptx::tc_wait_st();
    m_run = m_new;
    ptx::tc_fence_before();
    __syncthreads();  // every row's P (and rescaled O) is in tensor memory
    // ---- O += P . V
    if (warp == 0) {
      ptx::mbar_wait(&bar_v, c & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kTcAttnKeys / 16; ++kk) {
          // 16 keys = two 8-row swizzle atoms = 2048 bytes down the V tile; 8 columns of P
          const uint64_t vd = ptx::make_mnmajor_sw128_desc(smem_v + kk * 2048, kTcAttnBox);
          ptx::mma_f16_ts(tmem_base + 128u, tmem_base + kk * 8, vd, kIdescPV, (c | kk) != 0 ? 1u : 0u);
        }
        ptx::tc_commit(&bar_mma);
      }
      __syncwarp();
    }
    ptx::mbar_wait(&bar_mma, mma_phase);
    mma_phase ^= 1u;
    ptx::tc_fence_after();
    // V tile is free
    if (tid == 0 && c + 1 < nchunks) load_tile(smem_v, col_v, tok0 + kb0 + kTcAttnKeys, &bar_v);
  }

  // ---- context row / l  -> fp16 (each half writes its 64 dims)
  if (cta_live) {
    xch[half][rt] = l_half;
    __syncthreads();
    const float l_run = xch[0][rt] + xch[1][rt];
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    __half* out = ctx + (static_cast<size_t>(tok0) + row) * (static_cast<size_t>(heads) * kTcAttnDH) + h * kTcAttnDH + half * 64;
    uint32_t r0[32], r1[32];
    ptx::tmem_ld_32x32b_x32(tm_o, r0);
    ptx::tmem_ld_32x32b_x32(tm_o + 32, r1);
    ptx::tc_wait_ld();
    if (row < seq) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 v;
        v.x = pack_half2_rn(__uint_as_float(r0[j]) * inv, __uint_as_float(r0[j + 1]) * inv);
        v.y = pack_half2_rn(__uint_as_float(r0[j + 2]) * inv, __uint_as_float(r0[j + 3]) * inv);
        v.z = pack_half2_rn(__uint_as_float(r0[j + 4]) * inv, __uint_as_float(r0[j + 5]) * inv);
        v.w = pack_half2_rn(__uint_as_float(r0[j + 6]) * inv, __uint_as_float(r0[j + 7]) * inv);
        *reinterpret_cast<uint4*>(out + j) = v;
        v.x = pack_half2_rn(__uint_as_float(r1[j]) * inv, __uint_as_float(r1[j + 1]) * inv);
        v.y = pack_half2_rn(__uint_as_float(r1[j + 2]) * inv, __uint_as_float(r1[j + 3]) * inv);
        v.z = pack_half2_rn(__uint_as_float(r1[j + 4]) * inv, __uint_as_float(r1[j + 5]) * inv);
        v.w = pack_half2_rn(__uint_as_float(r1[j + 6]) * inv, __uint_as_float(r1[j + 7]) * inv);
        *reinterpret_cast<uint4*>(out + 32 + j) = v;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTcAttnTmemCols);
  }
}

}  // namespace lxg
