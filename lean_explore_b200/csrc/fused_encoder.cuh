// The query path of the BERT-class sentence encoder as ONE persistent kernel (<= 64 tokens per call:
// what EmbeddingClient.embed sends for a search query, reference
// src/lean_explore/util/embedding_client.py:88-101 called from search/engine.py:236).
//
// Why: at B = 1 the layered path (encoder_kernels.cuh) is 2 + 7 L dependent launches of 3-9 us each
// while the arithmetic is a 21 MB (MiniLM-L6) / 170 MB (bge-base) weight read - microseconds of HBM
// time.  Here the whole forward is one cooperative launch of one CTA per SM:
//
//   * Skinny GEMMs are computed transposed ("swap AB"): the weight tile W[128 features, 64 k] is the
//     M side of tcgen05.mma (M = 128), the tokens are the N side (N = tokens padded to 16 .. 64), so
//     the accumulator in TMEM is 128 lanes (features) x <= 64 columns (tokens) and no tensor-core
//     work is spent on padding rows.
//   * Weights are read exactly once per forward, by TMA, straight into a shared-memory ring that the
//     tensor core reads (SS MMA) - they never touch registers.  The list of (layer, phase, tile) jobs
//     a CTA owns is a pure function of its block index, so the TMA warp streams the weights of FUTURE
//     phases into the ring while the CTA waits at a phase boundary: weight traffic is decoupled from
//     the dependency chain, which only carries the small activations.
//   * Four phases per layer, separated by a grid-wide barrier (one counter per phase, release/acquire):
//       A  LayerNorm-on-load (previous layer's LN2, or the embedding sum + embedding LN) -> q|k|v of ONE
//          head -> attention of that head in the same CTA (mma.sync, fp32 softmax) -> ctx
//       B  attention.output.dense + bias + residual            -> pre-LN1 (fp32)
//       C  LayerNorm-on-load (LN1) -> intermediate.dense + erf-GELU
//       D  output.dense, K split in F / H slices; the last CTA to finish a tile sums the partial
//          slabs in slice order (deterministic) + bias + residual -> pre-LN2 (fp32)
//     LayerNorm needs whole rows, so every consumer CTA normalises the rows it loads itself (the rows
//     are <= 64 x 1024 values; recomputing beats one more grid barrier).
//   * Everything on the dependency chain is latency, not throughput, so the chain is kept short:
//     parameters (gamma / beta / bias) and residual rows are fetched before the phase barrier / the
//     accumulator is awaited; LayerNorm handles four rows per warp in lock step (the shuffle chains
//     overlap); each k-block's four K = 16 MMAs go to four independent TMEM accumulators, summed by
//     the epilogue (a chain of dependent N = 16 MMAs costs ~120 clocks per instruction).
//   * Pooling (mean over unmasked tokens / CLS) + L2 normalise run in CTA 0 after the last phase.
//
// Numerics are those of the layered path: fp16 operands, fp32 accumulation / LayerNorm / softmax,
// activations rounded to fp16 where that path stores them as fp16.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "encoder_kernels.cuh"
#include "ptx.cuh"

namespace lxg {

constexpr int kFusedMaxTokens = 64;
constexpr int kFusedComputeWarps = 8;
constexpr int kFusedComputeThreads = kFusedComputeWarps * 32;
constexpr int kFusedThreads = kFusedComputeThreads + 64;  // + TMA warp + MMA warp
constexpr int kFusedSlotBytes = 128 * 64 * 2;             // one k-block of a 128-row weight tile
constexpr int kFusedMaxSlots = 13;
constexpr int kFusedAcc = 4;                              // independent accumulators per weight tile
constexpr int kFusedTmemCols = 512;                       // 2 tiles x 4 accumulators x 64 token columns

struct alignas(64) FusedLayer {
  CUtensorMap map_qkv;               // wqkv [3H, H], box = head_dim rows x 64 columns
  CUtensorMap map_wo, map_w1, map_w2;  // box = 128 rows x 64 columns
  const float *bqkv, *bo, *ln1_g, *ln1_b, *b1, *b2, *ln2_g, *ln2_b;
};

struct FusedParams {
  const FusedLayer* layers;
  int num_layers;
  int tokens, seq, batch, tpad;  // tpad = tokens rounded up to 16
  int hidden, ffn, heads, vocab;
  float eps;
  const int* ids;
  const int* mask;
  const __half *word, *pos, *type0;
  const float *emb_g, *emb_b;
  __half *h0, *h1;    // [T, H] LayerNorm outputs kept for the residual adds (phase B reads h0, D reads h1)
  __half* ctx;        // [T, H]
  __half* act;        // [T, F] GELU(intermediate.dense)
  float *pre1, *pre2; // [T, H] pre-LayerNorm sums
  float* partial;     // [H/128][F/H][tpad][128] split-K slabs of phase D
  unsigned* sem;      // [H/128] slab counters (zero between launches)
  unsigned* bar;      // [4 L] phase counters (zero between launches)
  int pool_cls;
  float* out;         // [B, H]
  int nslots;         // weight ring depth
  unsigned long long* trace;  // optional [grid][4 L + 1][6] %globaltimer stamps of thread 0 (lxg_encoder_set_fused(enc, 2))
};

namespace fused {

__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(kFusedComputeThreads) : "memory"); }

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// One thread, after a CTA barrier that follows the CTA's stores (the barrier + this fence make them
// visible to whoever acquires the counter).
__device__ __forceinline__ void grid_arrive(unsigned* ctr) {
  __threadfence();
  atomicAdd(ctr, 1u);
}
__device__ __forceinline__ unsigned long long timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void grid_wait(const unsigned* ctr, unsigned target) {
  unsigned spins = 0;
  while (ld_acquire(ctr) < target) {
    if (++spins == (1u << 24)) __trap();  // a protocol bug must fail the launch, not hang the GPU
  }
}

// Byte offset of (token t, feature f) in the K-major, 128-byte-swizzled MMA operand of `tpad` rows:
// k-block f / 64 is a [tpad x 128 B] slab, rows are 128 B, the 16-byte chunk index is XORed with t & 7.
__device__ __forceinline__ uint32_t operand_offset(int t, int f, int tpad) {
  const int kb = f >> 6, c = (f & 63) >> 3;
  return static_cast<uint32_t>(kb * tpad * 128 + t * 128 + (((c ^ (t & 7)) << 4) | ((f & 7) << 1)));
}

// LayerNorm-on-load.  Warp w normalises tokens w, w + 8, w + 16, w + 24 in lock step (the loads and
// the two shuffle reductions of the four rows overlap), then the next four; lane holds the float4s
// lane, lane + 32, ... of a row.  g / bt: gamma / beta staged in shared memory before the phase barrier
// was awaited.  Writes fp16 into the swizzled operand (or row-major when `plain`), optionally also to
// global `hout` (the copy later residual adds read).
template <bool EMB, int NV>
__device__ __forceinline__ void stage_ln(const FusedParams& p, const float* src, const float* g, const float* bt, uint8_t* bsm,
                                         bool plain, __half* hout, int warp, int lane) {
  constexpr int R = NV >= 8 ? 2 : 4;  // rows in lock step (register budget)
  constexpr int H = NV * 128;
  for (int t0 = warp; t0 < p.tokens; t0 += R * kFusedComputeWarps) {
    float4 v[R][NV];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = t0 + r * kFusedComputeWarps;
      if (t < p.tokens) {
        if constexpr (EMB) {
          int id = p.ids[t];
          id = min(max(id, 0), p.vocab - 1);
          const uint2* w2 = reinterpret_cast<const uint2*>(p.word + static_cast<size_t>(id) * H);
          const uint2* p2 = reinterpret_cast<const uint2*>(p.pos + static_cast<size_t>(t % p.seq) * H);
          const uint2* t2 = reinterpret_cast<const uint2*>(p.type0);
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int j = lane + 32 * i;
            const uint2 a = __ldg(w2 + j), b = __ldg(p2 + j), c = __ldg(t2 + j);
            const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&a.x)), a1 = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
            const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(&b.x)), b1 = __half22float2(*reinterpret_cast<const __half2*>(&b.y));
            const float2 c0 = __half22float2(*reinterpret_cast<const __half2*>(&c.x)), c1 = __half22float2(*reinterpret_cast<const __half2*>(&c.y));
            v[r][i] = make_float4(a0.x + b0.x + c0.x, a0.y + b0.y + c0.y, a1.x + b1.x + c1.x, a1.y + b1.y + c1.y);
          }
        } else {
          const float4* s4 = reinterpret_cast<const float4*>(src + static_cast<size_t>(t) * H);
#pragma unroll
          for (int i = 0; i < NV; ++i) v[r][i] = __ldcg(s4 + lane + 32 * i);  // written by other CTAs in this launch: L2, not L1
        }
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[r][i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float mean[R], rstd[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      mean[r] = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) mean[r] += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < R; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      mean[r] *= 1.0f / H;
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float a = v[r][i].x - mean[r], b = v[r][i].y - mean[r], c = v[r][i].z - mean[r], d = v[r][i].w - mean[r];
        var += (a * a + b * b) + (c * c + d * d);
      }
      rstd[r] = var;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < R; ++r) rstd[r] += __shfl_xor_sync(0xffffffffu, rstd[r], o);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) rstd[r] = rsqrtf(rstd[r] * (1.0f / H) + p.eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int j = lane + 32 * i, f = 4 * j;
      const float4 gg = reinterpret_cast<const float4*>(g)[j], bb = reinterpret_cast<const float4*>(bt)[j];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int t = t0 + r * kFusedComputeWarps;
        if (t < p.tokens) {
          uint2 o;
          o.x = pack_half2((v[r][i].x - mean[r]) * rstd[r] * gg.x + bb.x, (v[r][i].y - mean[r]) * rstd[r] * gg.y + bb.y);
          o.y = pack_half2((v[r][i].z - mean[r]) * rstd[r] * gg.z + bb.z, (v[r][i].w - mean[r]) * rstd[r] * gg.w + bb.w);
          const uint32_t off = plain ? static_cast<uint32_t>((t * H + f) * 2) : operand_offset(t, f, p.tpad);
          *reinterpret_cast<uint2*>(bsm + off) = o;
          if (hout != nullptr) *reinterpret_cast<uint2*>(hout + static_cast<size_t>(t) * H + f) = o;
        }
      }
    }
  }
}

// fp16 activations [T, ld] (columns k0 .. k0 + H) -> swizzled operand.
__device__ __forceinline__ void stage_copy(const FusedParams& p, const __half* src, int ld, int k0, uint8_t* bsm, int tid) {
  const int cpr = p.hidden >> 3;  // 16-byte chunks per row
  const int total = p.tokens * cpr;
  for (int q0 = tid; q0 < total; q0 += 4 * kFusedComputeThreads) {  // four loads in flight per thread
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + u * kFusedComputeThreads;
      if (q < total) {
        const int t = q / cpr, cg = q - t * cpr;
        v[u] = __ldcg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(t) * ld + k0) + cg);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + u * kFusedComputeThreads;
      if (q < total) {
        const int t = q / cpr, cg = q - t * cpr;
        *reinterpret_cast<uint4*>(bsm + operand_offset(t, cg << 3, p.tpad)) = v[u];
      }
    }
  }
}

// Attention of one head over the call's tokens as ONE block-diagonal problem: query row r attends to
// key j iff tag[j] == r / seq (tag[j] = sequence of token j, or -1 when it is masked / padding).
// One warp per 16 query rows; scores and context on mma.sync m16n8k16 (see attention_mma_kernel).
template <int DH>
__device__ __forceinline__ void head_attention(const __half* qs, const __half* ks, const __half* vt, const int* tag, int tpad,
                                               int tokens, int seq, __half* ctx_head, int hidden, int warp, int lane) {
  constexpr int kKSteps = DH / 16, kOTiles = DH / 8, kKPitch = DH + 8;
  const int vpitch = tpad + 8;
  const int g = lane >> 2, t = lane & 3;
  const float scale = rsqrtf(static_cast<float>(DH)) * 1.4426950408889634f;
  for (int rt = warp; rt * 16 < tpad; rt += kFusedComputeWarps) {
    const int r0 = rt * 16 + g, r1 = r0 + 8;
    const int sq0 = r0 < tokens ? r0 / seq : -2, sq1 = r1 < tokens ? r1 / seq : -2;
    uint32_t qa[kKSteps][4];
#pragma unroll
    for (int kk = 0; kk < kKSteps; ++kk) {
      const int c = kk * 16 + 2 * t;
      qa[kk][0] = *reinterpret_cast<const uint32_t*>(qs + r0 * kKPitch + c);
      qa[kk][1] = *reinterpret_cast<const uint32_t*>(qs + r1 * kKPitch + c);
      qa[kk][2] = *reinterpret_cast<const uint32_t*>(qs + r0 * kKPitch + c + 8);
      qa[kk][3] = *reinterpret_cast<const uint32_t*>(qs + r1 * kKPitch + c + 8);
    }
    float o[kOTiles][4];
#pragma unroll
    for (int n = 0; n < kOTiles; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    const int ntiles = tpad >> 3;  // <= 8, even, warp-uniform
    float sc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
      if (j < ntiles) {
        const __half* kr = ks + (j * 8 + g) * kKPitch + 2 * t;
#pragma unroll
        for (int kk = 0; kk < kKSteps; ++kk)
          mma_m16n8k16(sc[j], qa[kk], *reinterpret_cast<const uint32_t*>(kr + kk * 16), *reinterpret_cast<const uint32_t*>(kr + kk * 16 + 8));
      }
    }
    float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < ntiles) {
        const int ta = tag[j * 8 + 2 * t], tb = tag[j * 8 + 2 * t + 1];
        sc[j][0] = ta == sq0 ? sc[j][0] * scale : -CUDART_INF_F;
        sc[j][1] = tb == sq0 ? sc[j][1] * scale : -CUDART_INF_F;
        sc[j][2] = ta == sq1 ? sc[j][2] * scale : -CUDART_INF_F;
        sc[j][3] = tb == sq1 ? sc[j][3] * scale : -CUDART_INF_F;
        m0 = fmaxf(m0, fmaxf(sc[j][0], sc[j][1]));
        m1 = fmaxf(m1, fmaxf(sc[j][2], sc[j][3]));
      }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float mu0 = m0 == -CUDART_INF_F ? 0.f : m0, mu1 = m1 == -CUDART_INF_F ? 0.f : m1;
    float l0 = 0.f, l1 = 0.f;
    uint32_t pa[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
      if (j < ntiles) {
        p0 = exp2f(sc[j][0] - mu0);
        p1 = exp2f(sc[j][1] - mu0);
        p2 = exp2f(sc[j][2] - mu1);
        p3 = exp2f(sc[j][3] - mu1);
      }
      l0 += p0 + p1;
      l1 += p2 + p3;
      pa[j >> 1][(j & 1) * 2] = pack_half2(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = pack_half2(p2, p3);
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk * 2 < ntiles) {
#pragma unroll
        for (int n = 0; n < kOTiles; ++n) {
          const __half* vr = vt + (n * 8 + g) * vpitch + kk * 16 + 2 * t;
          mma_m16n8k16(o[n], pa[kk], *reinterpret_cast<const uint32_t*>(vr), *reinterpret_cast<const uint32_t*>(vr + 8));
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
    __half* out0 = ctx_head + static_cast<size_t>(r0) * hidden + 2 * t;
    __half* out1 = ctx_head + static_cast<size_t>(r1) * hidden + 2 * t;
#pragma unroll
    for (int n = 0; n < kOTiles; ++n) {
      if (r0 < tokens) *reinterpret_cast<__half2*>(out0 + n * 8) = __floats2half2_rn(o[n][0] * inv0, o[n][1] * inv0);
      if (r1 < tokens) *reinterpret_cast<__half2*>(out1 + n * 8) = __floats2half2_rn(o[n][2] * inv1, o[n][3] * inv1);
    }
  }
}

// Sum of the four accumulators of tile `mt` over 8 token columns starting at `col` (this thread's lane).
__device__ __forceinline__ void load_acc8(uint32_t lane_taddr, int mt, int col, float (&v)[8]) {
  uint32_t r[kFusedAcc][8];
#pragma unroll
  for (int a = 0; a < kFusedAcc; ++a) ptx::tmem_ld_32x32b_x8(lane_taddr + (mt * kFusedAcc + a) * 64 + col, r[a]);
  ptx::tc_wait_ld();
#pragma unroll
  for (int j = 0; j < 8; ++j)
    v[j] = (__uint_as_float(r[0][j]) + __uint_as_float(r[1][j])) + (__uint_as_float(r[2][j]) + __uint_as_float(r[3][j]));
}

}  // namespace fused

// DH = head size (32 / 64), NV = hidden / 128.
template <int DH, int NV>
__global__ void __launch_bounds__(kFusedThreads, 1) bert_fused_kernel(const FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kFusedMaxSlots];
  __shared__ __align__(8) uint64_t empty_bar[kFusedMaxSlots];
  __shared__ __align__(8) uint64_t bready_bar;
  __shared__ __align__(8) uint64_t accfull_bar;
  __shared__ uint32_t tmem_base_holder;
  __shared__ int s_last;
  __shared__ float s_red[kFusedComputeWarps];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  constexpr int H = NV * 128;
  const int F = p.ffn, T = p.tokens, tpad = p.tpad;
  constexpr int KB = H >> 6;      // k-blocks of every job (phase D slices K = F into F / H jobs of H)
  const int KS = F / H;           // split-K factor of phase D
  const int njobs[4] = {p.heads, H >> 7, F >> 7, (H >> 7) * KS};
  constexpr int kTilesA = DH == 32 ? 1 : 2;  // head_dim 32: q|k|v stacked in one 128-row tile; 64: q|k, then v
  constexpr int kTmaWarp = kFusedComputeWarps, kMmaWarp = kFusedComputeWarps + 1;

  // shared-memory carve-up: weight ring | token operand | attention scratch | LayerNorm parameters
  const uint32_t base_u32 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base_u32 - ptx::smem_u32(smem_raw));
  const uint32_t ring_u32 = base_u32;
  const uint32_t bop_u32 = ring_u32 + p.nslots * kFusedSlotBytes;
  uint8_t* bop = base_ptr + p.nslots * kFusedSlotBytes;
  uint8_t* attn = bop + tpad * H * 2;
  __half* qs = reinterpret_cast<__half*>(attn);
  __half* ks = qs + tpad * (DH + 8);
  __half* vt = ks + tpad * (DH + 8);
  int* tag = reinterpret_cast<int*>(vt + DH * (tpad + 8));
  float* lng = reinterpret_cast<float*>(tag + tpad);  // LayerNorm gamma / beta of the rows being staged
  float* lnb = lng + H;

  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < p.nslots; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(&bready_bar, kFusedComputeThreads);
    ptx::mbar_init(&accfull_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == kTmaWarp) {
    ptx::tmem_alloc(&tmem_base_holder, kFusedTmemCols);
    ptx::tmem_relinquish();
  }
  if (warp < kFusedComputeWarps) {
    // rows >= tokens of the operand stay zero for the whole launch (their accumulator columns are never stored)
    for (int i = threadIdx.x; i < (tpad * H * 2) / 16; i += kFusedComputeThreads) reinterpret_cast<uint4*>(bop)[i] = make_uint4(0, 0, 0, 0);
    for (int j = threadIdx.x; j < tpad; j += kFusedComputeThreads) tag[j] = (j < T && p.mask[j] != 0) ? j / p.seq : -1;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;

  if (warp == kTmaWarp) {
    // ---------------- weight stream: every job this CTA owns, in program order, as far ahead as the ring allows
    const uint32_t full0 = ptx::smem_u32(&full_bar[0]), empty0 = ptx::smem_u32(&empty_bar[0]);
    uint32_t stage = 0, phase = 0;
    int jb = 0;
    for (int l = 0; l < p.num_layers; ++l) {
      const FusedLayer* L = p.layers + l;
      for (int ph = 0; ph < 4; ++ph) {
        const int n = njobs[ph];
        for (int i = ((cta - jb) % G + G) % G; i < n; i += G) {
          const int ntiles = ph == 0 ? kTilesA : 1;
          for (int mt = 0; mt < ntiles; ++mt) {
            for (int kb = 0; kb < KB; ++kb) {
              ptx::mbar_wait_a(empty0 + stage * 8, phase ^ 1u);
              if (ptx::elect_one()) {
                const uint32_t fb = full0 + stage * 8;
                const uint32_t dst = ring_u32 + stage * kFusedSlotBytes;
                if (ph == 0) {
                  if constexpr (DH == 32) {
                    ptx::mbar_arrive_expect_tx_a(fb, 3 * 32 * 128);
#pragma unroll
                    for (int w = 0; w < 3; ++w) ptx::tma_load_2d_a(dst + w * 32 * 128, &L->map_qkv, kb * 64, w * H + i * 32, fb, ptx::kEvictNormal);
                  } else {
                    if (mt == 0) {
                      ptx::mbar_arrive_expect_tx_a(fb, 2 * 64 * 128);
                      ptx::tma_load_2d_a(dst, &L->map_qkv, kb * 64, i * 64, fb, ptx::kEvictNormal);
                      ptx::tma_load_2d_a(dst + 64 * 128, &L->map_qkv, kb * 64, H + i * 64, fb, ptx::kEvictNormal);
                    } else {
                      ptx::mbar_arrive_expect_tx_a(fb, 64 * 128);
                      ptx::tma_load_2d_a(dst, &L->map_qkv, kb * 64, 2 * H + i * 64, fb, ptx::kEvictNormal);
                    }
                  }
                } else {
                  ptx::mbar_arrive_expect_tx_a(fb, kFusedSlotBytes);
                  if (ph == 1)
                    ptx::tma_load_2d_a(dst, &L->map_wo, kb * 64, i * 128, fb, ptx::kEvictNormal);
                  else if (ph == 2)
                    ptx::tma_load_2d_a(dst, &L->map_w1, kb * 64, i * 128, fb, ptx::kEvictNormal);
                  else
                    ptx::tma_load_2d_a(dst, &L->map_w2, (i % KS) * H + kb * 64, (i / KS) * 128, fb, ptx::kEvictNormal);
                }
              }
              __syncwarp();
              if (++stage == static_cast<uint32_t>(p.nslots)) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
        jb += n;
      }
    }
  } else if (warp == kMmaWarp) {
    // ---------------- tensor core: D[128 features, tpad tokens] (+)= W tile (ring) x operand^T; the four
    // K = 16 steps of a k-block accumulate into four different TMEM accumulators (independent chains)
    const uint32_t full0 = ptx::smem_u32(&full_bar[0]), empty0 = ptx::smem_u32(&empty_bar[0]);
    const uint32_t a_lo0 = ((ring_u32 & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_lo0 = ((bop_u32 & 0x3FFFFu) >> 4) | (1u << 16);
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t idesc = ptx::make_idesc_f16(128, tpad);
    uint32_t stage = 0, phase = 0, jcount = 0;
    int jb = 0;
    for (int l = 0; l < p.num_layers; ++l) {
      for (int ph = 0; ph < 4; ++ph) {
        const int n = njobs[ph];
        for (int i = ((cta - jb) % G + G) % G; i < n; i += G, ++jcount) {
          ptx::mbar_wait(&bready_bar, jcount & 1u);  // operand staged, previous accumulators drained
          ptx::tc_fence_after();
          const int ntiles = ph == 0 ? kTilesA : 1;
          for (int mt = 0; mt < ntiles; ++mt) {
            for (int kb = 0; kb < KB; ++kb) {
              ptx::mbar_wait_a(full0 + stage * 8, phase);
              ptx::tc_fence_after();
              if (ptx::elect_one()) {
                const uint32_t alo = a_lo0 + stage * (kFusedSlotBytes >> 4);
                const uint32_t blo = b_lo0 + kb * ((tpad * 128) >> 4);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                  const uint64_t adesc = (static_cast<uint64_t>(kDescHi) << 32) | (alo + k4 * 2);
                  const uint64_t bdesc = (static_cast<uint64_t>(kDescHi) << 32) | (blo + k4 * 2);
                  ptx::mma_f16_ss(tmem_base + (mt * kFusedAcc + k4) * 64, adesc, bdesc, idesc, kb != 0 ? 1u : 0u);
                }
                ptx::tc_commit_a(empty0 + stage * 8);
                if (mt == ntiles - 1 && kb == KB - 1) ptx::tc_commit(&accfull_bar);
              }
              __syncwarp();
              if (++stage == static_cast<uint32_t>(p.nslots)) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
        jb += n;
      }
    }
  } else {
    // ---------------- compute warps: operand staging, epilogues, attention, phase barriers.
    // Epilogue mapping: lane quarter q = warp & 3 <-> TMEM lanes (features) 32 q .. 32 q + 31,
    // hf = warp >> 2 <-> the first / second half of the token columns, in chunks of 8.
    const int tid = threadIdx.x;
    const int q = warp & 3, hf = warp >> 2;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int nch = tpad >> 4;  // chunks of 8 tokens per half: tokens (hf * nch + c) * 8 ..
    const int ntr = 4 * p.num_layers + 1;
    unsigned long long* tr = (p.trace != nullptr && tid == 0) ? p.trace + static_cast<size_t>(cta) * ntr * 6 : nullptr;
    uint32_t jcount = 0;
    int jb = 0, bar_idx = 0;
    for (int l = 0; l < p.num_layers; ++l) {
      const FusedLayer* L = p.layers + l;
      for (int ph = 0; ph < 4; ++ph, ++bar_idx) {
        const int n = njobs[ph];
        bool waited = bar_idx == 0;  // the first phase reads only weights and token ids
        if (tr) tr[bar_idx * 6 + 0] = fused::timer_ns();
        for (int i = ((cta - jb) % G + G) % G; i < n; i += G, ++jcount) {
          // ---- everything that does not depend on other CTAs comes BEFORE the phase barrier is awaited:
          // LayerNorm gamma / beta into shared memory, this thread's bias into a register
          float bias = 0.f, bias2 = 0.f;
          if (ph == 0) {
            const float* g = l == 0 ? p.emb_g : (L - 1)->ln2_g;
            const float* bt = l == 0 ? p.emb_b : (L - 1)->ln2_b;
            for (int j = tid; j < H / 4; j += kFusedComputeThreads) {
              reinterpret_cast<float4*>(lng)[j] = __ldg(reinterpret_cast<const float4*>(g) + j);
              reinterpret_cast<float4*>(lnb)[j] = __ldg(reinterpret_cast<const float4*>(bt) + j);
            }
            if constexpr (DH == 32) {
              if (q < 3) bias = __ldg(L->bqkv + q * H + i * DH + lane);
            } else {
              bias = __ldg(L->bqkv + (q >> 1) * H + i * DH + (q & 1) * 32 + lane);          // tile 0: q | k
              if (q < 2) bias2 = __ldg(L->bqkv + 2 * H + i * DH + (q & 1) * 32 + lane);     // tile 1: v
            }
          } else if (ph == 2) {
            for (int j = tid; j < H / 4; j += kFusedComputeThreads) {
              reinterpret_cast<float4*>(lng)[j] = __ldg(reinterpret_cast<const float4*>(L->ln1_g) + j);
              reinterpret_cast<float4*>(lnb)[j] = __ldg(reinterpret_cast<const float4*>(L->ln1_b) + j);
            }
            bias = __ldg(L->b1 + i * 128 + q * 32 + lane);
          } else if (ph == 1) {
            bias = __ldg(L->bo + i * 128 + q * 32 + lane);
          } else {
            bias = __ldg(L->b2 + (i / KS) * 128 + q * 32 + lane);
          }
          if (!waited) {
            if (tid == 0) fused::grid_wait(p.bar + bar_idx - 1, static_cast<unsigned>(G));
            waited = true;
          }
          fused::cbar();  // barrier passed, gamma / beta staged
          if (tr) tr[bar_idx * 6 + 1] = fused::timer_ns();
          // ---- stage the token operand
          if (ph == 0) {
            __half* hout = i == 0 ? p.h0 : nullptr;
            if (l == 0)
              fused::stage_ln<true, NV>(p, nullptr, lng, lnb, bop, false, hout, warp, lane);
            else
              fused::stage_ln<false, NV>(p, p.pre2, lng, lnb, bop, false, hout, warp, lane);
          } else if (ph == 1) {
            fused::stage_copy(p, p.ctx, H, 0, bop, tid);
          } else if (ph == 2) {
            fused::stage_ln<false, NV>(p, p.pre1, lng, lnb, bop, false, i == 0 ? p.h1 : nullptr, warp, lane);
          } else {
            fused::stage_copy(p, p.act, F, (i % KS) * H, bop, tid);
          }
          ptx::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's reads
          ptx::tc_fence_before();
          ptx::mbar_arrive(&bready_bar);
          if (tr) tr[bar_idx * 6 + 2] = fused::timer_ns();
          // residual rows (written a phase ago) on their way while the tensor core works
          const int f_out = (ph == 3 ? i / KS : i) * 128 + q * 32 + lane;
          __half res[4][8];
          if (ph == 1 || ph == 3) {
            const __half* hsrc = ph == 1 ? p.h0 : p.h1;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (c < nch) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const int t = (hf * nch + c) * 8 + j;
                  res[c][j] = t < T ? __ldcg(hsrc + static_cast<size_t>(t) * H + f_out) : __half();
                }
              }
          }
          ptx::mbar_wait(&accfull_bar, jcount & 1u);
          ptx::tc_fence_after();
          if (tr) tr[bar_idx * 6 + 3] = fused::timer_ns();

          // ---- epilogue: thread = feature (TMEM lane), registers = tokens
          if (ph == 0) {
#pragma unroll
            for (int mt = 0; mt < kTilesA; ++mt) {
              const int seg = DH == 32 ? q : mt * 2 + (q >> 1);  // 0 = q, 1 = k, 2 = v, 3 = unused lanes
              const int fh = DH == 32 ? lane : (q & 1) * 32 + lane;
              const float bs = mt == 0 ? bias : bias2;
              if (seg < 3) {
                for (int c = 0; c < nch; ++c) {
                  const int tk0 = (hf * nch + c) * 8;
                  float v[8];
                  fused::load_acc8(lane_taddr, mt, tk0, v);
                  if (seg == 2) {
                    __half* dst = vt + fh * (tpad + 8) + tk0;
#pragma unroll
                    for (int j = 0; j < 8; j += 2) *reinterpret_cast<__half2*>(dst + j) = __floats2half2_rn(v[j] + bs, v[j + 1] + bs);
                  } else {
                    __half* dst = (seg == 0 ? qs : ks) + tk0 * (DH + 8) + fh;
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[j * (DH + 8)] = __float2half_rn(v[j] + bs);
                  }
                }
              }
            }
            fused::cbar();
            fused::head_attention<DH>(qs, ks, vt, tag, tpad, T, p.seq, p.ctx + i * DH, H, warp, lane);
          } else if (ph == 1) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (c < nch) {
                const int tk0 = (hf * nch + c) * 8;
                float v[8];
                fused::load_acc8(lane_taddr, 0, tk0, v);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (tk0 + j < T) p.pre1[static_cast<size_t>(tk0 + j) * H + f_out] = v[j] + bias + __half2float(res[c][j]);
              }
          } else if (ph == 2) {
            for (int c = 0; c < nch; ++c) {
              const int tk0 = (hf * nch + c) * 8;
              float v[8];
              fused::load_acc8(lane_taddr, 0, tk0, v);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (tk0 + j < T) {
                  const float x = v[j] + bias;
                  p.act[static_cast<size_t>(tk0 + j) * F + f_out] = __float2half_rn(0.5f * x * (1.0f + erff(x * 0.70710678118654752f)));
                }
              }
            }
          } else {
            // split-K: slab out, the last CTA of the tile sums all slabs in slice order
            const int tile = i / KS, ksl = i % KS;
            float* slab = p.partial + (static_cast<size_t>(tile) * KS) * tpad * 128 + q * 32 + lane;
            if (KS > 1) {
              for (int c = 0; c < nch; ++c) {
                const int tk0 = (hf * nch + c) * 8;
                float v[8];
                fused::load_acc8(lane_taddr, 0, tk0, v);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (tk0 + j < T) slab[(static_cast<size_t>(ksl) * tpad + tk0 + j) * 128] = v[j];
              }
              __threadfence();
              fused::cbar();
              if (tid == 0) {
                const unsigned prev = atomicAdd(p.sem + tile, 1u);
                s_last = prev == static_cast<unsigned>(KS - 1);
                if (s_last) {
                  __threadfence();
                  p.sem[tile] = 0u;  // next use is a layer (four grid barriers) away
                }
              }
              fused::cbar();
            }
            if (KS == 1 || s_last) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (c < nch) {
                  const int tk0 = (hf * nch + c) * 8;
                  float own[8], acc[8];
                  fused::load_acc8(lane_taddr, 0, tk0, own);
#pragma unroll
                  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
                  for (int s = 0; s < KS; ++s) {  // slice order, whoever arrived last: deterministic
                    if (s == ksl) {
#pragma unroll
                      for (int j = 0; j < 8; ++j) acc[j] += own[j];
                    } else {
                      float o[8];
#pragma unroll
                      for (int j = 0; j < 8; ++j) o[j] = tk0 + j < T ? __ldcg(slab + (static_cast<size_t>(s) * tpad + tk0 + j) * 128) : 0.f;
#pragma unroll
                      for (int j = 0; j < 8; ++j) acc[j] += o[j];
                    }
                  }
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    if (tk0 + j < T) p.pre2[static_cast<size_t>(tk0 + j) * H + f_out] = acc[j] + bias + __half2float(res[c][j]);
                }
            }
          }
          ptx::tc_fence_before();  // accumulator reads done before the next job's MMAs (ordered by bready_bar)
          if (tr) tr[bar_idx * 6 + 4] = fused::timer_ns();
        }
        jb += n;
        fused::cbar();  // every compute thread's stores of this phase are issued
        if (tid == 0) fused::grid_arrive(p.bar + bar_idx);
        if (tr) tr[bar_idx * 6 + 5] = fused::timer_ns();
      }
    }
    // ---------------- pooling + L2 normalise (CTA 0)
    if (cta == 0) {
      const int last = 4 * p.num_layers - 1;
      {
        const FusedLayer* Ll = p.layers + p.num_layers - 1;
        for (int j = tid; j < H / 4; j += kFusedComputeThreads) {
          reinterpret_cast<float4*>(lng)[j] = __ldg(reinterpret_cast<const float4*>(Ll->ln2_g) + j);
          reinterpret_cast<float4*>(lnb)[j] = __ldg(reinterpret_cast<const float4*>(Ll->ln2_b) + j);
        }
      }
      if (tid == 0) fused::grid_wait(p.bar + last, static_cast<unsigned>(G));
      fused::cbar();
      fused::stage_ln<false, NV>(p, p.pre2, lng, lnb, bop, true, nullptr, warp, lane);
      fused::cbar();
      const __half* hs = reinterpret_cast<const __half*>(bop);
      constexpr int kPF = (H + kFusedComputeThreads - 1) / kFusedComputeThreads;  // features per thread
      for (int b = 0; b < p.batch; ++b) {
        float cnt = 0.f;
        for (int j = 0; j < p.seq; ++j) cnt += tag[b * p.seq + j] >= 0 ? 1.f : 0.f;
        cnt = fmaxf(cnt, 1e-9f);
        float v[kPF];
        float ss = 0.f;
#pragma unroll
        for (int u = 0; u < kPF; ++u) {
          v[u] = 0.f;
          const int f = tid + u * kFusedComputeThreads;
          if (f < H) {
            if (p.pool_cls) {
              v[u] = __half2float(hs[static_cast<size_t>(b) * p.seq * H + f]);
            } else {
              float acc = 0.f;
              for (int j = 0; j < p.seq; ++j)
                if (tag[b * p.seq + j] >= 0) acc += __half2float(hs[(static_cast<size_t>(b) * p.seq + j) * H + f]);
              v[u] = acc / cnt;
            }
            ss += v[u] * v[u];
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) s_red[warp] = ss;
        fused::cbar();
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < kFusedComputeWarps; ++w) tot += s_red[w];
        const float inv = 1.f / fmaxf(sqrtf(tot), 1e-12f);
#pragma unroll
        for (int u = 0; u < kPF; ++u) {
          const int f = tid + u * kFusedComputeThreads;
          if (f < H) p.out[static_cast<size_t>(b) * H + f] = v[u] * inv;
        }
        fused::cbar();
      }
      // every CTA has arrived at every phase counter: clear them for the next launch
      for (int j = tid; j <= last; j += kFusedComputeThreads) p.bar[j] = 0u;
      if (tr) tr[(last + 1) * 6 + 0] = fused::timer_ns();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kTmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kFusedTmemCols);
  }
}

}  // namespace lxg
