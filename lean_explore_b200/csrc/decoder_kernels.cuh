// Kernels of the Qwen3-class decoder backbone shared by the two models the reference ships with:
// Qwen/Qwen3-Embedding-0.6B (EmbeddingClient, reference src/lean_explore/util/embedding_client.py:58,
// 97-99: last-token pooling, L2 normalise) and Qwen/Qwen3-Reranker-0.6B (RerankerClient,
// src/lean_explore/util/reranker_client.py:110-141: "true"/"false" logits of the last token).
//   tok_emb -> L x [ RMSNorm, QKV GEMM, per-head q/k RMSNorm + RoPE, causal GQA attention,
//   o_proj GEMM (+= residual), RMSNorm, gate|up GEMM + SwiGLU, down GEMM (+= residual) ] -> RMSNorm
// exactly as transformers.Qwen3Model computes it (pre-norm, no biases, rotate_half RoPE with
// position = index in the padded sequence, causal AND key-padding mask).  The residual stream is
// fp32; GEMM operands are fp16, every norm / softmax / accumulation is fp32.  The GEMMs are
// gemm_tc_kernel (encoder_kernels.cuh) with the kEpiStore / kEpiAccF32 / kEpiSwiGLU epilogues.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include "encoder_kernels.cuh"

namespace lxg {

// ------------------------------------------------------------------ RMSNorm
// One warp per token: out = x * rsqrt(mean(x^2) + eps) * w   (fp32 in, fp16 out).  With ids != NULL
// the row is first gathered from the token-embedding table and written to the residual stream
// (the first layer's input_layernorm fused with embed_tokens).
static __global__ void __launch_bounds__(256)
rmsnorm_kernel(float* __restrict__ resid, const int* __restrict__ ids, const __half* __restrict__ tok_emb, int vocab,
               int tokens, int hidden, const float* __restrict__ w, float eps, __half* __restrict__ out) {
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= tokens) return;
  constexpr int kMax = 16;  // hidden <= 1024
  float2 v[kMax];
  const int n2 = hidden >> 1;
  float2* x2 = reinterpret_cast<float2*>(resid + static_cast<size_t>(t) * hidden);
  float ss = 0.f;
  if (ids != nullptr) {
    int id = ids[t];
    id = min(max(id, 0), vocab - 1);
    const __half2* e2 = reinterpret_cast<const __half2*>(tok_emb + static_cast<size_t>(id) * hidden);
#pragma unroll
    for (int i = 0; i < kMax; ++i) {
      const int j = i * 32 + lane;
      v[i] = make_float2(0.f, 0.f);
      if (j < n2) {
        v[i] = __half22float2(e2[j]);
        x2[j] = v[i];
        ss += v[i].x * v[i].x + v[i].y * v[i].y;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < kMax; ++i) {
      const int j = i * 32 + lane;
      v[i] = make_float2(0.f, 0.f);
      if (j < n2) {
        v[i] = x2[j];
        ss += v[i].x * v[i].x + v[i].y * v[i].y;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / hidden + eps);
  __half2* o2 = reinterpret_cast<__half2*>(out + static_cast<size_t>(t) * hidden);
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int j = i * 32 + lane;
    if (j < n2) {
      const float2 ww = reinterpret_cast<const float2*>(w)[j];
      o2[j] = __floats2half2_rn(v[i].x * rstd * ww.x, v[i].y * rstd * ww.y);
    }
  }
}

// Bulk variant (hidden % 128 == 0, no embedding gather): 16-byte accesses - lane l holds features
// [4 (32 i + l), +4) of pass i - same arithmetic per element, the sum of squares in a different
// (fixed) order.  6.8 -> ~5 us per call on 4096 x 1024 (the warp-per-token float2 version moves 3.7 TB/s).
static __global__ void __launch_bounds__(256)
rmsnorm_bulk_kernel(const float* __restrict__ resid, int tokens, int hidden, const float* __restrict__ w, float eps,
                    __half* __restrict__ out) {
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= tokens) return;
  constexpr int kMax = 8;  // hidden <= 1024
  float4 v[kMax];
  const int n4 = hidden >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(resid + static_cast<size_t>(t) * hidden);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int j = i * 32 + lane;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < n4) {
      v[i] = x4[j];
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / hidden + eps);
  uint2* o2 = reinterpret_cast<uint2*>(out + static_cast<size_t>(t) * hidden);
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int j = i * 32 + lane;
    if (j < n4) {
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + j);
      o2[j] = make_uint2(pack_half2(v[i].x * rstd * ww.x, v[i].y * rstd * ww.y), pack_half2(v[i].z * rstd * ww.z, v[i].w * rstd * ww.w));
    }
  }
}

// Skinny path (<= 128 tokens): one CTA per token.  Adds the split-K partial slabs of the preceding
// o_proj / down_proj to the residual row (slab order, so the sum is reproducible), writes the row
// back and normalises it.  All slab loads of a thread are independent and in flight together - a
// warp per token (rmsnorm_kernel) would walk them as one long dependent chain.
static __global__ void __launch_bounds__(256)
rmsnorm_partial_kernel(float* __restrict__ resid, int hidden, const float* __restrict__ w, float eps,
                       __half* __restrict__ out, const float* __restrict__ partial, int nsplit, size_t split_stride) {
  constexpr int kMaxSplit = 16;
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  const int t = blockIdx.x;
  const int n2 = hidden >> 1;
  __shared__ float red[8];
  float2* x2 = reinterpret_cast<float2*>(resid + static_cast<size_t>(t) * hidden);
  const float2* p2 = reinterpret_cast<const float2*>(partial + static_cast<size_t>(t) * hidden);
  float2 v[2];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {  // hidden <= 1024: two float2 per thread
    const int j = i * 256 + threadIdx.x;
    v[i] = make_float2(0.f, 0.f);
    if (j < n2) {
      float2 a[kMaxSplit];
#pragma unroll
      for (int sidx = 0; sidx < kMaxSplit; ++sidx)
        a[sidx] = sidx < nsplit ? p2[sidx * (split_stride >> 1) + j] : make_float2(0.f, 0.f);
      v[i] = x2[j];
#pragma unroll
      for (int sidx = 0; sidx < kMaxSplit; ++sidx) {
        v[i].x += a[sidx].x;
        v[i].y += a[sidx].y;
      }
      x2[j] = v[i];
      ss += v[i].x * v[i].x + v[i].y * v[i].y;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) tot += red[k];
  const float rstd = rsqrtf(tot / hidden + eps);
  __half2* o2 = reinterpret_cast<__half2*>(out + static_cast<size_t>(t) * hidden);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int j = i * 256 + threadIdx.x;
    if (j < n2) {
      const float2 ww = reinterpret_cast<const float2*>(w)[j];
      o2[j] = __floats2half2_rn(v[i].x * rstd * ww.x, v[i].y * rstd * ww.y);
    }
  }
}

// ------------------------------------------------------------------ q/k head RMSNorm + RoPE
// In place on the QKV projection [tokens, (heads + 2 kv_heads) * 128] (fp16): one warp per token,
// looping over its q and k heads (they share the position, so cos / sin are evaluated once per
// token - the kernel was instruction bound on sincosf when every (token, head) had its own warp).
// Lane l holds features 2l, 2l+1, 2l+64, 2l+65 so both rotate_half partners (i, i+64) sit in the
// same lane and every access is a half2.  cos/sin are evaluated in fp32 on angle = pos * inv_freq[i]
// with HF's fp32 inv_freq table (Qwen3RotaryEmbedding).
static __global__ void __launch_bounds__(256)
qk_norm_rope_kernel(__half* __restrict__ qkv, int tokens, int seq, const int* __restrict__ pos_of, int heads, int kv_heads, int hgroup,
                    const float* __restrict__ q_w, const float* __restrict__ k_w,
                    const float* __restrict__ inv_freq, float eps) {
  constexpr int DH = 128;
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  // a warp owns `hgroup` consecutive heads of one token: several for bulk batches (cos / sin once
  // per warp), one for a handful of tokens (more warps, shorter dependent chains)
  const int nh = heads + kv_heads;
  const int groups = (nh + hgroup - 1) / hgroup;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= tokens * groups) return;
  const int t = wid / groups, h_begin = (wid % groups) * hgroup, h_end = min(nh, h_begin + hgroup);
  // packed batches carry the position of every token; padded ones use the column index
  const float pos = static_cast<float>(pos_of != nullptr ? pos_of[t] : t % seq);
  const float2 f = *reinterpret_cast<const float2*>(inv_freq + 2 * lane);
  float s0, c0, s1, c1;
  sincosf(pos * f.x, &s0, &c0);
  sincosf(pos * f.y, &s1, &c1);
  const float2 qlo = *reinterpret_cast<const float2*>(q_w + 2 * lane), qhi = *reinterpret_cast<const float2*>(q_w + 64 + 2 * lane);
  const float2 klo = *reinterpret_cast<const float2*>(k_w + 2 * lane), khi = *reinterpret_cast<const float2*>(k_w + 64 + 2 * lane);
  __half* row = qkv + static_cast<size_t>(t) * (heads + 2 * kv_heads) * DH;
#pragma unroll 4
  for (int h = h_begin; h < h_end; ++h) {
    __half* p = row + h * DH;
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(p + 2 * lane));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(p + 64 + 2 * lane));
    float ss = lo.x * lo.x + lo.y * lo.y + hi.x * hi.x + hi.y * hi.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss / DH + eps);
    const float2 wlo = h < heads ? qlo : klo, whi = h < heads ? qhi : khi;
    const float x0 = lo.x * rstd * wlo.x, x1 = lo.y * rstd * wlo.y, x2 = hi.x * rstd * whi.x, x3 = hi.y * rstd * whi.y;
    *reinterpret_cast<__half2*>(p + 2 * lane) = __floats2half2_rn(x0 * c0 - x2 * s0, x1 * c1 - x3 * s1);
    *reinterpret_cast<__half2*>(p + 64 + 2 * lane) = __floats2half2_rn(x2 * c0 + x0 * s0, x3 * c1 + x1 * s1);
  }
}

// Bulk variant (hundreds of tokens and more): cos / sin come from a table [position][64] of (cos, sin)
// pairs built once per workspace by rope_table_kernel - the same sincosf(pos * inv_freq[i]) values -
// and every access is 16 bytes: 8 lanes cover one head (lane i of the 8 holds features 8i..8i+7 and
// their rotate_half partners 64+8i..64+8i+7), a warp 4 heads per step, `steps` steps per warp.
static __global__ void __launch_bounds__(256) rope_table_kernel(float2* __restrict__ tab, int positions,
                                                                const float* __restrict__ inv_freq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= positions * 64) return;
  float sn, cs;
  sincosf(static_cast<float>(i >> 6) * inv_freq[i & 63], &sn, &cs);
  tab[i] = make_float2(cs, sn);
}

static __global__ void __launch_bounds__(256)
qk_norm_rope_bulk_kernel(__half* __restrict__ qkv, int tokens, int seq, const int* __restrict__ pos_of, int heads, int kv_heads, int steps,
                         const float* __restrict__ q_w, const float* __restrict__ k_w, const float2* __restrict__ tab, float eps) {
  constexpr int DH = 128;
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  const int lane = threadIdx.x & 31, sub = lane >> 3, i8 = (lane & 7) * 8;
  const int nh = heads + kv_heads;
  const int groups = (nh + 4 * steps - 1) / (4 * steps);
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= tokens * groups) return;
  const int t = wid / groups, h0 = (wid % groups) * 4 * steps;
  const int pos = pos_of != nullptr ? pos_of[t] : t % seq;
  float cs[8], sn[8];
  {
    const float4* tp = reinterpret_cast<const float4*>(tab + static_cast<size_t>(pos) * 64 + i8);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 v = __ldg(tp + j);
      cs[2 * j] = v.x;
      sn[2 * j] = v.y;
      cs[2 * j + 1] = v.z;
      sn[2 * j + 1] = v.w;
    }
  }
  __half* row = qkv + static_cast<size_t>(t) * (heads + 2 * kv_heads) * DH;
  for (int st = 0; st < steps; ++st) {
    const int h = h0 + st * 4 + sub;
    const bool live = h < nh;  // (whole 8-lane groups)
    uint4 a = make_uint4(0, 0, 0, 0), b = a;
    __half* p = row + (live ? h : 0) * DH + i8;
    if (live) {
      a = *reinterpret_cast<const uint4*>(p);
      b = *reinterpret_cast<const uint4*>(p + 64);
    }
    float lo[8], hi[8];
    {
      const __half2* a2 = reinterpret_cast<const __half2*>(&a);
      const __half2* b2 = reinterpret_cast<const __half2*>(&b);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 x = __half22float2(a2[j]), y = __half22float2(b2[j]);
        lo[2 * j] = x.x;
        lo[2 * j + 1] = x.y;
        hi[2 * j] = y.x;
        hi[2 * j + 1] = y.y;
      }
    }
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += lo[j] * lo[j] + hi[j] * hi[j];
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    const float rstd = rsqrtf(ss / DH + eps);
    if (live) {
      const float* w = (h < heads ? q_w : k_w) + i8;
      const float4 wl0 = __ldg(reinterpret_cast<const float4*>(w)), wl1 = __ldg(reinterpret_cast<const float4*>(w) + 1);
      const float4 wh0 = __ldg(reinterpret_cast<const float4*>(w + 64)), wh1 = __ldg(reinterpret_cast<const float4*>(w + 64) + 1);
      const float wl[8] = {wl0.x, wl0.y, wl0.z, wl0.w, wl1.x, wl1.y, wl1.z, wl1.w};
      const float wh[8] = {wh0.x, wh0.y, wh0.z, wh0.w, wh1.x, wh1.y, wh1.z, wh1.w};
      uint4 oa, ob;
      __half2* oa2 = reinterpret_cast<__half2*>(&oa);
      __half2* ob2 = reinterpret_cast<__half2*>(&ob);
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float x0 = lo[j] * rstd * wl[j], x1 = lo[j + 1] * rstd * wl[j + 1];
        const float y0 = hi[j] * rstd * wh[j], y1 = hi[j + 1] * rstd * wh[j + 1];
        oa2[j >> 1] = __floats2half2_rn(x0 * cs[j] - y0 * sn[j], x1 * cs[j + 1] - y1 * sn[j + 1]);
        ob2[j >> 1] = __floats2half2_rn(y0 * cs[j] + x0 * sn[j], y1 * cs[j + 1] + x1 * sn[j + 1]);
      }
      *reinterpret_cast<uint4*>(p) = oa;
      *reinterpret_cast<uint4*>(p + 64) = ob;
    }
  }
}

// ------------------------------------------------------------------ causal GQA attention
// Flash-style: one CTA (4 warps) per (64 query rows, q head, sequence); warp w owns 16 query rows;
// three CTAs are resident per SM so one CTA's chunk loads overlap the others' MMAs.  Keys and
// values of the head's KV group are streamed through shared memory in chunks of 64 keys (both
// row-major; the P.V operand is read transposed with ldmatrix.trans), scores and context run on
// mma.sync m16n8k16 with an fp32 online softmax.  Key j is visible to query i iff j <= i and mask[j] != 0 (HF create_causal_mask with a
// padding mask); rows with no visible key (left padding) yield 0 and are never read downstream.
// Packed batches (cu != NULL, no padding tokens at all) give sequence b the tokens [cu[b], cu[b+1]).
constexpr int kCausalRows = 64;
constexpr int kCausalKeys = 64;
// two buffers x (K chunk + V chunk) of 64 rows x (128 + 8) halves, + bias[2][64]
constexpr int kCausalSmem = 4 * kCausalKeys * (128 + 8) * 2 + 2 * kCausalKeys * 4;

template <int DH>
__global__ void __launch_bounds__(kCausalRows * 2, 3)
attention_causal_kernel(const __half* __restrict__ qkv, const int* __restrict__ mask, const int* __restrict__ cu, int seq, int heads,
                        int kv_heads, __half* __restrict__ ctx) {
  constexpr int kKSteps = DH / 16;
  constexpr int kOTiles = DH / 8;
  constexpr int kKPitch = DH + 8;
  constexpr int kTile = kCausalKeys * kKPitch;  // halves per K (or V) chunk
  extern __shared__ __align__(16) uint8_t attn_smem[];  // [2 buffers][K chunk | V chunk] + bias[2][64]
  __half* const kv_sm = reinterpret_cast<__half*>(attn_smem);
  float* const bias_sm = reinterpret_cast<float*>(attn_smem + static_cast<size_t>(4) * kTile * sizeof(__half));
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  // packed batch (cu != NULL): sequence b is tokens [cu[b], cu[b+1]), every key is real
  const int tok0 = cu != nullptr ? cu[b] : b * seq;
  if (cu != nullptr) seq = cu[b + 1] - tok0;
  if (qb * kCausalRows >= seq) return;  // whole CTA, before any barrier
  const int kvh = h / (heads / kv_heads);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const size_t row_stride = static_cast<size_t>(heads + 2 * kv_heads) * DH;
  const __half* base = qkv + static_cast<size_t>(tok0) * row_stride;
  const __half* qbase = base + static_cast<size_t>(h) * DH;
  const __half* kbase = base + static_cast<size_t>(heads + kvh) * DH;
  const __half* vbase = base + static_cast<size_t>(heads + kv_heads + kvh) * DH;
  const int wrow0 = qb * kCausalRows + warp * 16;  // first query row of this warp
  const int r0 = wrow0 + g, r1 = r0 + 8;
  const bool active = wrow0 < seq;

  uint32_t qa[kKSteps][4];
#pragma unroll
  for (int kk = 0; kk < kKSteps; ++kk) {
    const int c = kk * 16 + 2 * t;
    qa[kk][0] = (active && r0 < seq) ? *reinterpret_cast<const uint32_t*>(qbase + r0 * row_stride + c) : 0u;
    qa[kk][1] = (active && r1 < seq) ? *reinterpret_cast<const uint32_t*>(qbase + r1 * row_stride + c) : 0u;
    qa[kk][2] = (active && r0 < seq) ? *reinterpret_cast<const uint32_t*>(qbase + r0 * row_stride + c + 8) : 0u;
    qa[kk][3] = (active && r1 < seq) ? *reinterpret_cast<const uint32_t*>(qbase + r1 * row_stride + c + 8) : 0u;
  }
  float o[kOTiles][4];
#pragma unroll
  for (int n = 0; n < kOTiles; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, l0 = 0.f, l1 = 0.f;
  const float scale = rsqrtf(static_cast<float>(DH)) * 1.4426950408889634f;  // softmax in base 2

  const int key_end = min(seq, (qb + 1) * kCausalRows);  // causal: no key beyond the block's last row
  const int nchunks = (key_end + kCausalKeys - 1) / kCausalKeys;
  const uint32_t smem_kv = ptx::smem_u32(kv_sm);
  // chunk c -> buffer c & 1, staged with cp.async (rows past the sequence are zero filled) one
  // chunk ahead of the MMAs
  auto stage_chunk = [&](int c) {
    const int kb = c * kCausalKeys;
    const uint32_t dst0 = smem_kv + static_cast<uint32_t>((c & 1) * 2 * kTile * 2);
    for (int i = threadIdx.x; i < kCausalKeys * (DH / 8); i += blockDim.x) {
      const int j = i / (DH / 8), cc = i % (DH / 8);
      const bool in = kb + j < seq;
      const size_t row = static_cast<size_t>(in ? kb + j : 0) * row_stride + 8 * cc;
      const uint32_t dst = dst0 + static_cast<uint32_t>((j * kKPitch + 8 * cc) * 2);
      const uint32_t bytes = in ? 16u : 0u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(kbase + row), "r"(bytes) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + kTile * 2), "l"(vbase + row), "r"(bytes) : "memory");
    }
    for (int j = threadIdx.x; j < kCausalKeys; j += blockDim.x)
      bias_sm[(c & 1) * kCausalKeys + j] = (kb + j < seq && (cu != nullptr || mask[tok0 + kb + j] != 0)) ? 0.f : -CUDART_INF_F;
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage_chunk(0);
  for (int c = 0; c < nchunks; ++c) {
    const int kb0 = c * kCausalKeys;
    if (c + 1 < nchunks) {
      stage_chunk(c + 1);  // its buffer was released by the barrier that ended iteration c - 1
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();  // chunk c (every thread's copies + bias) is visible
    const float* bias = bias_sm + (c & 1) * kCausalKeys;
    const uint32_t smem_ks = smem_kv + static_cast<uint32_t>((c & 1) * 2 * kTile * 2);
    const uint32_t smem_vs = smem_ks + kTile * 2;
    if (active && kb0 <= wrow0 + 15) {  // warp-uniform: else the chunk lies entirely in this warp's future

    float sc[8][4];
    // B fragments of Q.K^T with ldmatrix.x4: matrices = 8 keys x dims [32 kk2 + 8 m, +8), m = 0..3
    // -> (b0, b1) of k-steps 2 kk2 and 2 kk2 + 1
    const uint32_t krow = smem_ks + static_cast<uint32_t>(((lane & 7) * kKPitch + (lane >> 3) * 8) * 2);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
      for (int kk2 = 0; kk2 < kKSteps / 2; ++kk2) {
        uint32_t b0, b1, b2, b3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(krow + static_cast<uint32_t>((j * 8 * kKPitch + kk2 * 32) * 2)));
        mma_m16n8k16(sc[j], qa[2 * kk2], b0, b1);
        mma_m16n8k16(sc[j], qa[2 * kk2 + 1], b2, b3);
      }
    }
    float bm0 = -CUDART_INF_F, bm1 = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key0 = kb0 + j * 8 + 2 * t;
      const float b0 = bias[j * 8 + 2 * t], b1 = bias[j * 8 + 2 * t + 1];
      sc[j][0] = key0 <= r0 ? sc[j][0] * scale + b0 : -CUDART_INF_F;
      sc[j][1] = key0 + 1 <= r0 ? sc[j][1] * scale + b1 : -CUDART_INF_F;
      sc[j][2] = key0 <= r1 ? sc[j][2] * scale + b0 : -CUDART_INF_F;
      sc[j][3] = key0 + 1 <= r1 ? sc[j][3] * scale + b1 : -CUDART_INF_F;
      bm0 = fmaxf(bm0, fmaxf(sc[j][0], sc[j][1]));
      bm1 = fmaxf(bm1, fmaxf(sc[j][2], sc[j][3]));
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);
    const float mu0 = mn0 == -CUDART_INF_F ? 0.f : mn0, mu1 = mn1 == -CUDART_INF_F ? 0.f : mn1;
    const float corr0 = exp2f(m0 - mu0), corr1 = exp2f(m1 - mu1);
    m0 = mn0;
    m1 = mn1;
    float s0 = 0.f, s1 = 0.f;
    uint32_t pa[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f(sc[j][0] - mu0), p1 = exp2f(sc[j][1] - mu0);
      const float p2 = exp2f(sc[j][2] - mu1), p3 = exp2f(sc[j][3] - mu1);
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
    // context += P . V: B fragments of two feature tiles per ldmatrix.x4.trans
    // (lanes 0-15: keys kk*16 + lane of tile n, lanes 16-31: the same keys of tile n + 1)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t vrow = smem_vs + static_cast<uint32_t>(((kk * 16 + (lane & 15)) * kKPitch + (lane >> 4) * 8) * 2);
#pragma unroll
      for (int n = 0; n < kOTiles; n += 2) {
        uint32_t b0, b1, b2, b3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(vrow + n * 16));
        mma_m16n8k16(o[n], pa[kk], b0, b1);
        mma_m16n8k16(o[n + 1], pa[kk], b2, b3);
      }
    }
    }  // active
    __syncthreads();  // buffer c & 1 may be overwritten by the copies of chunk c + 2
  }
  if (!active) return;
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
  const size_t ctx_stride = static_cast<size_t>(heads) * DH;
  __half* out0 = ctx + (static_cast<size_t>(tok0) + r0) * ctx_stride + h * DH + 2 * t;
  __half* out1 = ctx + (static_cast<size_t>(tok0) + r1) * ctx_stride + h * DH + 2 * t;
#pragma unroll
  for (int n = 0; n < kOTiles; ++n) {
    if (r0 < seq) *reinterpret_cast<__half2*>(out0 + n * 8) = __floats2half2_rn(o[n][0] * inv0, o[n][1] * inv0);
    if (r1 < seq) *reinterpret_cast<__half2*>(out1 + n * 8) = __floats2half2_rn(o[n][2] * inv1, o[n][3] * inv1);
  }
}

// ------------------------------------------------------------------ heads on the last token
// One CTA per sequence.  The last token is the last position whose mask is set (position S-1
// under the left padding both reference clients use; the mask-based index also covers right
// padding, as sentence-transformers' Pooling(lasttoken) does).  Final RMSNorm in fp32, then
//   mode 0 (embedding): out[b, :] = x / max(||x||, 1e-12)       (Pooling(lasttoken) + Normalize)
//   mode 1 (reranker) : out[b] = softmax([false, true] logits)[1] with logits = x . lm_head[token]
//                       (reranker_client.py:127-139)
static __global__ void __launch_bounds__(256)
last_token_head_kernel(const float* __restrict__ resid, const int* __restrict__ mask, const int* __restrict__ cu, int seq, int hidden,
                       const float* __restrict__ norm_w, float eps, int mode, const __half* __restrict__ lm_head,
                       int token_true, int token_false, float* __restrict__ out,
                       const float* __restrict__ partial, int nsplit, size_t split_stride) {
  ptx::pdl_wait();
  const int b = blockIdx.x;
  __shared__ float red[3][8];
  __shared__ int s_last;
  extern __shared__ float xs[];
  if (threadIdx.x == 0) {
    if (cu != nullptr) {
      s_last = cu[b + 1] - 1;  // packed: absolute index of the sequence's last token
    } else {
      int last = seq - 1;
      while (last > 0 && mask[b * seq + last] == 0) --last;
      s_last = b * seq + last;
    }
  }
  __syncthreads();
  const float* x = resid + static_cast<size_t>(s_last) * hidden;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  auto block_sum = [&](float v, int slot) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[slot][warp] = v;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < nwarps; ++w) tot += red[slot][w];
    __syncthreads();
    return tot;
  };
  float ss = 0.f;
  for (int c = threadIdx.x; c < hidden; c += blockDim.x) {
    float v = x[c];
    for (int sidx = 0; sidx < nsplit; ++sidx)  // split-K partials of the last down_proj
      v += partial[static_cast<size_t>(sidx) * split_stride + static_cast<size_t>(s_last) * hidden + c];
    xs[c] = v;
    ss += v * v;
  }
  const float rstd = rsqrtf(block_sum(ss, 0) / hidden + eps);
  float n2 = 0.f, dt = 0.f, df = 0.f;
  for (int c = threadIdx.x; c < hidden; c += blockDim.x) {
    const float v = xs[c] * rstd * norm_w[c];
    xs[c] = v;
    n2 += v * v;
    if (mode == 1) {
      dt += v * __half2float(lm_head[static_cast<size_t>(token_true) * hidden + c]);
      df += v * __half2float(lm_head[static_cast<size_t>(token_false) * hidden + c]);
    }
  }
  if (mode == 0) {
    const float inv = 1.f / fmaxf(sqrtf(block_sum(n2, 0)), 1e-12f);
    for (int c = threadIdx.x; c < hidden; c += blockDim.x) out[static_cast<size_t>(b) * hidden + c] = xs[c] * inv;
  } else {
    const float lt = block_sum(dt, 1), lf = block_sum(df, 2);
    if (threadIdx.x == 0) out[b] = 1.f / (1.f + expf(lf - lt));
  }
}

}  // namespace lxg
