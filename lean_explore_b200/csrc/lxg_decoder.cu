// Qwen3-class decoder entry points of the C ABI (include/lxg.h): host logic only - workspace,
// tensor maps, launch sequence, CUDA-graph replay.  lxg_decoder_embed replaces
// SentenceTransformer.encode for the shipped Qwen/Qwen3-Embedding-0.6B (reference
// src/lean_explore/util/embedding_client.py:58,88-101), lxg_decoder_rerank replaces the
// AutoModelForCausalLM forward + true/false softmax of RerankerClient._compute_scores_sync
// (src/lean_explore/util/reranker_client.py:110-141).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lxg.h"
#include "common.h"
#include "attention_tc.cuh"
#include "decoder_kernels.cuh"
#include "gemm_host.cuh"

using namespace lxg;

struct lxg_decoder {
  int device = 0;  // the GPU that holds the weights; made current by every entry point
  lxg_qwen3_weights w{};
  std::vector<lxg_qwen3_layer> layers;
  std::mutex mu;
  int cap_tokens = 0;
  float* resid = nullptr;                                          // fp32 residual stream [cap, H]
  float* partial = nullptr;                                        // skinny path: split-K slabs [kSkinnySplits][128, H]
  __half *hn = nullptr, *qkv = nullptr, *ctx = nullptr, *act = nullptr;  // normed rows, QKV, context, SwiGLU output
  int *ids = nullptr, *mask = nullptr;
  int *pos = nullptr, *cu = nullptr;   // packed batches: position of every token, sequence offsets [b + 1]
  float2* rope_tab = nullptr;          // (cos, sin) [positions][64] for the bulk RoPE kernel
  int rope_positions = 0;
  int cu_cap = 0;
  int* stage = nullptr;                // pinned host staging: packed ids | pos | cu
  size_t stage_cap = 0;                // ints
  bool pack = true;                    // LXG_DECODER_PACK=0: always compute the padded rectangle
  bool pdl = true;                     // LXG_PDL=0: plain stream-ordered launches (no programmatic dependent launch)
  bool attn_tc = true;                 // LXG_ATTN_TC=0: mma.sync attention for every sequence length (A/B measurements)
  CUtensorMap map_hn{}, map_ctx{}, map_act{}, map_qkv{};
  CUtensorMap* map_resid_dev = nullptr;               // fp32 residual stream as 32 x 32 boxes (device copy of the map): TMA reduce epilogue of o_proj / down_proj
  CUtensorMap map_hn32{}, map_ctx32{}, map_act32{};  // 32-row boxes: the query path's GEMMs (gemm_host.cuh)
  int trace_layer = -1;                               // LXG_GEMM_TRACE=<layer>: device timeline of that layer's pair GEMMs -> stderr
  unsigned long long* trace = nullptr;                // [4 GEMMs][296 CTAs][16]
  bool rope_bulk = true;                              // LXG_ROPE_BULK=0: per-warp sincosf RoPE kernel for every batch size
  bool query_gemm = true;                             // LXG_QUERY_GEMM=0: 128 x 128 tiles for every row count
  std::vector<CUtensorMap> map_wqkv, map_wo, map_wgu, map_wdown;
  std::vector<CUtensorMap> map_wo64, map_wdown64;  // 64-row boxes: 256 x 128 pair tiles (gemm_host.cuh)
  std::vector<CUtensorMap> map_wqkv32, map_wgu64;  // query path: 32- / 64-column output tiles
  int launches = 0;
  int last_tokens = 0;  // tokens the last forward actually computed (after packing)
  struct Graph {
    int b, s, mode, tt, tf;
    cudaGraphExec_t exec;
  };
  std::vector<Graph> graphs;
  bool use_graphs = true;
  float* out_buf = nullptr;
  size_t out_cap = 0;  // floats
  cudaStream_t own = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
};

namespace {

constexpr int kHeadDim = 128;
constexpr int kSkinnyRows = 128;    // up to one row tile the o_proj / down_proj GEMMs are split along K ...
constexpr size_t kRopeTablePositions = 16384;
constexpr int kSkinnySplits = 16;   // ... into this many ranges (8 output tiles x 16 = 128 CTAs instead of 8)

void drop_graphs(lxg_decoder* e) {
  for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
}

void free_ws(lxg_decoder* e) {
  drop_graphs(e);
  cudaFree(e->resid);
  cudaFree(e->partial);
  e->partial = nullptr;
  cudaFree(e->hn);
  cudaFree(e->qkv);
  cudaFree(e->rope_tab);
  e->rope_tab = nullptr;
  e->rope_positions = 0;
  cudaFree(e->ctx);
  cudaFree(e->act);
  cudaFree(e->ids);
  cudaFree(e->mask);
  cudaFree(e->pos);
  e->pos = nullptr;
  e->resid = nullptr;
  e->hn = e->qkv = e->ctx = e->act = nullptr;
  e->ids = e->mask = nullptr;
  e->cap_tokens = 0;
}

int reserve_ws(lxg_decoder* e, int tokens) {
  if (tokens <= e->cap_tokens) return LXG_OK;
  free_ws(e);
  const size_t cap = (static_cast<size_t>(std::max(tokens, 256)) + 127) / 128 * 128;
  const size_t H = e->w.hidden, F = e->w.ffn;
  const size_t QKV = static_cast<size_t>(e->w.heads + 2 * e->w.kv_heads) * kHeadDim, C = static_cast<size_t>(e->w.heads) * kHeadDim;
  LXG_CUDA(cudaMalloc(&e->resid, cap * H * sizeof(float)));
  LXG_CUDA(cudaMalloc(&e->partial, static_cast<size_t>(kSkinnySplits) * kSkinnyRows * H * sizeof(float)));
  LXG_CUDA(cudaMalloc(&e->hn, cap * H * sizeof(__half)));
  LXG_CUDA(cudaMalloc(&e->qkv, cap * QKV * sizeof(__half)));
  LXG_CUDA(cudaMalloc(&e->ctx, cap * C * sizeof(__half)));
  LXG_CUDA(cudaMalloc(&e->act, cap * F * sizeof(__half)));
  LXG_CUDA(cudaMalloc(&e->ids, cap * sizeof(int)));
  LXG_CUDA(cudaMalloc(&e->mask, cap * sizeof(int)));
  LXG_CUDA(cudaMalloc(&e->pos, cap * sizeof(int)));
  // no position exceeds the token capacity; beyond kRopeTablePositions the per-warp sincosf kernel serves
  e->rope_positions = static_cast<int>(std::min<size_t>(cap, kRopeTablePositions));
  LXG_CUDA(cudaMalloc(&e->rope_tab, static_cast<size_t>(e->rope_positions) * 64 * sizeof(float2)));
  rope_table_kernel<<<(e->rope_positions * 64 + 255) / 256, 256>>>(e->rope_tab, e->rope_positions,
                                                                  reinterpret_cast<const float*>(e->w.inv_freq));
  LXG_CUDA(cudaGetLastError());
  // rows beyond the live tokens are read by TMA (never stored): keep them finite
  LXG_CUDA(cudaMemset(e->hn, 0, cap * H * sizeof(__half)));
  LXG_CUDA(cudaMemset(e->qkv, 0, cap * QKV * sizeof(__half)));
  LXG_CUDA(cudaMemset(e->ctx, 0, cap * C * sizeof(__half)));
  LXG_CUDA(cudaMemset(e->act, 0, cap * F * sizeof(__half)));
  // the table kernel and the memsets ran on the legacy stream; forwards run on non-blocking ones
  LXG_CUDA(cudaDeviceSynchronize());
  int rc;
  if ((rc = make_map(&e->map_hn, e->hn, static_cast<int>(cap), static_cast<int>(H))) != LXG_OK) return rc;
  if ((rc = make_map(&e->map_ctx, e->ctx, static_cast<int>(cap), static_cast<int>(C))) != LXG_OK) return rc;
  if ((rc = make_map(&e->map_act, e->act, static_cast<int>(cap), static_cast<int>(F))) != LXG_OK) return rc;
  if ((rc = make_map(&e->map_qkv, e->qkv, static_cast<int>(cap), static_cast<int>(QKV), kTcAttnRows)) != LXG_OK) return rc;
  {
    CUtensorMap m;
    if ((rc = make_map_f32_acc(&m, e->resid, static_cast<int>(cap), static_cast<int>(H))) != LXG_OK) return rc;
    if (!e->map_resid_dev) LXG_CUDA(cudaMalloc(&e->map_resid_dev, sizeof(CUtensorMap)));
    LXG_CUDA(cudaMemcpy(e->map_resid_dev, &m, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  }
  if ((rc = make_map(&e->map_hn32, e->hn, static_cast<int>(cap), static_cast<int>(H), 32)) != LXG_OK) return rc;
  if ((rc = make_map(&e->map_ctx32, e->ctx, static_cast<int>(cap), static_cast<int>(C), 32)) != LXG_OK) return rc;
  if ((rc = make_map(&e->map_act32, e->act, static_cast<int>(cap), static_cast<int>(F), 32)) != LXG_OK) return rc;
  e->cap_tokens = static_cast<int>(cap);
  return LXG_OK;
}

// 1 + 8 * layers + 1 launches on `st`, workspace pointers only (graph-capturable).
// Padded: tokens = b * s, `s` columns per sequence, key mask from e->mask.  Packed: `tokens` real
// tokens back to back, sequence i = [cu[i], cu[i+1]), `s` = the longest sequence.
int launch_forward(lxg_decoder* e, int b, int s, int tokens, bool packed, int mode, int tt, int tf, cudaStream_t st) {
  const int* pos_of = packed ? e->pos : nullptr;
  const int* cu = packed ? e->cu : nullptr;
  const int H = e->w.hidden, F = e->w.ffn, heads = e->w.heads, kvh = e->w.kv_heads;
  const int QKV = (heads + 2 * kvh) * kHeadDim, C = heads * kHeadDim;
  const float eps = e->w.rms_eps;
  int launches = 0;
  const int row_blocks = (tokens + 7) / 8;
  // skinny path (one row tile): o_proj / down_proj have only H / 128 output tiles, so K is split
  // and the partial slabs are added to the residual stream by the kernel that reads it next
  const bool skinny = tokens <= kSkinnyRows;
  // at most 32 rows (a query): narrow output tiles so that every SM takes part in the weight stream
  const bool query = e->query_gemm && tokens <= 32;
  const size_t slab = static_cast<size_t>(kSkinnyRows) * H;
  const bool pdl = e->pdl;
  int pending = 0;  // slabs waiting to be absorbed by the next RMSNorm / the head kernel
  const int hgroup = tokens >= 2048 ? 4 : 1;  // heads per RoPE warp (cos / sin are evaluated once per warp)
  const int rope_blocks = (tokens * ((heads + kvh + hgroup - 1) / hgroup) + 7) / 8;
  // bulk batches: table-driven, 16-byte accesses (decoder_kernels.cuh); positions are < s
  constexpr int kRopeBulkSteps = 2;  // 8 heads per warp
  const bool norm_bulk = e->rope_bulk && tokens >= 256 && H % 128 == 0;  // 16-byte RMSNorm kernel (same switch as the bulk RoPE)
  const bool rope_bulk = e->rope_bulk && tokens >= 256 && s <= e->rope_positions;
  const int rope_bulk_blocks = (tokens * ((heads + kvh + 4 * kRopeBulkSteps - 1) / (4 * kRopeBulkSteps)) + 7) / 8;
  // sequences of more than one 64-row tile: tcgen05 attention on 128-row tiles (attention_tc.cuh);
  // shorter ones (a query) stay on the mma.sync kernel, whose tile they do not even fill
  const bool attn_tc = e->attn_tc && s > kCausalRows;
  const dim3 attn_grid(attn_tc ? (s + kTcAttnRows - 1) / kTcAttnRows : (s + kCausalRows - 1) / kCausalRows, heads, b);
  if (attn_tc) LXG_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(&attention_causal_tc_kernel), kTcAttnSmem));
  else LXG_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(&attention_causal_kernel<kHeadDim>), kCausalSmem));
  for (int l = 0; l < e->w.layers; ++l) {
    const lxg_qwen3_layer& L = e->layers[l];
    // input_layernorm (layer 0: fused with the embed_tokens gather)
    // every kernel but the first (it follows the input copies) is a programmatic dependent launch:
    // its prologue overlaps the previous kernel's tail
    if (pending > 0)
      LXG_CUDA(lxg_launch(rmsnorm_partial_kernel, dim3(tokens), dim3(256), 0, st, pdl, e->resid, H, reinterpret_cast<const float*>(L.ln1), eps,
                          e->hn, static_cast<const float*>(e->partial), pending, slab));
    else if (norm_bulk && l > 0)
      LXG_CUDA(lxg_launch(rmsnorm_bulk_kernel, dim3(row_blocks), dim3(256), 0, st, pdl, static_cast<const float*>(e->resid), tokens, H,
                          reinterpret_cast<const float*>(L.ln1), eps, e->hn));
    else
      LXG_CUDA(lxg_launch(rmsnorm_kernel, dim3(row_blocks), dim3(256), 0, st, pdl && l > 0, e->resid, static_cast<const int*>(l == 0 ? e->ids : nullptr),
                          reinterpret_cast<const __half*>(e->w.tok_emb), e->w.vocab, tokens, H, reinterpret_cast<const float*>(L.ln1), eps, e->hn));
    pending = 0;
    GemmParams gp{};
    unsigned long long* const tr = (l == e->trace_layer && e->trace) ? e->trace : nullptr;
    gp.bias = nullptr;
    gp.residual = nullptr;
    gp.m = tokens;
    // q_proj | k_proj | v_proj
    gp.out = e->qkv;
    gp.n = QKV;
    gp.k = H;
    gp.trace = tr;
    if (query) LXG_CUDA((launch_gemm_query<kEpiStore, 32>(e->map_hn32, e->map_wqkv32[l], gp, st, pdl)));
    else LXG_CUDA(launch_gemm<kEpiStore>(e->map_hn, e->map_wqkv[l], gp, st, pdl));
    if (rope_bulk)
      LXG_CUDA(lxg_launch(qk_norm_rope_bulk_kernel, dim3(rope_bulk_blocks), dim3(256), 0, st, pdl, e->qkv, tokens, s, pos_of, heads, kvh,
                          kRopeBulkSteps, reinterpret_cast<const float*>(L.q_norm), reinterpret_cast<const float*>(L.k_norm),
                          static_cast<const float2*>(e->rope_tab), eps));
    else
      LXG_CUDA(lxg_launch(qk_norm_rope_kernel, dim3(rope_blocks), dim3(256), 0, st, pdl, e->qkv, tokens, s, pos_of, heads, kvh, hgroup,
                          reinterpret_cast<const float*>(L.q_norm), reinterpret_cast<const float*>(L.k_norm),
                          reinterpret_cast<const float*>(e->w.inv_freq), eps));
    if (attn_tc)
      LXG_CUDA(lxg_launch(attention_causal_tc_kernel, attn_grid, dim3(kTcAttnThreads), kTcAttnSmem, st, pdl, e->map_qkv,
                          static_cast<const int*>(e->mask), cu, s, heads, kvh, e->ctx));
    else
      LXG_CUDA(lxg_launch(attention_causal_kernel<kHeadDim>, attn_grid, dim3(kCausalRows * 2), kCausalSmem, st, pdl, static_cast<const __half*>(e->qkv),
                          static_cast<const int*>(e->mask), cu, s, heads, kvh, e->ctx));
    // o_proj, accumulated onto the residual stream
    gp.n = H;
    gp.k = C;
    gp.trace = tr ? tr + 296 * 16 : nullptr;
    if (skinny) {
      gp.out = e->partial;
      gp.ksplit = std::min(kSkinnySplits, C / kGemmBK);
      gp.split_stride = slab;
      if (query) LXG_CUDA((launch_gemm_query<kEpiPartial, 128>(e->map_ctx32, e->map_wo[l], gp, st, pdl)));
      else LXG_CUDA(launch_gemm<kEpiPartial>(e->map_ctx, e->map_wo[l], gp, st, pdl));
      pending = gp.ksplit;
      gp.ksplit = 0;
    } else {
      gp.out = e->resid;
      LXG_CUDA(launch_gemm<kEpiAccF32>(e->map_ctx, e->map_wo[l], gp, st, pdl, &e->map_wo64[l], e->map_resid_dev));
    }
    if (pending > 0)
      LXG_CUDA(lxg_launch(rmsnorm_partial_kernel, dim3(tokens), dim3(256), 0, st, pdl, e->resid, H, reinterpret_cast<const float*>(L.ln2), eps,
                          e->hn, static_cast<const float*>(e->partial), pending, slab));
    else if (norm_bulk)
      LXG_CUDA(lxg_launch(rmsnorm_bulk_kernel, dim3(row_blocks), dim3(256), 0, st, pdl, static_cast<const float*>(e->resid), tokens, H,
                          reinterpret_cast<const float*>(L.ln2), eps, e->hn));
    else
      LXG_CUDA(lxg_launch(rmsnorm_kernel, dim3(row_blocks), dim3(256), 0, st, pdl, e->resid, static_cast<const int*>(nullptr),
                          static_cast<const __half*>(nullptr), 0, tokens, H, reinterpret_cast<const float*>(L.ln2), eps, e->hn));
    pending = 0;
    // gate_proj | up_proj (interleaved) + SwiGLU
    gp.out = e->act;
    gp.n = 2 * F;
    gp.k = H;
    gp.trace = tr ? tr + 2 * 296 * 16 : nullptr;
    if (query) LXG_CUDA((launch_gemm_query<kEpiSwiGLU, 64>(e->map_hn32, e->map_wgu64[l], gp, st, pdl)));
    else LXG_CUDA(launch_gemm<kEpiSwiGLU>(e->map_hn, e->map_wgu[l], gp, st, pdl));
    // down_proj, accumulated onto the residual stream
    gp.n = H;
    gp.k = F;
    gp.trace = tr ? tr + 3 * 296 * 16 : nullptr;
    if (skinny) {
      gp.out = e->partial;
      gp.ksplit = std::min(kSkinnySplits, F / kGemmBK);
      gp.split_stride = slab;
      if (query) LXG_CUDA((launch_gemm_query<kEpiPartial, 128>(e->map_act32, e->map_wdown[l], gp, st, pdl)));
      else LXG_CUDA(launch_gemm<kEpiPartial>(e->map_act, e->map_wdown[l], gp, st, pdl));
      pending = gp.ksplit;
      gp.ksplit = 0;
    } else {
      gp.out = e->resid;
      LXG_CUDA(launch_gemm<kEpiAccF32>(e->map_act, e->map_wdown[l], gp, st, pdl, &e->map_wdown64[l], e->map_resid_dev));
    }
    launches += 8;
  }
  LXG_CUDA(lxg_launch(last_token_head_kernel, dim3(b), dim3(256), H * sizeof(float), st, pdl, static_cast<const float*>(e->resid),
                      static_cast<const int*>(e->mask), cu, s, H, reinterpret_cast<const float*>(e->w.final_norm), eps, mode,
                      reinterpret_cast<const __half*>(e->w.lm_head), tt, tf, e->out_buf, static_cast<const float*>(e->partial), pending, slab));
  ++launches;
  e->launches = launches;
  return LXG_OK;
}

int run(lxg_decoder* e, const int32_t* ids, const int32_t* mask, int32_t b, int32_t s, int mode, int tt, int tf, float* out,
        void* stream) {
  if (!e || !ids || !mask || !out) return set_error(LXG_EINVAL, "NULL argument");
  if (b < 0 || s <= 0) return set_error(LXG_EINVAL, "b must be >= 0 and s >= 1");
  if (b > 65535) return set_error(LXG_EUNSUPPORTED, "more than 65535 sequences per call");
  if (mode == 1) {
    if (!e->w.lm_head) return set_error(LXG_EINVAL, "this decoder was created without lm_head weights");
    if (tt < 0 || tt >= e->w.vocab || tf < 0 || tf >= e->w.vocab) return set_error(LXG_EINVAL, "true/false token id outside the vocabulary");
  }
  if (b == 0) return LXG_OK;
  DeviceGuard guard(e->device);
  NvtxRange nvtx("lxg_decoder forward (embed / rerank)");
  std::lock_guard<std::mutex> lock(e->mu);
  cudaStream_t caller = reinterpret_cast<cudaStream_t>(stream);
  const long long tokens_ll = static_cast<long long>(b) * s;
  if (tokens_ll > (1 << 20)) return set_error(LXG_EUNSUPPORTED, "more than 1M tokens per call");
  const int tokens = static_cast<int>(tokens_ll);
  int rc = reserve_ws(e, tokens);
  if (rc != LXG_OK) return rc;
  if (!e->own) {
    LXG_CUDA(cudaStreamCreateWithFlags(&e->own, cudaStreamNonBlocking));
    LXG_CUDA(cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming));
    LXG_CUDA(cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming));
  }
  cudaStream_t st = e->own;
  LXG_CUDA(cudaEventRecord(e->ev_in, caller));
  LXG_CUDA(cudaStreamWaitEvent(st, e->ev_in, 0));
  const int H = e->w.hidden;
  const bool ids_dev = is_device_ptr(ids), mask_dev = is_device_ptr(mask), out_dev = is_device_ptr(out);
  // ---- packing: with host inputs and more than one sequence, padding tokens are dropped on the
  // host (every mask row must be one contiguous run of ones - what a tokenizer's left / right
  // padding produces) and only the real tokens go through the layers.  RoPE is relative and pad
  // keys are masked in the padded formulation, so both compute the same function.
  int live_tokens = tokens, max_len = s;
  bool packed = false;
  if (e->pack && b > 1 && !ids_dev && !mask_dev) {
    const size_t need = static_cast<size_t>(2) * tokens + b + 1;
    if (need > e->stage_cap) {
      if (e->stage) cudaFreeHost(e->stage);
      e->stage = nullptr;
      e->stage_cap = 0;
      LXG_CUDA(cudaMallocHost(&e->stage, need * sizeof(int)));
      e->stage_cap = need;
    }
    if (b + 1 > e->cu_cap) {
      cudaFree(e->cu);
      e->cu = nullptr;
      e->cu_cap = 0;
      LXG_CUDA(cudaMalloc(&e->cu, static_cast<size_t>(b + 1) * sizeof(int)));
      e->cu_cap = b + 1;
    }
    // the previous call's H2D copies from the staging buffer must have drained
    LXG_CUDA(cudaStreamSynchronize(st));
    int* pid = e->stage;
    int* ppos = e->stage + tokens;
    int* pcu = e->stage + 2 * static_cast<size_t>(tokens);
    int n = 0, longest = 1;
    bool ok = true;
    for (int i = 0; i < b && ok; ++i) {
      const int32_t* mrow = mask + static_cast<size_t>(i) * s;
      const int32_t* irow = ids + static_cast<size_t>(i) * s;
      int first = 0;
      while (first < s && mrow[first] == 0) ++first;
      int last = s - 1;
      while (last >= first && mrow[last] == 0) --last;
      for (int j = first; j <= last; ++j) ok = ok && mrow[j] != 0;
      pcu[i] = n;
      if (last < first) {  // no token at all: keep one (padding) token so the row stays defined
        pid[n] = irow[s - 1];
        ppos[n] = 0;
        ++n;
        continue;
      }
      for (int j = first; j <= last; ++j) {
        pid[n] = irow[j];
        ppos[n] = j - first;
        ++n;
      }
      longest = std::max(longest, last - first + 1);
    }
    pcu[b] = n;
    // worth it only for a bulk batch that is at least 10 % padding: a packed forward is launched
    // kernel by kernel, small rectangular batches are better served by the captured graph
    if (ok && tokens >= 512 && static_cast<long long>(n) * 10 <= static_cast<long long>(tokens) * 9) {
      packed = true;
      live_tokens = n;
      max_len = longest;
      LXG_CUDA(cudaMemcpyAsync(e->ids, pid, n * sizeof(int), cudaMemcpyHostToDevice, st));
      LXG_CUDA(cudaMemcpyAsync(e->pos, ppos, n * sizeof(int), cudaMemcpyHostToDevice, st));
      LXG_CUDA(cudaMemcpyAsync(e->cu, pcu, (b + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    }
  }
  if (!packed) {
    LXG_CUDA(cudaMemcpyAsync(e->ids, ids, tokens * sizeof(int), ids_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    LXG_CUDA(cudaMemcpyAsync(e->mask, mask, tokens * sizeof(int), mask_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  }
  e->last_tokens = live_tokens;
  const size_t out_floats = mode == 0 ? static_cast<size_t>(b) * H : static_cast<size_t>(b);
  if (out_floats > e->out_cap) {
    drop_graphs(e);
    cudaFree(e->out_buf);
    e->out_buf = nullptr;
    e->out_cap = 0;
    const size_t cap = std::max(out_floats, static_cast<size_t>(64) * H);
    LXG_CUDA(cudaMalloc(&e->out_buf, cap * sizeof(float)));
    e->out_cap = cap;
  }
  bool ran = false;
  if (packed) {
    // ragged shapes rarely repeat and a bulk forward is not launch bound: plain launches
    rc = launch_forward(e, b, max_len, live_tokens, true, mode, tt, tf, st);
    if (rc != LXG_OK) return rc;
    ran = true;
  } else if (e->use_graphs) {
    cudaGraphExec_t exec = nullptr;
    for (auto& g : e->graphs)
      if (g.b == b && g.s == s && g.mode == mode && g.tt == tt && g.tf == tf) exec = g.exec;
    if (exec) {
      LXG_CUDA(cudaGraphLaunch(exec, st));
      ran = true;
    } else {
      // first call with this shape: run eagerly (sets kernel attributes), then capture for later calls
      rc = launch_forward(e, b, s, tokens, false, mode, tt, tf, st);
      if (rc != LXG_OK) return rc;
      ran = true;
      cudaGraph_t graph = nullptr;
      if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        const int rc2 = launch_forward(e, b, s, tokens, false, mode, tt, tf, st);
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc2 == LXG_OK && ce == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
          if (e->graphs.size() >= 64) drop_graphs(e);
          e->graphs.push_back({b, s, mode, tt, tf, exec});
        } else {
          e->use_graphs = false;
        }
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
      } else {
        cudaGetLastError();
        e->use_graphs = false;
      }
    }
  }
  if (!ran) {
    rc = launch_forward(e, b, s, tokens, false, mode, tt, tf, st);
    if (rc != LXG_OK) return rc;
  }
  if (e->trace) {  // diagnostics: timeline of the traced layer's GEMMs, CTAs 0 / 72 / 146 (even CTA of a pair), us since the CTA started
    std::vector<unsigned long long> h(4 * 296 * 16);
    LXG_CUDA(cudaStreamSynchronize(st));
    LXG_CUDA(cudaMemcpy(h.data(), e->trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    static const char* names[4] = {"qkv", "o_proj", "gate|up", "down"};
    static const char* slots[13] = {"start", "prologue", "dependency", "-", "-", "-", "acc t0", "acc t1", "acc t2",
                                    "drain t0", "drain t1", "drain t2", "end"};
    for (int g = 0; g < 4; ++g) {
      unsigned long long first = ~0ull, last = 0;
      for (int c = 0; c < 296; ++c) {
        const unsigned long long* t = &h[(static_cast<size_t>(g) * 296 + c) * 16];
        if (t[0]) first = std::min(first, t[0]);
        last = std::max(last, t[12]);
      }
      if (last == 0) continue;
      std::fprintf(stderr, "[lxg] %s GEMM of layer %d: %.2f us from the first CTA's start to the last CTA's end\n", names[g], e->trace_layer,
                   (last - first) * 1e-3);
      for (int c : {0, 72, 146}) {
        const unsigned long long* t = &h[(static_cast<size_t>(g) * 296 + c) * 16];
        if (!t[0]) continue;
        std::fprintf(stderr, "[lxg]   CTA %3d (+%.2f us):", c, (t[0] - first) * 1e-3);
        for (int k = 1; k < 13; ++k)
          if (t[k]) std::fprintf(stderr, " %s %.2f", slots[k], (t[k] - t[0]) * 1e-3);
        std::fprintf(stderr, "\n");
      }
    }
    LXG_CUDA(cudaMemset(e->trace, 0, h.size() * sizeof(unsigned long long)));
  }
  LXG_CUDA(cudaMemcpyAsync(out, e->out_buf, out_floats * sizeof(float), out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  if (!out_dev) {
    LXG_CUDA(cudaStreamSynchronize(st));
  } else {
    LXG_CUDA(cudaEventRecord(e->ev_out, st));
    LXG_CUDA(cudaStreamWaitEvent(caller, e->ev_out, 0));
  }
  return LXG_OK;
}

}  // namespace

extern "C" {

int lxg_decoder_create(lxg_decoder** out, const lxg_qwen3_weights* w) {
  if (!out) return set_error(LXG_EINVAL, "out is NULL");
  *out = nullptr;
  if (!w || !w->layer) return set_error(LXG_EINVAL, "weights are NULL");
  if (!lxg::encode_tensor_map_ready()) return set_error(LXG_EINVAL, "lxg_init has not been called");
  if (w->head_dim != kHeadDim) return set_error(LXG_EUNSUPPORTED, "decoder geometry: head_dim must be 128 (every Qwen3 size)");
  if (w->hidden <= 0 || w->hidden > 1024 || w->hidden % kGemmBN != 0 || w->ffn <= 0 || w->ffn % 64 != 0 || w->layers <= 0 ||
      w->heads <= 0 || w->kv_heads <= 0 || w->heads % w->kv_heads != 0 || w->vocab <= 0)
    return set_error(LXG_EUNSUPPORTED,
                     "decoder geometry: hidden must be a multiple of 128 and <= 1024, ffn a multiple of 64, heads a multiple of kv_heads");
  const void* globals[] = {w->tok_emb, w->final_norm, w->inv_freq};
  for (const void* p : globals)
    if (!p || !is_device_ptr(p)) return set_error(LXG_EINVAL, "tok_emb / final_norm / inv_freq must be device memory");
  if (w->lm_head && !is_device_ptr(w->lm_head)) return set_error(LXG_EINVAL, "lm_head must be device memory");
  const int device = device_of_ptr(w->tok_emb);
  DeviceGuard guard(device);
  if (lxg::num_sms() == 0) return set_error(LXG_EINVAL, "lxg_init has not been called for the device that holds the weights");
  lxg_decoder* e = new lxg_decoder();
  e->device = device;
  e->w = *w;
  e->layers.assign(w->layer, w->layer + w->layers);
  e->w.layer = e->layers.data();
  const int H = w->hidden, F = w->ffn, QKV = (w->heads + 2 * w->kv_heads) * kHeadDim, C = w->heads * kHeadDim;
  e->map_wqkv.resize(w->layers);
  e->map_wo.resize(w->layers);
  e->map_wgu.resize(w->layers);
  e->map_wdown.resize(w->layers);
  e->map_wo64.resize(w->layers);
  e->map_wdown64.resize(w->layers);
  e->map_wqkv32.resize(w->layers);
  e->map_wgu64.resize(w->layers);
  for (int l = 0; l < w->layers; ++l) {
    const lxg_qwen3_layer& L = e->layers[l];
    const void* ptrs[] = {L.ln1, L.wqkv, L.q_norm, L.k_norm, L.wo, L.ln2, L.wgu, L.wdown};
    for (const void* p : ptrs)
      if (!p || !is_device_ptr(p)) {
        delete e;
        return set_error(LXG_EINVAL, "layer " + std::to_string(l) + ": weight pointer is not device memory");
      }
    int rc;
    if ((rc = make_map(&e->map_wqkv[l], L.wqkv, QKV, H)) != LXG_OK || (rc = make_map(&e->map_wo[l], L.wo, H, C)) != LXG_OK ||
        (rc = make_map(&e->map_wgu[l], L.wgu, 2 * F, H)) != LXG_OK || (rc = make_map(&e->map_wdown[l], L.wdown, H, F)) != LXG_OK ||
        (rc = make_map(&e->map_wo64[l], L.wo, H, C, 64)) != LXG_OK || (rc = make_map(&e->map_wdown64[l], L.wdown, H, F, 64)) != LXG_OK ||
        (rc = make_map(&e->map_wqkv32[l], L.wqkv, QKV, H, 32)) != LXG_OK || (rc = make_map(&e->map_wgu64[l], L.wgu, 2 * F, H, 64)) != LXG_OK) {
      delete e;
      return rc;
    }
  }
  const char* pk = std::getenv("LXG_DECODER_PACK");
  e->pack = !(pk && pk[0] == '0');
  const char* pd = std::getenv("LXG_PDL");
  e->pdl = !(pd && pd[0] == '0');
  const char* gt = std::getenv("LXG_GEMM_TRACE");
  if (gt && gt[0]) {
    e->trace_layer = std::atoi(gt);
    e->use_graphs = false;
    if (cudaMalloc(&e->trace, 4 * 296 * 16 * sizeof(unsigned long long)) == cudaSuccess)
      cudaMemset(e->trace, 0, 4 * 296 * 16 * sizeof(unsigned long long));
    else
      e->trace = nullptr;
  }
  const char* rb = std::getenv("LXG_ROPE_BULK");
  e->rope_bulk = !(rb && rb[0] == '0');
  const char* qg = std::getenv("LXG_QUERY_GEMM");
  e->query_gemm = !(qg && qg[0] == '0');
  const char* at = std::getenv("LXG_ATTN_TC");
  e->attn_tc = !(at && at[0] == '0');
  *out = e;
  return LXG_OK;
}

int lxg_decoder_destroy(lxg_decoder* e) {
  if (!e) return LXG_OK;
  DeviceGuard guard(e->device);
  free_ws(e);
  cudaFree(e->out_buf);
  cudaFree(e->trace);
  cudaFree(e->map_resid_dev);
  cudaFree(e->cu);
  if (e->stage) cudaFreeHost(e->stage);
  if (e->own) cudaStreamDestroy(e->own);
  if (e->ev_in) cudaEventDestroy(e->ev_in);
  if (e->ev_out) cudaEventDestroy(e->ev_out);
  delete e;
  return LXG_OK;
}

int lxg_decoder_last_launches(const lxg_decoder* e) { return e ? e->launches : -1; }
int lxg_decoder_last_tokens(const lxg_decoder* e) { return e ? e->last_tokens : -1; }

int lxg_decoder_embed(lxg_decoder* e, const int32_t* ids, const int32_t* mask, int32_t b, int32_t s, float* out, void* stream) {
  return run(e, ids, mask, b, s, 0, 0, 0, out, stream);
}

int lxg_decoder_rerank(lxg_decoder* e, const int32_t* ids, const int32_t* mask, int32_t b, int32_t s, int32_t token_true,
                       int32_t token_false, float* scores, void* stream) {
  return run(e, ids, mask, b, s, 1, token_true, token_false, scores, stream);
}

}  // extern "C"
