// Sentence-encoder entry points of the C ABI (include/lxg.h): host logic only - workspace,
// tensor maps, launch sequence.  Replaces SentenceTransformer.encode as called by
// EmbeddingClient.embed (reference src/lean_explore/util/embedding_client.py:88-101).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lxg.h"
#include "common.h"
#include "encoder_kernels.cuh"
#include "gemm_host.cuh"

using namespace lxg;

struct lxg_encoder {
  int device = 0;  // the GPU that holds the weights; made current by every entry point
  lxg_bert_weights w{};
  std::vector<lxg_bert_layer> layers;
  std::mutex mu;
  int cap_tokens = 0;  // workspace capacity (tokens)
  __half *h = nullptr, *qkv = nullptr, *ctx = nullptr, *ffn = nullptr;
  float* pre = nullptr;
  int *ids = nullptr, *mask = nullptr;
  CUtensorMap map_h{}, map_ctx{}, map_ffn{};  // A operands (activations)
  std::vector<CUtensorMap> map_wqkv, map_wo, map_w1, map_w2;
  int launches = 0;
  // One captured CUDA graph per (b, s, pool): a query-time forward pass is 2 + 7 * layers tiny
  // kernels, i.e. launch bound; replaying a graph removes the per-launch host cost and most of the
  // inter-kernel gaps.  All graph nodes use workspace pointers only (ids/mask/out are staged).
  struct Graph {
    int b, s, pool;
    cudaGraphExec_t exec;
  };
  std::vector<Graph> graphs;
  bool use_graphs = true;
  float* out_buf = nullptr;  // [out_cap, H] pooled vectors (graph nodes cannot point at caller memory)
  int out_cap = 0;
  // The forward runs on a private stream ordered after / before the caller's stream by events:
  // the caller's stream may be the legacy default stream, which cannot be captured.
  cudaStream_t own = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
};

namespace {

void drop_graphs(lxg_encoder* e) {
  for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
}

void free_ws(lxg_encoder* e) {
  drop_graphs(e);  // they reference the workspace
  cudaFree(e->out_buf);
  e->out_buf = nullptr;
  e->out_cap = 0;
  cudaFree(e->h);
  cudaFree(e->qkv);
  cudaFree(e->ctx);
  cudaFree(e->ffn);
  cudaFree(e->pre);
  cudaFree(e->ids);
  cudaFree(e->mask);
  e->h = e->qkv = e->ctx = e->ffn = nullptr;
  e->pre = nullptr;
  e->ids = e->mask = nullptr;
  e->cap_tokens = 0;
}

int reserve_ws(lxg_encoder* e, int tokens) {
  if (tokens <= e->cap_tokens) return LXG_OK;
  free_ws(e);
  const int cap = (std::max(tokens, 256) + 127) / 128 * 128;
  const size_t H = e->w.hidden, F = e->w.ffn;
  LXG_CUDA(cudaMalloc(&e->h, cap * H * sizeof(__half)));
  LXG_CUDA(cudaMalloc(&e->qkv, cap * 3 * H * sizeof(__half)));
  LXG_CUDA(cudaMalloc(&e->ctx, cap * H * sizeof(__half)));
  LXG_CUDA(cudaMalloc(&e->ffn, cap * F * sizeof(__half)));
  LXG_CUDA(cudaMalloc(&e->pre, cap * H * sizeof(float)));
  LXG_CUDA(cudaMalloc(&e->ids, cap * sizeof(int)));
  LXG_CUDA(cudaMalloc(&e->mask, cap * sizeof(int)));
  // rows beyond the live tokens are read by TMA (never stored): keep them finite
  LXG_CUDA(cudaMemset(e->h, 0, cap * H * sizeof(__half)));
  LXG_CUDA(cudaMemset(e->ctx, 0, cap * H * sizeof(__half)));
  LXG_CUDA(cudaMemset(e->ffn, 0, cap * F * sizeof(__half)));
  int rc;
  if ((rc = make_map(&e->map_h, e->h, cap, static_cast<int>(H))) != LXG_OK) return rc;
  if ((rc = make_map(&e->map_ctx, e->ctx, cap, static_cast<int>(H))) != LXG_OK) return rc;
  if ((rc = make_map(&e->map_ffn, e->ffn, cap, static_cast<int>(F))) != LXG_OK) return rc;
  e->cap_tokens = cap;
  return LXG_OK;
}

// The forward pass proper: 2 + 7 * layers launches on `st`, workspace pointers only.
int launch_forward(lxg_encoder* e, int b, int s, int pool, cudaStream_t st) {
  const int tokens = b * s;
  const int H = e->w.hidden, F = e->w.ffn, heads = e->w.heads, dh = H / heads;
  int launches = 0;
  const int warps_per_block = 8;
  const int row_blocks = (tokens + warps_per_block - 1) / warps_per_block;
  // the first kernel follows the input copies; every later one is a programmatic dependent launch
  // (its prologue overlaps the previous kernel's tail, also inside the captured graph; LXG_PDL=0 disables)
  static const bool pdl = [] {
    const char* v = std::getenv("LXG_PDL");
    return !(v && v[0] == '0');
  }();
  LXG_CUDA(lxg_launch(embed_ln_kernel, dim3(row_blocks), dim3(256), 0, st, false, static_cast<const int*>(e->ids), tokens, s, H, e->w.vocab,
                      reinterpret_cast<const __half*>(e->w.word_emb), reinterpret_cast<const __half*>(e->w.pos_emb),
                      reinterpret_cast<const __half*>(e->w.type_emb), reinterpret_cast<const float*>(e->w.emb_ln_g),
                      reinterpret_cast<const float*>(e->w.emb_ln_b), e->w.ln_eps, e->h));
  ++launches;
  // attention: tensor-core kernel for head sizes 32 / 64 (every shipped model), scalar fallback otherwise
  const bool attn_mma = dh == 32 || dh == 64;
  const int seq_pad = (s + 15) / 16 * 16;
  const int attn_warps = std::min(8, seq_pad / 16);
  const size_t attn_smem = attn_mma
      ? static_cast<size_t>(seq_pad) * (dh + 8) * sizeof(__half) + static_cast<size_t>(dh) * (seq_pad + 8) * sizeof(__half) +
            static_cast<size_t>(seq_pad) * sizeof(float)
      : static_cast<size_t>(2) * s * (dh + 2) * sizeof(__half) + static_cast<size_t>(s) * sizeof(float) +
            static_cast<size_t>(kAttnThreads / 32) * s * sizeof(float);
  if (attn_smem > 200 * 1024) return set_error(LXG_EUNSUPPORTED, "sequence too long for the attention kernel's shared memory");
  const int attn_which = !attn_mma ? 0 : (dh == 32 ? 1 : 2);
  if (attn_which == 0)
    LXG_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(&attention_scalar_kernel), attn_smem));
  else if (attn_which == 1)
    LXG_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(&attention_mma_kernel<32>), attn_smem));
  else
    LXG_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(&attention_mma_kernel<64>), attn_smem));
  for (int l = 0; l < e->w.layers; ++l) {
    const lxg_bert_layer& L = e->layers[l];
    GemmParams gp{};
    // QKV projection
    gp.bias = reinterpret_cast<const float*>(L.bqkv);
    gp.out = e->qkv;
    gp.m = tokens;
    gp.n = 3 * H;
    gp.k = H;
    LXG_CUDA(launch_gemm<kEpiStore>(e->map_h, e->map_wqkv[l], gp, st, pdl));
    const __half* qkv_c = e->qkv;
    const int* mask_c = e->mask;
    if (attn_which == 1)
      LXG_CUDA(lxg_launch(attention_mma_kernel<32>, dim3(b * heads), dim3(attn_warps * 32), attn_smem, st, pdl, qkv_c, mask_c, s, H, heads, e->ctx));
    else if (attn_which == 2)
      LXG_CUDA(lxg_launch(attention_mma_kernel<64>, dim3(b * heads), dim3(attn_warps * 32), attn_smem, st, pdl, qkv_c, mask_c, s, H, heads, e->ctx));
    else
      LXG_CUDA(lxg_launch(attention_scalar_kernel, dim3(b * heads), dim3(kAttnThreads), attn_smem, st, pdl, qkv_c, mask_c, s, H, heads, e->ctx));
    // attention.output.dense + residual -> LayerNorm
    gp.bias = reinterpret_cast<const float*>(L.bo);
    gp.residual = e->h;
    gp.out = e->pre;
    gp.n = H;
    gp.k = H;
    LXG_CUDA(launch_gemm<kEpiResid>(e->map_ctx, e->map_wo[l], gp, st, pdl));
    LXG_CUDA(lxg_launch(layernorm_kernel, dim3(row_blocks), dim3(256), 0, st, pdl, static_cast<const float*>(e->pre), tokens, H,
                        reinterpret_cast<const float*>(L.ln1_g), reinterpret_cast<const float*>(L.ln1_b), e->w.ln_eps, e->h));
    // intermediate.dense + GELU
    gp.bias = reinterpret_cast<const float*>(L.b1);
    gp.residual = nullptr;
    gp.out = e->ffn;
    gp.n = F;
    gp.k = H;
    LXG_CUDA(launch_gemm<kEpiGelu>(e->map_h, e->map_w1[l], gp, st, pdl));
    // output.dense + residual -> LayerNorm
    gp.bias = reinterpret_cast<const float*>(L.b2);
    gp.residual = e->h;
    gp.out = e->pre;
    gp.n = H;
    gp.k = F;
    LXG_CUDA(launch_gemm<kEpiResid>(e->map_ffn, e->map_w2[l], gp, st, pdl));
    LXG_CUDA(lxg_launch(layernorm_kernel, dim3(row_blocks), dim3(256), 0, st, pdl, static_cast<const float*>(e->pre), tokens, H,
                        reinterpret_cast<const float*>(L.ln2_g), reinterpret_cast<const float*>(L.ln2_b), e->w.ln_eps, e->h));
    launches += 7;
  }
  LXG_CUDA(lxg_launch(pool_normalize_kernel, dim3(b), dim3(256), H * sizeof(float), st, pdl, static_cast<const __half*>(e->h),
                      static_cast<const int*>(e->mask), s, H, pool == LXG_POOL_CLS ? 1 : 0, e->out_buf));
  ++launches;
  e->launches = launches;
  return LXG_OK;
}

}  // namespace

extern "C" {

int lxg_encoder_create(lxg_encoder** out, const lxg_bert_weights* w) {
  if (!out) return set_error(LXG_EINVAL, "out is NULL");
  *out = nullptr;
  if (!w || !w->layer) return set_error(LXG_EINVAL, "weights are NULL");
  if (!lxg::encode_tensor_map_ready()) return set_error(LXG_EINVAL, "lxg_init has not been called");
  if (w->hidden <= 0 || w->hidden > 1024 || w->hidden % kGemmBN != 0 || w->ffn % kGemmBN != 0 || w->layers <= 0 ||
      w->heads <= 0 || w->hidden % w->heads != 0)
    return set_error(LXG_EUNSUPPORTED, "encoder geometry: hidden and ffn must be multiples of 128, hidden <= 1024");
  const int dh = w->hidden / w->heads;
  if (dh > 64 || dh % 2 != 0) return set_error(LXG_EUNSUPPORTED, "encoder geometry: head size must be even and <= 64");
  if (!w->word_emb || !w->pos_emb || !w->type_emb || !w->emb_ln_g || !w->emb_ln_b)
    return set_error(LXG_EINVAL, "embedding weights are NULL");
  const int device = device_of_ptr(w->word_emb);
  if (device < 0) return set_error(LXG_EINVAL, "embedding weights are not device memory");
  DeviceGuard guard(device);
  if (lxg::num_sms() == 0) return set_error(LXG_EINVAL, "lxg_init has not been called for the device that holds the weights");
  lxg_encoder* e = new lxg_encoder();
  e->device = device;
  e->w = *w;
  e->layers.assign(w->layer, w->layer + w->layers);
  e->w.layer = e->layers.data();
  const int H = w->hidden, F = w->ffn;
  e->map_wqkv.resize(w->layers);
  e->map_wo.resize(w->layers);
  e->map_w1.resize(w->layers);
  e->map_w2.resize(w->layers);
  for (int l = 0; l < w->layers; ++l) {
    const lxg_bert_layer& L = e->layers[l];
    const void* ptrs[] = {L.wqkv, L.bqkv, L.wo, L.bo, L.ln1_g, L.ln1_b, L.w1, L.b1, L.w2, L.b2, L.ln2_g, L.ln2_b};
    for (const void* p : ptrs)
      if (!p || !is_device_ptr(p)) {
        delete e;
        return set_error(LXG_EINVAL, "layer " + std::to_string(l) + ": weight pointer is not device memory");
      }
    int rc;
    if ((rc = make_map(&e->map_wqkv[l], L.wqkv, 3 * H, H)) != LXG_OK || (rc = make_map(&e->map_wo[l], L.wo, H, H)) != LXG_OK ||
        (rc = make_map(&e->map_w1[l], L.w1, F, H)) != LXG_OK || (rc = make_map(&e->map_w2[l], L.w2, H, F)) != LXG_OK) {
      delete e;
      return rc;
    }
  }
  *out = e;
  return LXG_OK;
}

int lxg_encoder_destroy(lxg_encoder* e) {
  if (!e) return LXG_OK;
  DeviceGuard guard(e->device);
  free_ws(e);
  if (e->own) cudaStreamDestroy(e->own);
  if (e->ev_in) cudaEventDestroy(e->ev_in);
  if (e->ev_out) cudaEventDestroy(e->ev_out);
  delete e;
  return LXG_OK;
}

int lxg_encoder_last_launches(const lxg_encoder* e) { return e ? e->launches : -1; }

int lxg_encode(lxg_encoder* e, const int32_t* ids, const int32_t* mask, int32_t b, int32_t s, int pool, float* out,
               void* stream) {
  if (!e || !ids || !mask || !out) return set_error(LXG_EINVAL, "NULL argument");
  if (b < 0 || s <= 0) return set_error(LXG_EINVAL, "b must be >= 0 and s >= 1");
  if (s > e->w.max_pos) return set_error(LXG_EINVAL, "sequence longer than the position table");
  if (pool != LXG_POOL_MEAN && pool != LXG_POOL_CLS) return set_error(LXG_EINVAL, "bad pooling mode");
  if (b == 0) return LXG_OK;
  DeviceGuard guard(e->device);
  std::lock_guard<std::mutex> lock(e->mu);
  cudaStream_t caller = reinterpret_cast<cudaStream_t>(stream);
  const long long tokens_ll = static_cast<long long>(b) * s;
  if (tokens_ll > (1 << 22)) return set_error(LXG_EUNSUPPORTED, "more than 4M tokens per call");
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
  LXG_CUDA(cudaMemcpyAsync(e->ids, ids, tokens * sizeof(int), ids_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  LXG_CUDA(cudaMemcpyAsync(e->mask, mask, tokens * sizeof(int), mask_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  // pooled vectors land in a workspace buffer first (fixed address: graph-capturable)
  if (b > e->out_cap) {
    drop_graphs(e);
    cudaFree(e->out_buf);
    e->out_buf = nullptr;
    e->out_cap = 0;
    const int cap = std::max(b, 64);
    LXG_CUDA(cudaMalloc(&e->out_buf, static_cast<size_t>(cap) * H * sizeof(float)));
    e->out_cap = cap;
  }
  cudaGraphExec_t exec = nullptr;
  if (e->use_graphs) {
    for (auto& g : e->graphs)
      if (g.b == b && g.s == s && g.pool == pool) exec = g.exec;
    if (!exec) {
      // first call with this shape: run once eagerly (sets kernel attributes, validates the
      // launch configuration), then capture the same sequence for every later call
      rc = launch_forward(e, b, s, pool, st);
      if (rc != LXG_OK) return rc;
      cudaGraph_t graph = nullptr;
      if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        const int rc2 = launch_forward(e, b, s, pool, st);
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc2 == LXG_OK && ce == cudaSuccess && graph &&
            cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
          if (e->graphs.size() >= 32) drop_graphs(e);
          e->graphs.push_back({b, s, pool, exec});
        } else {
          exec = nullptr;
          e->use_graphs = false;  // capture unsupported here: stay on plain launches
        }
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
      } else {
        cudaGetLastError();
        e->use_graphs = false;
      }
      exec = nullptr;  // this call already ran eagerly
    } else {
      LXG_CUDA(cudaGraphLaunch(exec, st));
    }
  }
  if (!e->use_graphs) {
    rc = launch_forward(e, b, s, pool, st);
    if (rc != LXG_OK) return rc;
  }
  const size_t out_bytes = static_cast<size_t>(b) * H * sizeof(float);
  LXG_CUDA(cudaMemcpyAsync(out, e->out_buf, out_bytes, out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  if (!out_dev) {
    LXG_CUDA(cudaStreamSynchronize(st));
  } else {
    LXG_CUDA(cudaEventRecord(e->ev_out, st));
    LXG_CUDA(cudaStreamWaitEvent(caller, e->ev_out, 0));
  }
  return LXG_OK;
}

}  // extern "C"
