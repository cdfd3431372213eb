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
#include "fused_encoder.cuh"
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
  // Query path (<= 64 tokens): the whole forward as one cooperative kernel (fused_encoder.cuh).
  // 0 = not probed yet, 1 = available, -1 = geometry / device cannot run it (layered path only).
  int fused_state = 0;
  int fused_grid = 0;
  FusedLayer* fused_layers = nullptr;  // device array, one entry per layer (tensor maps + vectors)
  uint8_t* fused_ws = nullptr;         // activations, split-K slabs, counters
  FusedParams fused_base{};
  bool last_fused = false;
  bool fused_allowed = true;  // lxg_encoder_set_fused
  bool fused_trace = false;   // lxg_encoder_set_fused(enc, 2): per-phase %globaltimer stamps of every CTA
  unsigned long long* trace_buf = nullptr;
};

namespace {

void drop_graphs(lxg_encoder* e) {
  for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
}

constexpr size_t kFusedSmemLimit = 226 * 1024;  // dynamic part; the kernel also has ~300 B of static shared memory

// shared memory of the fused kernel besides the weight ring: token operand + attention scratch + alignment
size_t fused_fixed_smem(int tpad, int H, int dh) {
  return static_cast<size_t>(tpad) * H * 2 + static_cast<size_t>(2) * tpad * (dh + 8) * 2 + static_cast<size_t>(dh) * (tpad + 8) * 2 +
         static_cast<size_t>(tpad) * 4 + static_cast<size_t>(2) * H * 4 + 1024 + 256;
}

using FusedKernel = void (*)(FusedParams);
// instantiated geometries: (head size, hidden / 128); anything else stays on the layered kernels
FusedKernel fused_kernel_for(int dh, int nv) {
  if (dh == 32 && nv == 1) return bert_fused_kernel<32, 1>;
  if (dh == 32 && nv == 3) return bert_fused_kernel<32, 3>;   // MiniLM-L6 / L12 (H = 384, 12 heads)
  if (dh == 64 && nv == 6) return bert_fused_kernel<64, 6>;   // BERT-base / bge-base (H = 768, 12 heads)
  if (dh == 64 && nv == 8) return bert_fused_kernel<64, 8>;   // BERT-large / bge-large (H = 1024, 16 heads)
  return nullptr;
}

bool fused_enabled() {
  static const bool on = [] {
    const char* v = std::getenv("LXG_FUSED");
    return !(v && v[0] == '0');
  }();
  return on;
}

// One-time set-up of the single-kernel query path; leaves fused_state = -1 when this model cannot use it.
int fused_prepare(lxg_encoder* e) {
  e->fused_state = -1;
  const int H = e->w.hidden, F = e->w.ffn, heads = e->w.heads, dh = H / heads, L = e->w.layers;
  if (H % 128 != 0 || F % H != 0 || F % 128 != 0 || L > 64) return LXG_OK;
  const FusedKernel kernel = fused_kernel_for(dh, H / 128);
  if (kernel == nullptr) return LXG_OK;
  std::vector<FusedLayer> host(L);
  for (int l = 0; l < L; ++l) {
    const lxg_bert_layer& W = e->layers[l];
    int rc;
    if ((rc = make_map(&host[l].map_qkv, W.wqkv, 3 * H, H, dh)) != LXG_OK || (rc = make_map(&host[l].map_wo, W.wo, H, H)) != LXG_OK ||
        (rc = make_map(&host[l].map_w1, W.w1, F, H)) != LXG_OK || (rc = make_map(&host[l].map_w2, W.w2, H, F)) != LXG_OK)
      return rc;
    host[l].bqkv = reinterpret_cast<const float*>(W.bqkv);
    host[l].bo = reinterpret_cast<const float*>(W.bo);
    host[l].ln1_g = reinterpret_cast<const float*>(W.ln1_g);
    host[l].ln1_b = reinterpret_cast<const float*>(W.ln1_b);
    host[l].b1 = reinterpret_cast<const float*>(W.b1);
    host[l].b2 = reinterpret_cast<const float*>(W.b2);
    host[l].ln2_g = reinterpret_cast<const float*>(W.ln2_g);
    host[l].ln2_b = reinterpret_cast<const float*>(W.ln2_b);
  }
  LXG_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(kernel), kFusedSmemLimit));
  int per_sm = 0;
  LXG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kFusedThreads, kFusedSmemLimit));
  int coop = 0;
  LXG_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->device));
  if (per_sm < 1 || !coop || lxg::num_sms() < 8) return LXG_OK;
  e->fused_grid = lxg::num_sms();
  LXG_CUDA(cudaMalloc(&e->fused_layers, L * sizeof(FusedLayer)));
  LXG_CUDA(cudaMemcpy(e->fused_layers, host.data(), L * sizeof(FusedLayer), cudaMemcpyHostToDevice));
  // workspace: h0 | h1 | ctx (fp16 [64, H]) | act (fp16 [64, F]) | pre1 | pre2 (fp32 [64, H]) | slabs | sem | bar
  const size_t T = kFusedMaxTokens;
  const size_t b_h = T * H * 2, b_act = T * F * 2, b_pre = T * H * 4, b_slab = static_cast<size_t>(H / 128) * (F / H) * T * 128 * 4;
  const size_t b_ctr = (static_cast<size_t>(H / 128) + 4 * L + 64) * 4;
  const size_t total = 3 * b_h + b_act + 2 * b_pre + b_slab + b_ctr;
  LXG_CUDA(cudaMalloc(&e->fused_ws, total));
  LXG_CUDA(cudaMemset(e->fused_ws, 0, total));
  uint8_t* q = e->fused_ws;
  FusedParams& fp = e->fused_base;
  fp.layers = e->fused_layers;
  fp.num_layers = L;
  fp.hidden = H;
  fp.ffn = F;
  fp.heads = heads;
  fp.vocab = e->w.vocab;
  fp.eps = e->w.ln_eps;
  fp.word = reinterpret_cast<const __half*>(e->w.word_emb);
  fp.pos = reinterpret_cast<const __half*>(e->w.pos_emb);
  fp.type0 = reinterpret_cast<const __half*>(e->w.type_emb);
  fp.emb_g = reinterpret_cast<const float*>(e->w.emb_ln_g);
  fp.emb_b = reinterpret_cast<const float*>(e->w.emb_ln_b);
  fp.h0 = reinterpret_cast<__half*>(q), q += b_h;
  fp.h1 = reinterpret_cast<__half*>(q), q += b_h;
  fp.ctx = reinterpret_cast<__half*>(q), q += b_h;
  fp.act = reinterpret_cast<__half*>(q), q += b_act;
  fp.pre1 = reinterpret_cast<float*>(q), q += b_pre;
  fp.pre2 = reinterpret_cast<float*>(q), q += b_pre;
  fp.partial = reinterpret_cast<float*>(q), q += b_slab;
  fp.sem = reinterpret_cast<unsigned*>(q), q += static_cast<size_t>(H / 128) * 4;
  fp.bar = reinterpret_cast<unsigned*>(q);
  e->fused_state = 1;
  return LXG_OK;
}

// Whether this call goes through the single kernel.
bool fused_applies(const lxg_encoder* e, int tokens) {
  if (e->fused_state != 1 || tokens > kFusedMaxTokens) return false;
  const int tpad = (tokens + 15) / 16 * 16;
  const int H = e->w.hidden, dh = H / e->w.heads;
  // measured cross-over with the layered kernels (profiles/r2e_encoder_query.json): the per-phase cost
  // of the single kernel grows with the rows every CTA normalises; H >= 768 only wins up to 32 tokens
  if (H > 512 && tokens > 32) return false;
  return fused_fixed_smem(tpad, H, dh) + 3 * static_cast<size_t>(kFusedSlotBytes) <= kFusedSmemLimit;
}

int launch_fused(lxg_encoder* e, int b, int s, int pool, cudaStream_t st) {
  FusedParams fp = e->fused_base;
  const int tokens = b * s, H = e->w.hidden, dh = H / e->w.heads;
  fp.tokens = tokens;
  fp.seq = s;
  fp.batch = b;
  fp.tpad = (tokens + 15) / 16 * 16;
  fp.ids = e->ids;
  fp.mask = e->mask;
  fp.pool_cls = pool == LXG_POOL_CLS ? 1 : 0;
  fp.out = e->out_buf;
  fp.trace = nullptr;
  if (e->fused_trace) {
    const size_t n = static_cast<size_t>(e->fused_grid) * (4 * e->w.layers + 1) * 6;
    if (!e->trace_buf) LXG_CUDA(cudaMalloc(&e->trace_buf, n * sizeof(unsigned long long)));
    LXG_CUDA(cudaMemsetAsync(e->trace_buf, 0, n * sizeof(unsigned long long), st));
    fp.trace = e->trace_buf;
  }
  const size_t fixed = fused_fixed_smem(fp.tpad, H, dh);
  fp.nslots = static_cast<int>(std::min<size_t>(kFusedMaxSlots, (kFusedSmemLimit - fixed) / kFusedSlotBytes));
  const size_t smem = fixed + static_cast<size_t>(fp.nslots) * kFusedSlotBytes;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(e->fused_grid);
  cfg.blockDim = dim3(kFusedThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the phase barriers cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LXG_CUDA(cudaLaunchKernelEx(&cfg, fused_kernel_for(dh, H / 128), fp));
  e->launches = 1;
  return LXG_OK;
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
  cudaFree(e->fused_layers);
  cudaFree(e->fused_ws);
  cudaFree(e->trace_buf);
  if (e->own) cudaStreamDestroy(e->own);
  if (e->ev_in) cudaEventDestroy(e->ev_in);
  if (e->ev_out) cudaEventDestroy(e->ev_out);
  delete e;
  return LXG_OK;
}

int lxg_encoder_last_launches(const lxg_encoder* e) { return e ? e->launches : -1; }

int lxg_encoder_set_fused(lxg_encoder* e, int enabled) {
  if (!e) return set_error(LXG_EINVAL, "NULL encoder");
  std::lock_guard<std::mutex> lock(e->mu);
  e->fused_allowed = enabled != 0;
  e->fused_trace = enabled == 2;
  return LXG_OK;
}

int lxg_encoder_read_trace(lxg_encoder* e, uint64_t* out, int32_t capacity, int32_t* grid, int32_t* phases) {
  if (!e || !out || !grid || !phases) return set_error(LXG_EINVAL, "NULL argument");
  DeviceGuard guard(e->device);
  std::lock_guard<std::mutex> lock(e->mu);
  if (!e->trace_buf) return set_error(LXG_EINVAL, "no trace recorded (lxg_encoder_set_fused(enc, 2), then a call of <= 64 tokens)");
  *grid = e->fused_grid;
  *phases = 4 * e->w.layers + 1;
  const size_t n = static_cast<size_t>(*grid) * *phases * 6;
  if (static_cast<size_t>(capacity) < n) return set_error(LXG_EINVAL, "trace buffer too small");
  if (e->own) LXG_CUDA(cudaStreamSynchronize(e->own));
  LXG_CUDA(cudaMemcpy(out, e->trace_buf, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return LXG_OK;
}

int lxg_encode(lxg_encoder* e, const int32_t* ids, const int32_t* mask, int32_t b, int32_t s, int pool, float* out,
               void* stream) {
  if (!e || !ids || !mask || !out) return set_error(LXG_EINVAL, "NULL argument");
  if (b < 0 || s <= 0) return set_error(LXG_EINVAL, "b must be >= 0 and s >= 1");
  if (s > e->w.max_pos) return set_error(LXG_EINVAL, "sequence longer than the position table");
  if (pool != LXG_POOL_MEAN && pool != LXG_POOL_CLS) return set_error(LXG_EINVAL, "bad pooling mode");
  if (b == 0) return LXG_OK;
  DeviceGuard guard(e->device);
  NvtxRange nvtx("lxg_encode");
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
  const bool want_fused = fused_enabled() && e->fused_allowed;
  if (e->fused_state == 0 && want_fused) {
    rc = fused_prepare(e);
    if (rc != LXG_OK) return rc;
  }
  const bool fused = want_fused && fused_applies(e, tokens);
  e->last_fused = fused;
  if (fused) {
    rc = launch_fused(e, b, s, pool, st);
    if (rc != LXG_OK) return rc;
  }
  cudaGraphExec_t exec = nullptr;
  if (!fused && e->use_graphs) {
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
  if (!fused && !e->use_graphs) {
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
