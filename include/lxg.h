/* lxg.h - C ABI of the B200 semantic-search hot path of lean-explore.
 *
 * This is the whole drop-in boundary: plain pointers and sizes, no torch / numpy / C++ types.
 * Every entry point names the reference interface it replaces (paths relative to the
 * reference repository justincasher/lean-explore @ 3a52d6b).  The reference delegates this
 * arithmetic to the third-party wheels faiss-cpu (>=1.7) and sentence-transformers (>=2.2);
 * the Python host side in lean_explore_b200/ binds these symbols with ctypes
 * (INTEGRATION.md shows the stub a maintainer would add to the reference).
 *
 * Conventions
 *   - every function returns 0 on success, a negative LXG_E* code otherwise, and never
 *     aborts or exits (the MCP server maps start-up failures to sys.exit(1) itself,
 *     src/lean_explore/mcp/server.py:153-181).  lxg_last_error() gives the message of the
 *     last failure on the calling thread.
 *   - nothing is ever printed to stdout (stdout is the MCP JSON-RPC channel,
 *     src/lean_explore/mcp/server.py:33-38).
 *   - device memory handed in (corpus, weights) stays owned by the caller (a torch.Tensor on
 *     the Python side) and must outlive the handle.
 *   - `stream` is a cudaStream_t passed as void*; NULL is the legacy default stream.
 *   - handles may be used from any host thread and with any stream; calls on one handle are
 *     serialised on the host (per-handle mutex) AND on the device (a search waits for the previous
 *     search of the same handle, whatever stream that ran on, before it touches the shared
 *     workspaces).  Every entry point makes the handle's device current for the duration of the
 *     call and restores the caller's device, so executor threads that start on device 0 are fine;
 *     call lxg_init once for every device that will hold a handle.
 *   - there is no CPU fallback: without an sm_100 device lxg_init fails.
 */
#ifndef LXG_H_
#define LXG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LXG_OK 0
#define LXG_EINVAL (-1)      /* bad argument */
#define LXG_ECUDA (-2)       /* CUDA runtime / driver error */
#define LXG_ENODEVICE (-3)   /* no sm_100 GPU */
#define LXG_EUNSUPPORTED (-4)/* shape outside what the kernels cover */
#define LXG_ETIES (-5)       /* > 16384 rows tie with the k-th score: exact result not representable */

#define LXG_F32 0
#define LXG_F16 1

#define LXG_POOL_MEAN 0      /* sentence-transformers Pooling(mean) - all-MiniLM-L6-v2 */
#define LXG_POOL_CLS 1       /* sentence-transformers Pooling(cls)  - bge-base-en-v1.5 */

typedef struct lxg_index lxg_index;
typedef struct lxg_encoder lxg_encoder;

/* Library / device bring-up.  Selects `device`, checks it is compute capability 10.x and
 * resolves the driver's tensor-map encoder.  Idempotent.  Replaces nothing in the reference
 * (faiss-cpu needs no device); it is where "no CPU fallback" is enforced. */
int lxg_init(int device);

/* Message of the last failure on this thread ("" if none). */
const char* lxg_last_error(void);

/* ABI version, bumped on any signature change. */
int lxg_abi_version(void);

/* ---- flat inner-product index ------------------------------------------------------
 * lxg_index_create replaces faiss.read_index()/IndexFlatIP.add() as used by
 * SearchEngine._ensure_faiss_loaded (src/lean_explore/search/engine.py:151-161) on the matrix
 * that extract/index.py:59-71 defines: `corpus_dev` is that [n, d] row-major matrix in device
 * memory (fp32 as stored by the reference, or fp16), row i <-> faiss label i + row_offset.
 * The rows are NOT normalised here (the reference does not either, extract/index.py:103-116).
 * `row_offset` is the global row number of row 0 (row-sharded multi-GPU indexes). */
int lxg_index_create(lxg_index** out, const void* corpus_dev, int64_t n, int32_t d, int dtype,
                     int64_t row_offset);
int lxg_index_destroy(lxg_index* index);
int64_t lxg_index_ntotal(const lxg_index* index); /* faiss.Index.ntotal */
int32_t lxg_index_d(const lxg_index* index);      /* faiss.Index.d */

/* lxg_search replaces faiss.normalize_L2(x) + index.search(x, k)
 * (src/lean_explore/search/engine.py:242 and :250) with IndexFlatIP semantics:
 *   x        [nq, d] float32 queries, host or device memory (detected), NOT modified;
 *   normalize != 0 applies faiss.normalize_L2 first (zero rows stay zero);
 *   D_out    [nq, k] float32 inner products, best first; I_out [nq, k] int64 row ids;
 *            host or device memory (detected); rows with fewer than k results are padded
 *            with D = -FLT_MAX (-3.4028235e38), I = -1 exactly as FAISS does.
 * Ties are ordered by ascending row id.  Ids are exact (certified against exact arithmetic),
 * scores are the correctly rounded fp32 of the exact inner product.
 * With host outputs the call returns after the results have landed; with device outputs it
 * is asynchronous on `stream`. */
int lxg_search(lxg_index* index, const float* x, int32_t nq, int32_t k, int normalize,
               float* D_out, int64_t* I_out, void* stream);

/* Same, plus optional exact fp64 scores D64_out [nq, k] (device memory) used to merge
 * row-shards without losing the order of scores that collide in fp32.  D64_out and I_out may be
 * the two planes of one packed [2, nq, k] 8-byte buffer - the unit lxg_merge_topk_packed consumes
 * after an all-gather. */
int lxg_search_ex(lxg_index* index, const float* x, int32_t nq, int32_t k, int normalize,
                  float* D_out, int64_t* I_out, double* D64_out, void* stream);

/* For callers of the asynchronous (device-output) form: waits for the last search on the handle
 * and reports what a synchronous call would have returned - LXG_ETIES if a query had more exact
 * ties with its k-th score than the exact path can hold, else LXG_OK; *uncertified (may be NULL)
 * receives the number of queries the exact path re-did. */
int lxg_index_sync(lxg_index* index, int32_t* uncertified);

/* In-place faiss.normalize_L2(x) on a device or host [nq, d] float32 matrix
 * (src/lean_explore/search/engine.py:242) for callers that want the normalised queries. */
int lxg_normalize_l2(float* x, int32_t nq, int32_t d, void* stream);

/* Merge step of a row-sharded index: Dg/Ig are the all-gathered per-shard results
 * [shards, nq, k] (exact fp64 scores, global ids, -1 padded; device memory); writes the global
 * top-k to D_out/I_out (device).  Every per-shard list must be ordered best first (score
 * descending, ties by ascending id, padding last) - what lxg_search_ex writes.  No reference
 * analogue (the reference is single-process); it is the exchange step of SURVEY.md section 8(e). */
int lxg_merge_topk(const double* Dg, const int64_t* Ig, int32_t nq, int32_t k, int32_t shards,
                   float* D_out, int64_t* I_out, void* stream);

/* Same merge, reading the all-gather buffer in place: `gathered` is [shards][2][nq][k] 8-byte
 * words, plane 0 of a shard = the bits of its fp64 scores (lxg_search_ex's D64_out), plane 1 =
 * its int64 ids (I_out) - each rank contributes one contiguous [2, nq, k] block, so the exchange
 * is exactly one all-gather with no repacking on either side. */
int lxg_merge_topk_packed(const int64_t* gathered, int32_t nq, int32_t k, int32_t shards,
                          float* D_out, int64_t* I_out, void* stream);

/* Counters of the last lxg_search on this handle (tests, bench.py). */
typedef struct lxg_search_stats {
  int32_t kernel_launches;   /* kernels launched by the call */
  int32_t slices;            /* corpus slices per query block */
  int32_t query_blocks;      /* blocks of 128 queries */
  int32_t kp;                /* candidates kept per (slice, query) */
  int32_t uncertified;       /* queries re-done by the exact path (-1 if not read back) */
  int32_t tile_rows;         /* corpus rows per tcgen05 accumulator tile */
} lxg_search_stats;
int lxg_index_last_stats(const lxg_index* index, lxg_search_stats* out);

/* Optional per-kernel device timing (bench.py's roofline numbers).  While enabled, every
 * lxg_search records CUDA events around its kernels on the stream it launches on;
 * lxg_index_get_timing synchronises them, returns the sums since the last get and resets. */
typedef struct lxg_timing {
  int32_t calls;
  float scan_ms;   /* pass 1: TMA + tcgen05 GEMM + threshold top-k' (scan_topk_kernel alone) */
  float merge_ms;  /* pass 2: merge + exact re-score + certificate */
  float exact_ms;  /* exact collectors for uncertified queries */
  float prep_ms;   /* query preparation: faiss.normalize_L2 + fp16 conversion (prep_queries_kernel) */
} lxg_timing;
int lxg_index_set_timing(lxg_index* index, int enable);
int lxg_index_get_timing(lxg_index* index, lxg_timing* out);

/* Test hook: raw tensor-core scores of pass 1 (scores_dev [nq, n] float32, scaled by
 * qscale_dev[q] * scan_scale) so the TMA/tcgen05 data path can be checked in isolation. */
int lxg_debug_scores(lxg_index* index, const float* x_dev, int32_t nq, int normalize,
                     float* scores_dev, float* qscale_dev, float* scan_scale_host, void* stream);

/* Test / measurement hook: selects code paths of pass 1 that the planner would otherwise pick by
 * shape (each argument: 0 / 1 sets, -1 leaves unchanged).  no_level: never use the cross-list
 * level (thresholds from list compaction only); force_single: never pair CTAs (cta_group::1
 * only); perf_mode: see LXG_SCAN_PERF_MODE in DESIGN.md (results are NOT produced when != 0).
 * perf_mode 16 / 17 only switch the three-stage merge of small batches off / on (results are
 * identical either way; LXG_MERGE_SPLIT=0 in the environment does the same).
 * The same switches are read from the environment (LXG_SCAN_NOLEVEL, LXG_SCAN_SINGLE,
 * LXG_SCAN_PERF_MODE) by lxg_init. */
int lxg_debug_config(int no_level, int force_single, int perf_mode);

/* How lxg_search would lay out pass 1 for a corpus shape and a batch (pure host arithmetic: needs no
 * device and no lxg_init; `sms` = streaming multiprocessors of the target GPU, 148 on a B200).  Exposes the
 * planning rules of DESIGN.md section 4.1 to the CPU test suite: query blocks of 128, corpus slices, two
 * candidate lists per slice, the tracker depth and the tracker ranks ("classes", each standing for
 * level_weight rows of a list) the level warps select over, list capacity and the merge pool. */
typedef struct lxg_plan_info {
  int32_t kp;               /* candidates certified per query: k plus the margin */
  int32_t query_blocks;     /* blocks of 128 queries */
  int32_t slices;           /* corpus slices per query block */
  int32_t lists;            /* candidate lists per query */
  int32_t tile_rows;        /* corpus rows per accumulator tile */
  int32_t pair;             /* 1: CTA pairs (tcgen05 cta_group::2) */
  int32_t level_depth;      /* tracker depth (0: no cross-list level, per-list compaction only) */
  int32_t level_classes;    /* tracker ranks read per list */
  int32_t level_rank[8];    /* 1-based rank of every class */
  int32_t level_weight[8];  /* rows of the list a class stands for */
  int32_t list_capacity;    /* entries per candidate list */
  int32_t merge_pool;       /* shared-memory pool of pass 2 (entries) */
} lxg_plan_info;
int lxg_debug_plan(int64_t n, int32_t d, int dtype, int32_t sms, int32_t nq, int32_t k, lxg_plan_info* out);

/* ---- sentence encoder (BERT-class) --------------------------------------------------
 * Replaces SentenceTransformer.encode inside EmbeddingClient.embed
 * (src/lean_explore/util/embedding_client.py:88-101): transformer forward -> pooling ->
 * L2 normalise.  Device pointers owned by the caller: matrices (nn.Linear weights [out, in]
 * row-major as HF BertModel stores them, embedding tables) are fp16; vectors (biases, LayerNorm
 * gamma / beta) are fp32.  hidden and ffn must be multiples of 128, head size even and <= 64. */
typedef struct lxg_bert_layer {
  const void *wqkv, *bqkv;   /* [3H, H], [3H]  (query|key|value stacked) */
  const void *wo, *bo;       /* [H, H], [H] */
  const void *ln1_g, *ln1_b; /* attention.output.LayerNorm */
  const void *w1, *b1;       /* intermediate.dense [F, H], [F] */
  const void *w2, *b2;       /* output.dense [H, F], [H] */
  const void *ln2_g, *ln2_b; /* output.LayerNorm */
} lxg_bert_layer;

typedef struct lxg_bert_weights {
  int32_t hidden, layers, heads, ffn, vocab, max_pos;
  float ln_eps;
  const void *word_emb, *pos_emb, *type_emb; /* [vocab,H], [max_pos,H], [2,H] */
  const void *emb_ln_g, *emb_ln_b;
  const lxg_bert_layer* layer;               /* host array of `layers` entries */
} lxg_bert_weights;

int lxg_encoder_create(lxg_encoder** out, const lxg_bert_weights* w);
int lxg_encoder_destroy(lxg_encoder* enc);
int lxg_encoder_last_launches(const lxg_encoder* enc); /* kernels launched by the last lxg_encode */
/* Calls of at most 64 tokens - what EmbeddingClient.embed sends for a search query
 * (search/engine.py:236) - run the whole forward as ONE cooperative kernel
 * (csrc/fused_encoder.cuh) when the geometry allows it (head size 32 / 64, ffn a multiple of
 * hidden); lxg_encoder_last_launches then reports 1.  enabled = 0 keeps this handle on the layered
 * kernels (A/B measurements, tests); the environment variable LXG_FUSED=0 does so process-wide. */
int lxg_encoder_set_fused(lxg_encoder* enc, int enabled);
/* Tracing aid: after lxg_encoder_set_fused(enc, 2) every single-kernel call records, per CTA and
 * phase (4 per layer + pooling), six %globaltimer stamps (ns) of its first thread: phase entered,
 * phase barrier passed, operand staged, accumulator ready, epilogue done, arrived.  out receives
 * grid x phases x 6 values (zero = the CTA had no job in that phase). */
int lxg_encoder_read_trace(lxg_encoder* enc, uint64_t* out, int32_t capacity, int32_t* grid, int32_t* phases);
/* ids/mask: [b, s] int32 token ids / attention mask (host or device); out: [b, H] float32
 * unit vectors (host or device).  pool = LXG_POOL_MEAN | LXG_POOL_CLS. */
int lxg_encode(lxg_encoder* enc, const int32_t* ids, const int32_t* mask, int32_t b, int32_t s,
               int pool, float* out, void* stream);

/* ---- Qwen3-class decoder backbone: embedding model and reranker ------------------------
 * The two models the reference ships with share one backbone (RMSNorm, rotary positions,
 * grouped-query causal attention with per-head q/k RMSNorm, SwiGLU):
 *   lxg_decoder_embed  replaces SentenceTransformer("Qwen/Qwen3-Embedding-0.6B").encode inside
 *                      EmbeddingClient.embed (src/lean_explore/util/embedding_client.py:58,88-101):
 *                      transformer -> Pooling(lasttoken) -> Normalize; out [b, hidden] float32.
 *   lxg_decoder_rerank replaces the AutoModelForCausalLM forward and the true/false softmax of
 *                      RerankerClient._compute_scores_sync
 *                      (src/lean_explore/util/reranker_client.py:110-141): scores [b] float32 =
 *                      softmax([logit_false, logit_true])[1] of the LAST position's logits.
 * ids / mask are [b, s] int32 (host or device), padded on the left as both reference clients
 * do (padding_side="left"); right padding is handled too (the last position whose mask is set
 * is used).  Positions are 0..s-1 over the padded sequence exactly as transformers.Qwen3Model
 * assigns them when no position_ids are passed.
 * Device pointers owned by the caller: matrices fp16 row-major [out, in] as nn.Linear stores
 * them, vectors fp32.  head_dim must be 128, hidden a multiple of 128 and <= 1024, ffn a
 * multiple of 64. */
typedef struct lxg_qwen3_layer {
  const void* ln1;      /* input_layernorm.weight [H] */
  const void* wqkv;     /* [(heads + 2 kv_heads) * 128, H]: q_proj | k_proj | v_proj stacked */
  const void* q_norm;   /* self_attn.q_norm.weight [128] */
  const void* k_norm;   /* self_attn.k_norm.weight [128] */
  const void* wo;       /* o_proj [H, heads * 128] */
  const void* ln2;      /* post_attention_layernorm.weight [H] */
  const void* wgu;      /* [2 F, H]: gate_proj / up_proj rows interleaved in groups of 32
                           (rows 64 i .. 64 i + 31 = gate[32 i ..], rows 64 i + 32 .. = up[32 i ..]) */
  const void* wdown;    /* down_proj [H, F] */
} lxg_qwen3_layer;

typedef struct lxg_qwen3_weights {
  int32_t hidden, layers, heads, kv_heads, head_dim, ffn, vocab;
  float rms_eps;
  const void* tok_emb;    /* embed_tokens [vocab, H] fp16 */
  const void* lm_head;    /* [vocab, H] fp16 (== tok_emb when tied); NULL for embedding-only use */
  const void* final_norm; /* norm.weight [H] fp32 */
  const void* inv_freq;   /* rotary inv_freq [64] fp32, as Qwen3RotaryEmbedding computes it */
  const lxg_qwen3_layer* layer; /* host array of `layers` entries */
} lxg_qwen3_weights;

typedef struct lxg_decoder lxg_decoder;
int lxg_decoder_create(lxg_decoder** out, const lxg_qwen3_weights* w);
int lxg_decoder_destroy(lxg_decoder* dec);
int lxg_decoder_last_launches(const lxg_decoder* dec);
/* tokens the last forward computed: for a host batch of >= 512 padded tokens of which >= 10 % are
 * padding, the padding tokens are dropped before the first layer (sequences are packed back to
 * back), so this is sum(mask), not b * s */
int lxg_decoder_last_tokens(const lxg_decoder* dec);
int lxg_decoder_embed(lxg_decoder* dec, const int32_t* ids, const int32_t* mask, int32_t b, int32_t s,
                      float* out, void* stream);
int lxg_decoder_rerank(lxg_decoder* dec, const int32_t* ids, const int32_t* mask, int32_t b, int32_t s,
                       int32_t token_true, int32_t token_false, float* scores, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LXG_H_ */
