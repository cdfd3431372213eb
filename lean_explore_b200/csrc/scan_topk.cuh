// Pass 1 of the flat inner-product search (replaces faiss.normalize_L2 + Index.search,
// reference call sites: src/lean_explore/search/engine.py:242 and :250).
//
// prep_queries_kernel (one warp per query): the fused L2-normalise (FAISS fvec_renorm_L2
// semantics, engine.py:242), a power-of-two scale and the fp16 conversion of the query block.
//
// scan_topk_kernel: one persistent CTA per (corpus slice, block of 128 queries), 12 warps:
//   * warp 8: TMA producer.  Streams the slice of the fp16 corpus ("scan copy") through a ring
//     of 128B-swizzled shared-memory stages (N_T rows x 64 dims each) - all ~192 KB of shared
//     memory is corpus pipeline.
//   * warp 9: one elected thread issues tcgen05.mma (M=128 queries, N=N_T corpus rows, K=16),
//     A (the fp16 queries) from TENSOR MEMORY, B from the swizzled stage, fp32 accumulators in
//     TMEM, double buffered.
//   * warps 0-7: epilogue, two groups of four.  Group g owns half of the columns of every
//     accumulator tile; thread t owns query t of the block: tcgen05.ld 32 scores at a time (lane =
//     query, column = corpus row), max-tree per 8 columns against the thread's threshold; survivors
//     are appended to the (slice, group, query) candidate list in global memory (L2 resident) and
//     feed a register tracker of the best scores the list has seen, which the thread publishes.
//   * warps 10-11: LEVEL warps.  Thresholds come from a CROSS-LIST LEVEL: the kp-th largest of the
//     union of all lists' published trackers (at least kp rows score >= it, so nothing below it can
//     be among the query's best kp).  The CTAs that scan a query block deal its queries among their
//     level warps; each selects the level of its queries over and over, off the epilogue's critical
//     path, and the epilogue threads pick the latest value up from shared memory at every tile.
//     Without a level (too few lists for trackers of 8) a list that fills up is compacted by its warp
//     to its best kp entries (exact k-th-largest by bit bisection) and that raises the threshold.
// Pass 2 (rescore.cuh) merges the lists, re-scores the survivors exactly and certifies that the
// answer equals the exact top-k.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <type_traits>
#include "ptx.cuh"

namespace lxg {

// Epilogue warp groups.  Measured on B200 (same box, r1p): 2 groups (8 epilogue warps, 64 columns
// each at N_T = 128) 0.336 ms on cfg2 / 2.61 ms on cfg3; 4 groups (16 warps, 32 columns each, 96
// registers per thread, twice the lists) 0.394 / 2.83 ms.  More warps do not help: the kernel sits
// at the chip's power wall and the extra per-tile bookkeeping costs more than the latency it hides.
#ifndef LXG_EPI_GROUPS
#define LXG_EPI_GROUPS 2
#endif
constexpr int kGroups = LXG_EPI_GROUPS;   // epilogue warp groups: each owns N_T / kGroups columns of every tile
constexpr int kEpiWarps = 4 * kGroups;    // four warps (the four TMEM lane quarters) per group
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kScanThreads = kEpiThreads + 128;  // + one TMA warp + one MMA warp + two level warps
constexpr int kKC = 64;            // fp16 elements per 128-byte swizzled row
constexpr int kStageRing = 196608; // bytes of shared memory used as corpus pipeline (kASm == 0)
constexpr int kScanSmemMax = 229376;  // A chunks in shared memory + pipeline when kASm > 0
constexpr int kAChunkBytes = 16384;   // one k-chunk of the query block: 128 rows x 128 bytes
constexpr int kStageBytes = 32768; // one pipeline stage: N_T rows x (32768 / (128 N_T)) k-chunks
constexpr int kQueryBlock = 128;   // queries per CTA == UMMA M
constexpr int kTrack = 8;          // register tracker: best kTrack scores a list has seen
constexpr uint32_t kLvlNone = 0u;            // no level (yet): ordered keys of real scores are never 0
constexpr uint32_t kLvlSkip = 0xFFFFFFFFu;   // (pass 2) identity of the min over level slots
constexpr uint32_t kKeyNegInf = 0x007FFFFFu; // float_to_key(-inf): an empty tracker slot
constexpr int kLvlMaxWords = 1280;           // tracker words per query the level warp selects over (40 per lane)

struct ScanParams {
  const __half* xh;   // [query blocks * 128, dpad] prepared queries (normalised, scaled, fp16, zero padded)
  uint2* cand;        // [lists, nq, cap] (score bits, row);  list = slice * kGroups + epilogue group
  int* cand_count;    // [lists, nq]
  float* slice_thr;   // [lists, nq] final threshold of the list: every dropped row scored <= it
  float* dbg_scores;  // optional [nq, n] raw tensor-core scores (tests only), else nullptr
  uint32_t* lvl;      // [nq] cross-list level of every query (ordered key, 0 = none yet), raised by the level warps
  uint32_t* trk;      // [nq, lists, kTrack] published trackers: ordered keys of the best scores each list has seen
  int lvl_r;          // tracker depth (1, 2, 4 or 8 real slots); lists * lvl_r >= kp; 0 disables the cross-list level
  int lvl_lg;         // log2 of the number of tracker ranks ("classes") per list the level warp reads
  uint32_t lvl_slot;  // class c reads tracker slot (lvl_slot >> 4c) & 15 ...
  uint32_t lvl_w;     // ... which stands for (lvl_w >> 4c) & 15 rows of the list
  int lists;          // slices * kGroups
  int lvl_sleep_ns;
  unsigned long long* lvl_dbg;  // level-warp diagnostics (LXG_DEBUG_COUNTS), else nullptr
  int nq, dpad, num_kc;
  int n;
  int num_tiles, slices, tiles_per_slice;
  int kp, cap, keep_max;
  int* progress;      // [slices, readers] tiles issued by every reader (CTA or CTA pair) of a slice, or nullptr
  int sync_window;    // a reader does not run more than this many tiles ahead of the slowest one
  int perf_mode;      // measurements only (LXG_SCAN_PERF_MODE): 1 = epilogue releases tiles unread, 2 = no candidate ever passes, 3 = TMEM reads only
};

__device__ __forceinline__ uint32_t float_to_key(uint32_t b) {
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ uint32_t key_to_float_bits(uint32_t k) {
  return (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
}

// ||x||^2 of one fp32 row by a full warp, accumulated in fp64 in a fixed order (lane-strided
// partial sums, xor tree), rounded once to fp32: the `nr` of FAISS' fvec_renorm_L2.  Shared by
// prep_queries_kernel and normalize_l2_kernel so both produce bit-identical normalised queries.
__device__ __forceinline__ float warp_row_norm_sq(const float* __restrict__ xr, int d, int lane) {
  double s = 0.0;
  for (int i = lane; i < d; i += 32) {
    const double a = xr[i];
    s = fma(a, a, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return static_cast<float>(s);
}
// inv_nr = 1.0 / sqrtf(nr) evaluated in double and rounded to float, as FAISS writes it.
__device__ __forceinline__ float inv_norm(float nr) {
  return nr > 0.0f ? static_cast<float>(1.0 / static_cast<double>(sqrtf(nr))) : 1.0f;
}

// In-place faiss.normalize_L2 (engine.py:242): one warp per row.
__global__ void __launch_bounds__(256) normalize_l2_kernel(float* __restrict__ x, int nq, int d) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  float* xr = x + static_cast<size_t>(q) * d;
  const float inv = inv_norm(warp_row_norm_sq(xr, d, lane));
  for (int i = lane; i < d; i += 32) xr[i] = xr[i] * inv;
}

// Query preparation, one warp per row of the padded query block matrix:
//   xn  = normalize_L2(x) (fp32, what the exact re-score uses),
//   xh  = fp16(xn * 2^e), e chosen so that max|xn * 2^e| lands in [1,2): exact scaling, keeps
//         fp16 out of the subnormals; rows >= nq and columns >= d are zero,
//   qscale = 2^e, qnorm = ||xn||_2 rounded up (used only in the certificate's error bound).
// It also clears the per-search counters (zero_a / zero_b: published levels, flag and exact-list
// counters) so that a search needs no memset nodes on its stream.
__global__ void __launch_bounds__(256)
prep_queries_kernel(const float* __restrict__ x, float* __restrict__ xn, __half* __restrict__ xh,
                    float* __restrict__ qscale, float* __restrict__ qnorm, int nq, int nq_pad, int d,
                    int dpad, int normalize, uint32_t* __restrict__ zero_a, int zero_a_words,
                    uint32_t* __restrict__ zero_b, int zero_b_words) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < zero_a_words + zero_b_words; i += gridDim.x * blockDim.x) {
    if (i < zero_a_words) zero_a[i] = 0u; else zero_b[i - zero_a_words] = 0u;
  }
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= nq_pad) return;
  __half* hr = xh + static_cast<size_t>(q) * dpad;
  if (q >= nq) {
    for (int i = lane; i < dpad; i += 32) hr[i] = __float2half_rn(0.0f);
    return;
  }
  const float* xr = x + static_cast<size_t>(q) * d;
  float* nr = xn + static_cast<size_t>(q) * d;
  const float inv = normalize ? inv_norm(warp_row_norm_sq(xr, d, lane)) : 1.0f;
  float amax = 0.0f;
  double n2 = 0.0;
  for (int i = lane; i < d; i += 32) {
    const float v = xr[i] * inv;
    nr[i] = v;
    amax = fmaxf(amax, fabsf(v));
    n2 = fma(static_cast<double>(v), static_cast<double>(v), n2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    n2 += __shfl_xor_sync(0xffffffffu, n2, o);
  }
  float sq = 1.0f;
  if (amax > 0.0f && amax < CUDART_INF_F) sq = ldexpf(1.0f, -ilogbf(amax));
  for (int i = lane; i < dpad; i += 32) hr[i] = __float2half_rn(i < d ? xr[i] * inv * sq : 0.0f);
  if (lane == 0) {
    qscale[q] = sq;
    qnorm[q] = static_cast<float>(sqrt(n2)) * 1.0000002f;
  }
}

// Warp-cooperative compaction of one candidate list (c > kp entries on entry).
// Finds a cut key by bit bisection, starting at the highest bit in which the keys differ.
//   keep_max == kp : exact - keeps the kp largest scores (ties: first in list order == lowest
//                    row) and returns the kp-th largest score;
//   keep_max >  kp : the bisection stops as soon as kp <= #(key >= cut) <= keep_max and keeps all
//                    of those (cheaper: no tie handling, ~3x fewer steps).
// Returns the cut score (every dropped entry scores <= it); *kept = entries left in the list.
template <int PER_LANE>
__device__ __forceinline__ float warp_compact_regs(uint2* __restrict__ buf, int c, int kp, int keep_max,
                                                   int lane, int* kept) {
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint2 e[PER_LANE];
  uint32_t key[PER_LANE];
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
#pragma unroll
  for (int m = 0; m < PER_LANE; ++m) {
    const int i = m * 32 + lane;
    if (i < c) {
      e[m] = __ldcg(buf + i);
      key[m] = float_to_key(e[m].x);
      kmin = min(kmin, key[m]);
      kmax = max(kmax, key[m]);
    } else {
      e[m] = make_uint2(0, 0);
      key[m] = 0;  // below every real key (float_to_key never returns 0 for a finite score)
    }
  }
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  uint32_t prefix = kmax;
  bool exact = true;  // prefix is the exact kp-th largest key (tie handling needed)
  if (kmin != kmax) {
    const int hb = 31 - __clz(kmin ^ kmax);
    prefix = (hb == 31) ? 0u : (kmax & ~((2u << hb) - 1u));
    for (int bit = hb; bit >= 0; --bit) {
      const uint32_t cnd = prefix | (1u << bit);
      int mine = 0;
#pragma unroll
      for (int m = 0; m < PER_LANE; ++m) mine += (key[m] >= cnd) ? 1 : 0;
      const int ge = __reduce_add_sync(0xffffffffu, mine);
      if (ge >= kp) {
        prefix = cnd;
        if (ge <= keep_max && keep_max > kp) {
          exact = false;
          break;
        }
      }
    }
  }
  int gt = 0;
#pragma unroll
  for (int m = 0; m < PER_LANE; ++m) gt += (key[m] > prefix) ? 1 : 0;
  const int need_eq = exact ? kp - __reduce_add_sync(0xffffffffu, gt) : 0x7fffffff;
  int w = 0, eq_seen = 0;
#pragma unroll
  for (int m = 0; m < PER_LANE; ++m) {
    const bool is_eq = (key[m] == prefix) && (m * 32 + lane < c);
    const uint32_t be = __ballot_sync(0xffffffffu, is_eq);
    const bool keep = (key[m] > prefix) || (is_eq && (eq_seen + __popc(be & lt_mask) < need_eq));
    const uint32_t bk = __ballot_sync(0xffffffffu, keep);
    if (keep) __stcg(buf + w + __popc(bk & lt_mask), e[m]);
    w += __popc(bk);
    eq_seen += __popc(be);
  }
  *kept = w;
  return __uint_as_float(key_to_float_bits(prefix));
}

__device__ __noinline__ float warp_compact(uint2* __restrict__ buf, int c, int kp, int keep_max, int lane,
                                           int* kept) {
  __syncwarp();
  float r;
  if (c <= 128) {
    r = warp_compact_regs<4>(buf, c, kp, keep_max, lane, kept);
  } else if (c <= 256) {
    r = warp_compact_regs<8>(buf, c, kp, keep_max, lane, kept);
  } else if (c <= 384) {
    r = warp_compact_regs<12>(buf, c, kp, keep_max, lane, kept);
  } else {
    // large k: keys stay in memory, always exact
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t prefix = 0;
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cnd = prefix | (1u << bit);
      int mine = 0;
      for (int i = lane; i < c; i += 32) mine += (float_to_key(__ldcg(buf + i).x) >= cnd) ? 1 : 0;
      if (__reduce_add_sync(0xffffffffu, mine) >= kp) prefix = cnd;
    }
    int gt = 0;
    for (int i = lane; i < c; i += 32) gt += (float_to_key(__ldcg(buf + i).x) > prefix) ? 1 : 0;
    const int need_eq = kp - __reduce_add_sync(0xffffffffu, gt);
    int w = 0, eq_seen = 0;
    for (int base = 0; base < c; base += 32) {
      const int i = base + lane;
      uint2 e = make_uint2(0, 0);
      uint32_t key = 0;
      if (i < c) {
        e = __ldcg(buf + i);
        key = float_to_key(e.x);
      }
      const bool is_eq = (i < c) && (key == prefix);
      const uint32_t be = __ballot_sync(0xffffffffu, is_eq);
      const bool keep = (key > prefix) || (is_eq && (eq_seen + __popc(be & lt_mask) < need_eq));
      const uint32_t bk = __ballot_sync(0xffffffffu, keep);
      __syncwarp();  // every lane has read chunk `base` before anyone overwrites [w, w+32) <= base+32
      if (keep) __stcg(buf + w + __popc(bk & lt_mask), e);
      w += __popc(bk);
      eq_seen += __popc(be);
      __syncwarp();
    }
    *kept = w;
    r = __uint_as_float(key_to_float_bits(prefix));
  }
  __syncwarp();
  return r;
}

// State of one epilogue thread's candidate list.
struct ListState {
  uint2* buf;
  uint2* wp;  // next free entry
  float thr;  // scores <= thr are dropped
  __device__ __forceinline__ int count() const { return static_cast<int>(wp - buf); }
};

// Compacts the lists of every lane whose list is fuller than `limit` (warp-cooperative).
__device__ __forceinline__ void compact_full_lists(ListState& ls, int limit, int kp, int keep_max, int lane) {
  uint32_t need = __ballot_sync(0xffffffffu, ls.count() > limit);
  while (need) {
    const int src = __ffs(need) - 1;
    need &= need - 1;
    uint2* b = reinterpret_cast<uint2*>(
        __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(ls.buf), src));
    const int bc = __shfl_sync(0xffffffffu, ls.count(), src);
    int kept;
    const float tnew = warp_compact(b, bc, kp, keep_max, lane, &kept);
    if (lane == src) {
      ls.thr = fmaxf(ls.thr, tnew);  // a level may already have raised it above the list's cut
      ls.wp = ls.buf + kept;
    }
  }
}

// ---- register tracker: the kTrack best scores a list has seen (sorted, t[0] best).  It starts
// with kTrack - r entries of +inf, so t[kTrack-1] is always the r-th best REAL score (no dynamic
// register indexing).  It may miss scores (it is fed the maximum of each 8-column group that
// produced a candidate): the r-th best of a subset is still a valid lower bound of the r-th best
// of the list.
__device__ __forceinline__ void track_insert(float (&t)[kTrack], float v, bool two_slots) {
  if (v > t[kTrack - 1]) {
    // t is sorted descending: new t[k] = min(t[k-1], max(v, t[k])) - every slot independently
    if (two_slots) {  // r <= 2 (warp-uniform): the other slots hold the +inf sentinels
      const float hi = fmaxf(v, t[kTrack - 2]);
      t[kTrack - 1] = fminf(t[kTrack - 2], fmaxf(v, t[kTrack - 1]));
      t[kTrack - 2] = hi;
    } else {
      float nt[kTrack];
      nt[0] = fmaxf(v, t[0]);
#pragma unroll
      for (int k = 1; k < kTrack; ++k) nt[k] = fminf(t[k - 1], fmaxf(v, t[k]));
#pragma unroll
      for (int k = 0; k < kTrack; ++k) t[k] = nt[k];
    }
  }
}

// ---- cross-list level.  Every list publishes its tracker (kTrack ordered keys, best
// first behind the sentinels) after each tile that changed it.  If the published values of ALL lists
// of a query hold at least kp entries >= L - counting a rank-r slot as r rows of its list - then at
// least kp corpus rows score >= L and no row scoring <= L can be among the query's best kp, whichever
// slice it lives in.  The largest such L is the kp-th largest of the (weighted) union.  A slot is
// monotone in time, so a reader that sees a mix of old and new slots of one list only undercounts.
//
// The selection runs on a dedicated LEVEL WARP per CTA, off the epilogue's critical path: the CTAs
// that scan one query block (one per slice) deal the block's queries among themselves, each computes
// the level of its queries and raises p.lvl[q]; every level warp then copies the block's 128 levels
// into shared memory, where the epilogue threads pick them up at their next tile.
//
// One query: v = lists << lg words; word i is rank class (i & (2^lg - 1)) of list i >> lg.  i = m * 32 +
// lane, so a lane's class - its tracker slot and the rows it stands for - does not depend on m.
//
// Selection = bisection on the ordered keys, from the highest bit in which the query's keys differ
// down to bit kLvlLowBit (the result is the kp-th largest rounded DOWN - by less than 0.1 % of its
// value - still a valid level), stopping early once at most ~12 % more than kp entries reach the
// candidate.  Registers only: the SM's shared
// memory bandwidth belongs to the tensor cores.  NQ queries run interleaved (the chain of compare /
// warp-reduce steps is latency, not work).
constexpr int kLvlWarps = 2;
constexpr int kLvlLowBit = 13;
// rec: tracker words of the first query, qstride words between queries.  out[u] = ordered key of a
// valid level, 0 if there is none (or the query's lists have not all published and need_complete).
template <int PL, int NQ>
__device__ __noinline__ void level_select(const uint32_t* __restrict__ rec, size_t qstride, int nqv, int v, int lg,
                                             int slot_lane, int w_lane, int kp, int lane, bool need_complete,
                                             uint32_t (&out)[NQ]) {
  uint32_t key[NQ][PL];
  uint32_t prefix[NQ];
  bool done[NQ];
  int nvalid = 0;
#pragma unroll
  for (int m = 0; m < PL; ++m) nvalid += (m * 32 + lane < v) ? 1 : 0;
  const bool enough = __reduce_add_sync(0xffffffffu, nvalid * w_lane) >= kp;
#pragma unroll
  for (int u = 0; u < NQ; ++u) {
#pragma unroll
    for (int m = 0; m < PL; ++m) {
      const int i = m * 32 + lane;
      key[u][m] = (u < nqv && i < v) ? __ldcg(rec + u * qstride + ((i >> lg) << 3) + slot_lane) : 0u;
    }
  }
  int hbmax = -1;
#pragma unroll
  for (int u = 0; u < NQ; ++u) {
    uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
#pragma unroll
    for (int m = 0; m < PL; ++m) {
      if (m * 32 + lane < v) {
        kmin = min(kmin, key[u][m]);
        kmax = max(kmax, key[u][m]);
      }
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    // kmin == 0: some list has not published yet
    const bool act = enough && u < nqv && kmax > kKeyNegInf && (kmin != 0u || !need_complete);
    prefix[u] = act ? kmax : 0u;  // all keys equal: the level is that key
    done[u] = !act || kmin == kmax;
    if (!done[u]) {
      const int hb = 31 - __clz(kmin ^ kmax);
      prefix[u] = (hb == 31) ? 0u : (kmax & ~((2u << hb) - 1u));
      hbmax = max(hbmax, hb);
    }
  }
  // above a query's own highest differing bit a step changes nothing: the bit is either part of the
  // common prefix already (candidate == prefix) or above every key (count 0)
  for (int bit = hbmax; bit >= kLvlLowBit; --bit) {
    bool all_done = true;
#pragma unroll
    for (int u = 0; u < NQ; ++u) {
      const uint32_t cnd = prefix[u] | (1u << bit);
      int mine = 0;
#pragma unroll
      for (int m = 0; m < PL; ++m) mine += (key[u][m] >= cnd) ? 1 : 0;
      const int ge = __reduce_add_sync(0xffffffffu, mine * w_lane);
      if (!done[u] && ge >= kp) {
        prefix[u] = cnd;
        done[u] = ge <= kp + (kp >> 3);
      }
      all_done = all_done && done[u];
    }
    if (all_done) break;
  }
#pragma unroll
  for (int u = 0; u < NQ; ++u) out[u] = prefix[u] > kKeyNegInf ? prefix[u] : 0u;
}

// The cheap level of NQ queries (stride qstride words): the minimum over the lists of the tracker
// slot that holds their r-th best, lists * r >= kp (that many rows reach it).  0 while some list has
// not published.  All loads are issued before the first reduction: one L2 round trip per call.
template <int PL, int NQ>
__device__ __noinline__ void level_min(const uint32_t* __restrict__ rec, size_t qstride, int nqv, int lists, int slot,
                                          int lane, uint32_t (&out)[NQ]) {
  uint32_t kmin[NQ];
#pragma unroll
  for (int u = 0; u < NQ; ++u) {
    kmin[u] = 0xFFFFFFFFu;
#pragma unroll
    for (int m = 0; m < PL; ++m) {
      const int i = m * 32 + lane;
      if (u < nqv && i < lists) kmin[u] = min(kmin[u], __ldcg(rec + u * qstride + (i << 3) + slot));
    }
  }
#pragma unroll
  for (int u = 0; u < NQ; ++u) {
    const uint32_t r = __reduce_min_sync(0xffffffffu, kmin[u]);
    out[u] = (u < nqv && r > kKeyNegInf) ? r : 0u;
  }
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// A level warp's loop (see above).  which: 0..kLvlWarps-1; thr_sh: the block's 128 levels as floats
// (-inf = none yet); epi_done counts the epilogue warps that have finished their slice.
// A round = the selection over the union for every owned query, a few queries per call, then a copy of
// the block's levels to shared memory.  Until every query of the block has a level - the epilogue warps
// wait for that before they take anything from their first tile - and whenever the selection could do
// no better (lists * depth == kp), a round computes the CHEAP level instead / first: the minimum over
// the lists of their ceil(kp / lists)-th best, one L2 round trip for up to 16 queries.  (Measured on
// cfg2, Q = 1024: also running the cheap level in every later round costs 0.05 ms of a 0.32 ms scan.)
struct LevelArgs {
  unsigned long long* dbg;  // diagnostics: [0] rounds, [1] clocks spent selecting, [2] clocks alive (summed over level warps)
  uint32_t* lvl;
  const uint32_t* trk;
  int lists, lg, kp, nq, depth;
  uint32_t slot, w;
  int sleep_ns;  // experiment knob (LXG_LVL_SLEEP): pause between rounds; 0 = default schedule
};
__device__ __noinline__ void level_service(const LevelArgs p, int qblock, int slice, int slices, int which, float* thr_sh,
                                           volatile int* epi_done, int lane) {
  const int q0 = qblock * kQueryBlock;
  const int nlive = max(0, min(kQueryBlock, p.nq - q0));
  const int lg = p.lg;
  const int cls = lane & ((1 << lg) - 1);
  const int slot_lane = static_cast<int>((p.slot >> (4 * cls)) & 15u);
  const int w_lane = static_cast<int>((p.w >> (4 * cls)) & 15u);
  const int v = p.lists << lg;
  const int kp = p.kp;
  // the block's queries are dealt to the level warps of the CTAs that scan it
  const int owner = slice * kLvlWarps + which, owners = slices * kLvlWarps;
  const bool owned = owner < nlive;
  const size_t qwords = static_cast<size_t>(p.lists) * kTrack;
  // cheap level: rank r = ceil(kp / lists) of every list (lists * r >= kp rows reach the minimum)
  const int min_rank = (kp + p.lists - 1) / p.lists;
  const bool have_min = min_rank <= p.depth;
  const int min_slot = kTrack - p.depth + min_rank - 1;
  const bool select_useful = p.lists * p.depth > kp;  // else the selection IS the minimum
  const unsigned long long t0 = global_timer_ns();
  const long long c0 = clock64();
  long long csel = 0;
  int round = 0;
  bool have_all = false;  // every query of this warp's half of the block has had a level

  auto copy_levels = [&]() {
    bool all = true;
#pragma unroll
    for (int j = 0; j < kQueryBlock / 32 / kLvlWarps; ++j) {
      const int t = (which * (kQueryBlock / 32 / kLvlWarps) + j) * 32 + lane;
      if (t < nlive) {
        const uint32_t key = __ldcg(p.lvl + q0 + t);
        all = all && key != kLvlNone;
        if (key != kLvlNone) {
          const float f = __uint_as_float(key_to_float_bits(key));
          if (f > thr_sh[t]) thr_sh[t] = f;
        }
      }
    }
    have_all = have_all || __all_sync(0xffffffffu, all);
  };
  auto min_levels = [&](auto pl_tag, auto nq_tag) {
    constexpr int PL = decltype(pl_tag)::value, NQ = decltype(nq_tag)::value;
    for (int t = owner; t < nlive; t += owners * NQ) {  // warp-uniform
      uint32_t key[NQ];
      level_min<PL, NQ>(p.trk + (q0 + t) * qwords, owners * qwords, min(NQ, (nlive - t + owners - 1) / owners), p.lists,
                        min_slot, lane, key);
#pragma unroll
      for (int u = 0; u < NQ; ++u)
        if (lane == u && key[u] != 0u) atomicMax(p.lvl + q0 + t + u * owners, key[u]);
    }
  };
  auto select_levels = [&](auto pl_tag, auto nq_tag, bool need_complete) {
    constexpr int PL = decltype(pl_tag)::value, NQ = decltype(nq_tag)::value;
    for (int t = owner; t < nlive; t += owners * NQ) {  // warp-uniform
      uint32_t key[NQ];
      level_select<PL, NQ>(p.trk + (q0 + t) * qwords, owners * qwords, min(NQ, (nlive - t + owners - 1) / owners), v, lg,
                           slot_lane, w_lane, kp, lane, need_complete, key);
#pragma unroll
      for (int u = 0; u < NQ; ++u)
        if (lane == u && key[u] != 0u) atomicMax(p.lvl + q0 + t + u * owners, key[u]);
    }
  };

  for (;;) {
    const bool stop = *epi_done >= kEpiWarps;
    // The first levels wait (bounded) for every list's first tile: a level over a part of the lists
    // would let the first tiles dump most of their rows.  (The cheap level always needs all lists.)
    const bool early = !have_all && global_timer_ns() - t0 < 40000ull;
    const long long cs = clock64();
    if (owned) {
      if (have_min && (early || !select_useful)) {
        if (p.lists <= 64) min_levels(std::integral_constant<int, 2>{}, std::integral_constant<int, 16>{});
        else min_levels(std::integral_constant<int, 10>{}, std::integral_constant<int, 2>{});
        copy_levels();
      }
      if (select_useful) {
        if (v <= 160) select_levels(std::integral_constant<int, 5>{}, std::integral_constant<int, 4>{}, early);
        else if (v <= 320) select_levels(std::integral_constant<int, 10>{}, std::integral_constant<int, 4>{}, early);
        else if (v <= 640) select_levels(std::integral_constant<int, 20>{}, std::integral_constant<int, 2>{}, early);
        else select_levels(std::integral_constant<int, kLvlMaxWords / 32>{}, std::integral_constant<int, 1>{}, early);
      }
    }
    csel += clock64() - cs;
    ++round;
    copy_levels();
    have_all = have_all || clock64() - c0 > 150000;  // bounded like the first levels
    if (stop) break;
    // Rounds run back to back for the first ~35 us, then an eighth of the scan's age apart.  Early
    // on the level rises with every tile and its lag is paid in candidates (cfg2, Q = 1024: 0.29 ms
    // with no pause, 0.30 ms with 20 us pauses, 0.57 ms with 100 us); after thousands of tiles it
    // hardly moves, and on the long power-limited scans (cfg3 / cfg4) two warps that never rest cost
    // 3-4 % of the clock.
    if (have_all) {
      const long long alive = clock64() - c0;
      long long pause = p.sleep_ns > 0 ? static_cast<long long>(p.sleep_ns) * 2 : (alive < 65536 ? 400 : alive >> 3);
      for (; pause > 0 && *epi_done < kEpiWarps; pause -= 4000) __nanosleep(static_cast<unsigned>(min(pause, 4000ll)) >> 1);
    }
  }
  if (p.dbg != nullptr && lane == 0) {
    atomicAdd(p.dbg, static_cast<unsigned long long>(round));
    atomicAdd(p.dbg + 1, static_cast<unsigned long long>(csel));
    atomicAdd(p.dbg + 2, static_cast<unsigned long long>(clock64() - c0));
  }
}

__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// One chunk (32 or 16 columns) of the accumulator: per 8-column group a max tree against the
// threshold; a group with a survivor appends its survivors to the list and feeds the tracker.
template <bool kFeedTracker = true, int NC>
__device__ __forceinline__ void scan_chunk(const uint32_t (&r)[NC], ListState& ls, float (&tk)[kTrack],
                                           int base_row, bool two_slots) {
#pragma unroll
  for (int g = 0; g < NC / 8; ++g) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
    const float m = fmax3(fmax3(v[0], v[1], v[2]), fmax3(v[3], v[4], v[5]), fmaxf(v[6], v[7]));
    if (m > ls.thr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (v[j] > ls.thr) {
          __stcg(ls.wp, make_uint2(r[g * 8 + j], static_cast<uint32_t>(base_row + g * 8 + j)));
          ++ls.wp;
        }
      }
      if (kFeedTracker) track_insert(tk, m, two_slots);
    }
  }
}

// First tile of a list, pass 1: feed the tracker with every 8-column maximum, append nothing.
template <int NC>
__device__ __forceinline__ void track_chunk(const uint32_t (&r)[NC], float (&tk)[kTrack], bool two_slots) {
#pragma unroll
  for (int g = 0; g < NC / 8; ++g) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
    track_insert(tk, fmax3(fmax3(v[0], v[1], v[2]), fmax3(v[3], v[4], v[5]), fmaxf(v[6], v[7])), two_slots);
  }
}

// Cold variant (last, partial tile of the corpus; debug dump): columns >= valid are TMA zero fill.
template <int NC>
__device__ __forceinline__ void scan_chunk_careful(uint32_t taddr, ListState& ls, float (&tk)[kTrack], int base_row,
                                                   int valid, float* __restrict__ dbg_row, bool two_slots) {
  uint32_t r[NC];
  ptx::tmem_ld_cols(taddr, r);
  ptx::tc_wait_ld();
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    if (j < valid) {
      if (dbg_row != nullptr) dbg_row[base_row + j] = __uint_as_float(r[j]);
    } else {
      r[j] = 0xFF800000u;  // -inf
    }
  }
  scan_chunk<true>(r, ls, tk, base_row, two_slots);
}

// N_T: corpus rows per accumulator tile (UMMA N).  128 when the fp16 queries need <= 256 TMEM
// columns (d <= 512), 64 when they need up to 384 (d <= 768): A + 2 accumulators <= 512 columns.
// kASm: k-chunks (64 dims) of the query block kept in SHARED memory instead of tensor memory, for
// d beyond what fits TMEM next to two N_T = 128 accumulators (8 chunks = 512 dims): chunks
// 0..7 are TS MMAs (A from TMEM), chunks 8.. are SS MMAs (A from a 128B-swizzled K-major smem
// image written by the epilogue threads).  kASm = 8 covers d <= 1024 (the shipped
// Qwen3-Embedding-0.6B), kASm = 4 is an alternative to N_T = 64 for d <= 768.
// kPair: two CTAs of a cluster (two adjacent blocks of 128 queries) run ONE tcgen05.mma
// cta_group::2 stream (M = 256): each CTA stages only N_T/2 rows of every corpus tile, so a corpus
// byte crosses L2 -> SM once per 256 queries instead of once per 128.
template <int N_T, bool kPair, int kASm = 0>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_topk_kernel(const __grid_constant__ CUtensorMap tmap, const ScanParams p) {
  static_assert(kASm == 0 || N_T == 128, "shared-memory A chunks go with 128-row tiles");
  constexpr int kTmChunks = (512 - 2 * N_T) / 32;        // A k-chunks that live in tensor memory
  constexpr int kRingBytes = kASm == 0 ? kStageRing : kScanSmemMax - kASm * kAChunkBytes;
  constexpr int kBoxRows = kPair ? N_T / 2 : N_T;       // corpus rows this CTA stages per tile
  constexpr int kBoxBytes = kBoxRows * 128;             // one TMA box: kBoxRows rows x 64 fp16
  constexpr int kStageBytesT = kPair ? kStageBytes / 2 : kStageBytes;
  constexpr int kKcPerStage = kStageBytesT / kBoxBytes;  // k-chunks (boxes) per pipeline stage
  constexpr int kStages = kRingBytes / kStageBytesT;
  static_assert(kASm == 0 || kTmChunks % kKcPerStage == 0, "a stage is all-TS or all-SS");
  // Epilogue group g owns accumulator columns [g * kGroupCols, (g+1) * kGroupCols) of every tile,
  // read in chunks of kChunkCols columns.  (Computing every tile as N = N_T/2 halves with
  // per-group barriers was measured and dropped: N = 64 MMAs with A in tensor memory run at
  // ~2/3 of the N = 128 rate - 0.331 vs 0.270 ms for the bare cfg2 pipeline.)
  constexpr int kGroupCols = N_T / kGroups;
  constexpr int kChunkCols = kGroupCols < 32 ? kGroupCols : 32;
  constexpr int kGroupChunks = kGroupCols / kChunkCols;
  static_assert(kChunkCols == 32 || kChunkCols == 16, "tcgen05.ld shapes used: x32, x16");
  constexpr uint32_t kACol0 = 2 * N_T;  // TMEM columns: [0,N_T) acc0, [N_T,2N_T) acc1, then A
  constexpr uint32_t kIdesc = ptx::make_idesc_f16(kPair ? 256 : 128, N_T);
  constexpr uint32_t kAccArrivals = (kPair ? 2 : 1) * kEpiWarps;  // one arrival per epilogue warp
  constexpr uint32_t kAArrivals = (kPair ? 2 : 1) * kEpiWarps;    // every epilogue warp stores part of A
  constexpr int kTmaWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1, kLvlWarp = kEpiWarps + 2;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ __align__(8) uint64_t a_ready_bar;
  __shared__ uint32_t tmem_base_holder;
  __shared__ float thr_sh[kQueryBlock];  // per query: the latest cross-list level (written by the level warp)
  __shared__ int epi_done;               // epilogue warps that have finished their slice

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qblock = blockIdx.x;
  const int slice = blockIdx.y;
  const uint32_t rank = kPair ? ptx::cluster_ctarank() : 0u;  // position in the CTA pair

  // 1024-byte aligned: [A chunks kTmChunks.. (kASm x 16 KB)] [stage ring] (SWIZZLE_128B atoms are 1024 bytes).
  const uint32_t asm_u32 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ring_u32 = asm_u32 + kASm * kAChunkBytes;

  const int tile_begin = slice * p.tiles_per_slice;
  const int tile_end = min(p.num_tiles, tile_begin + p.tiles_per_slice);
  const int my_tiles = max(0, tile_end - tile_begin);

  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], kAccArrivals);
    }
    ptx::mbar_init(&a_ready_bar, kAArrivals);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < kQueryBlock) thr_sh[threadIdx.x] = -CUDART_INF_F;
  if (threadIdx.x == 0) epi_done = 0;
  if (warp == kTmaWarp) {
    if (lane == 0) ptx::prefetch_tensormap(&tmap);
    if constexpr (kPair) {
      ptx::tmem_alloc_pair(&tmem_base_holder, 512);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(&tmem_base_holder, 512);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if constexpr (kPair) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;

  if (warp == kTmaWarp) {
    // ------------------------------------------------------------ TMA producer
    // Pair mode: both CTAs load their half of the tile; the bytes of both halves are counted on
    // the even CTA's full barrier (it issues the MMAs), which therefore expects 2x the bytes.
    // The whole warp walks the loop (warp-uniform control flow); one elected lane issues.
    {
      const uint32_t full0 = ptx::opaque(ptx::smem_u32(&full_bar[0]));
      const uint32_t empty0 = ptx::opaque(ptx::smem_u32(&empty_bar[0]));
      const uint32_t ring0 = ptx::opaque(ring_u32);
      uint32_t stage = 0, phase = 0;
      // Every reader of a slice (the query blocks / pairs with the same blockIdx.y) streams the same
      // corpus rows; DRAM serves them once only while the readers stay within the L2's reach of each
      // other.  The leader therefore publishes its tile count every few tiles and holds its loads
      // while the slowest reader is more than sync_window tiles behind (bounded wait; all CTAs of a
      // launch are co-resident).  Measured on cfg3: without it the four pairs of a slice drift apart
      // and the corpus is read twice from DRAM.
      const int readers = static_cast<int>(gridDim.x) / (kPair ? 2 : 1);
      const int reader = static_cast<int>(blockIdx.x) / (kPair ? 2 : 1);
      int* const prog = (p.progress != nullptr && readers > 1 && rank == 0) ? p.progress + slice * readers : nullptr;
      const int check_every = max(1, p.sync_window >> 1);
      for (int t = tile_begin; t < tile_end; ++t) {
        if (prog != nullptr && (t - tile_begin) % check_every == 0) {
          const int mine = t - tile_begin;
          if (lane == 0) __stcg(prog + reader, mine);
          if (mine > p.sync_window) {
            const unsigned long long t0 = global_timer_ns();
            for (;;) {
              int v = lane < readers ? __ldcg(prog + lane) : 0x7fffffff;
              v = __reduce_min_sync(0xffffffffu, v);
              if (v + p.sync_window >= mine || global_timer_ns() - t0 > 20000ull) break;
              __nanosleep(200);
            }
          }
        }
        const int row = t * N_T + static_cast<int>(rank) * kBoxRows;
        for (int kc0 = 0; kc0 < p.num_kc; kc0 += kKcPerStage) {
          const int nb = min(kKcPerStage, p.num_kc - kc0);
          ptx::mbar_wait_a(empty0 + stage * 8, phase ^ 1u);
          if (ptx::elect_one()) {
            const uint32_t fb = full0 + stage * 8;
            if (rank == 0) ptx::mbar_arrive_expect_tx_a(fb, nb * kBoxBytes * (kPair ? 2 : 1));
            const uint32_t dst = ring0 + stage * kStageBytesT;
#pragma unroll
            for (int b = 0; b < kKcPerStage; ++b) {
              if (b < nb) {
                if constexpr (kPair)
                  ptx::tma_load_2d_pair_a(dst + b * kBoxBytes, &tmap, (kc0 + b) * kKC, row, fb, ptx::kEvictNormal);
                else
                  ptx::tma_load_2d_a(dst + b * kBoxBytes, &tmap, (kc0 + b) * kKC, row, fb, ptx::kEvictNormal);
              }
            }
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      if (prog != nullptr && lane == 0) __stcg(prog + reader, 0x3fffffff);  // done: never the slowest
    }
  } else if (warp == kMmaWarp) {
    // -------------------------------------------------------------- MMA issuer
    // (warp-uniform loop, one elected lane issues: in divergent code every tcgen05 instruction
    // would be wrapped in its own elect loop)
    if (rank == 0) {
      if constexpr (kPair && kASm > 0) ptx::mbar_wait_cluster(&a_ready_bar, 0); else ptx::mbar_wait(&a_ready_bar, 0);
      ptx::tc_fence_after();
      const uint32_t full0 = ptx::opaque(ptx::smem_u32(&full_bar[0]));
      const uint32_t empty0 = ptx::opaque(ptx::smem_u32(&empty_bar[0]));
      const uint32_t tfull0 = ptx::opaque(ptx::smem_u32(&tmem_full_bar[0]));
      const uint32_t tempty0 = ptx::opaque(ptx::smem_u32(&tmem_empty_bar[0]));
      // shared-memory matrix descriptor of (stage 0, box 0, k step 0); see make_kmajor_sw128_desc
      const uint32_t desc_lo0 = ptx::opaque(((ring_u32 & 0x3FFFFu) >> 4) | (1u << 16));
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_tmem0 = ptx::opaque(tmem_base + kACol0);
      const uint32_t adesc_lo0 = ptx::opaque(((asm_u32 & 0x3FFFFu) >> 4) | (1u << 16));
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const uint32_t acc = it & 1;
        ptx::mbar_wait_a(tempty0 + acc * 8, ((it >> 1) & 1) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * N_T;
        for (int kc0 = 0; kc0 < p.num_kc; kc0 += kKcPerStage) {
          const int nb = min(kKcPerStage, p.num_kc - kc0);
          ptx::mbar_wait_a(full0 + stage * 8, phase);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t lo = desc_lo0 + stage * (kStageBytesT >> 4);
            if (kASm == 0 || kc0 < kTmChunks) {
              const uint32_t a_kc = a_tmem0 + kc0 * 32;
#pragma unroll
              for (int b = 0; b < kKcPerStage; ++b) {
                if (b < nb) {
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4) {
                    // K advances 16 fp16 = 32 bytes inside the swizzle atom (+2 in >>4 units) and
                    // 8 TMEM columns in A.
                    const uint64_t bdesc =
                        (static_cast<uint64_t>(kDescHi) << 32) | (lo + b * (kBoxBytes >> 4) + k4 * 2);
                    const uint32_t accum = (kc0 | b | k4) != 0 ? 1u : 0u;
                    if constexpr (kPair)
                      ptx::mma_f16_ts_pair(d_tmem, a_kc + b * 32 + k4 * 8, bdesc, kIdesc, accum);
                    else
                      ptx::mma_f16_ts(d_tmem, a_kc + b * 32 + k4 * 8, bdesc, kIdesc, accum);
                  }
                }
              }
            } else {
              // A from shared memory: chunk kc lives at asm + (kc - kTmChunks) * 16 KB, same
              // K-major 128B-swizzled layout (and descriptor) as a corpus box
              const uint32_t alo = adesc_lo0 + (kc0 - kTmChunks) * (kAChunkBytes >> 4);
#pragma unroll
              for (int b = 0; b < kKcPerStage; ++b) {
                if (b < nb) {
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4) {
                    const uint64_t bdesc =
                        (static_cast<uint64_t>(kDescHi) << 32) | (lo + b * (kBoxBytes >> 4) + k4 * 2);
                    const uint64_t adesc =
                        (static_cast<uint64_t>(kDescHi) << 32) | (alo + b * (kAChunkBytes >> 4) + k4 * 2);
                    if constexpr (kPair)
                      ptx::mma_f16_ss_pair(d_tmem, adesc, bdesc, kIdesc, 1u);
                    else
                      ptx::mma_f16_ss(d_tmem, adesc, bdesc, kIdesc, 1u);
                  }
                }
              }
            }
            // stage reusable (in both CTAs) once these MMAs retire
            const bool last = kc0 + kKcPerStage >= p.num_kc;
            if constexpr (kPair) {
              ptx::tc_commit_pair_a(empty0 + stage * 8, 3);
              if (last) ptx::tc_commit_pair_a(tfull0 + acc * 8, 3);
            } else {
              ptx::tc_commit_a(empty0 + stage * 8);
              if (last) ptx::tc_commit_a(tfull0 + acc * 8);
            }
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp >= kLvlWarp) {
    // ------------------------------------------------------------- level warps
    if (p.lvl_r > 0) {
      const LevelArgs la{p.lvl_dbg, p.lvl, p.trk, p.lists, p.lvl_lg, p.kp, p.nq, p.lvl_r, p.lvl_slot, p.lvl_w, p.lvl_sleep_ns};
      level_service(la, qblock, slice, static_cast<int>(gridDim.y), warp - kLvlWarp, thr_sh, &epi_done, lane);
    }
  } else {
    // ------ epilogue warps: group g scans columns [g*kGroupCols, (g+1)*kGroupCols) of every
    // accumulator tile, one query per thread
    const int grp = warp >> 2;
    const int t = threadIdx.x & (kQueryBlock - 1);  // TMEM lane == query of the block
    const int q = qblock * kQueryBlock + t;
    const bool live = q < p.nq;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;

    // ---- A operand: this thread's prepared fp16 query row -> tensor memory (lane t, d/2 columns);
    // the 128-byte k-chunks of the row are dealt round robin to the groups
    {
      const uint4* xrow = reinterpret_cast<const uint4*>(p.xh + static_cast<size_t>(q) * p.dpad);
      for (int kc = grp; kc < min(p.num_kc, kASm > 0 ? kTmChunks : p.num_kc); kc += kGroups) {
        uint32_t r[32];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint4 v = __ldg(xrow + kc * 8 + u);
          r[4 * u] = v.x;
          r[4 * u + 1] = v.y;
          r[4 * u + 2] = v.z;
          r[4 * u + 3] = v.w;
        }
        ptx::tmem_st_32x32b_x32(tmem_base + lane_base + kACol0 + kc * 32, r);
      }
      if constexpr (kASm > 0) {
        // chunks kTmChunks..: K-major rows of 128 bytes, 16-byte unit u of row t stored at unit
        // u ^ (t & 7) (CU_TENSOR_MAP_SWIZZLE_128B, what the MMA descriptor expects)
        for (int kc = kTmChunks + grp; kc < p.num_kc; kc += kGroups) {
          const uint32_t rowaddr = asm_u32 + (kc - kTmChunks) * kAChunkBytes + t * 128;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint4 v = __ldg(xrow + kc * 8 + u);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + ((u ^ (t & 7)) << 4)), "r"(v.x), "r"(v.y),
                         "r"(v.z), "r"(v.w)
                         : "memory");
          }
        }
        ptx::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async proxy
      }
    }
    ptx::tc_wait_st();
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if constexpr (kPair && kASm > 0) ptx::mbar_arrive_cluster_release(&a_ready_bar, 0);  // orders the A image in shared memory
      else if constexpr (kPair) ptx::mbar_arrive_cluster(&a_ready_bar, 0);
      else ptx::mbar_arrive(&a_ready_bar);
    }

    // ---- threshold scan
    const int list_id = slice * kGroups + grp;
    const size_t list = static_cast<size_t>(list_id) * p.nq + (live ? q : 0);
    ListState ls;
    ls.buf = p.cand + list * p.cap;
    ls.wp = ls.buf;
    ls.thr = (live && p.perf_mode != 2) ? -CUDART_INF_F : CUDART_INF_F;
    float tk[kTrack];
#pragma unroll
    for (int i = 0; i < kTrack; ++i) tk[i] = (i < kTrack - p.lvl_r) ? CUDART_INF_F : -CUDART_INF_F;
    float published = -CUDART_INF_F;
    const bool two_slots = p.lvl_r <= 2;
    const int kp = p.kp, cap = p.cap, keep_max = p.keep_max, n = p.n, lvl_r = p.lvl_r;
    float* dbg_row = (p.dbg_scores != nullptr && live) ? p.dbg_scores + static_cast<size_t>(q) * n : nullptr;
    const bool dbg = p.dbg_scores != nullptr;
    // where this list publishes its tracker (query-major [nq][lists][kTrack]: the level warp reads one
    // query's trackers with coalesced loads)
    uint4* const trk_mine =
        (lvl_r > 0 && live) ? reinterpret_cast<uint4*>(p.trk + (static_cast<size_t>(q) * p.lists + list_id) * kTrack) : nullptr;
    auto publish = [&]() {
      if (trk_mine != nullptr && tk[kTrack - 1] > published) {
        published = tk[kTrack - 1];
        uint32_t w[kTrack];
#pragma unroll
        for (int i = 0; i < kTrack; ++i) w[i] = float_to_key(__float_as_uint(tk[i]));
        __stcg(trk_mine, make_uint4(w[0], w[1], w[2], w[3]));
        __stcg(trk_mine + 1, make_uint4(w[4], w[5], w[6], w[7]));
      }
    };
    const volatile float* const vthr = thr_sh;

    const uint32_t tfull0 = ptx::opaque(ptx::smem_u32(&tmem_full_bar[0]));
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t acc = it & 1;
      ptx::mbar_wait_a(tfull0 + acc * 8, (it >> 1) & 1);
      ptx::tc_fence_after();
      if (lvl_r > 0) ls.thr = fmaxf(ls.thr, vthr[t]);  // the level warp keeps raising it
      const int row0 = (tile_begin + it) * N_T + grp * kGroupCols;  // corpus row of the group's column 0
      const uint32_t tile_addr = tmem_base + lane_base + acc * N_T + grp * kGroupCols;
      if (it == 0 && lvl_r > 0 && lvl_r <= kGroupCols / 8 && !dbg && p.perf_mode == 0 && (tile_begin + 1) * N_T <= n) {
        // ---- first tile: instead of dumping all of it into the list (no threshold exists yet),
        // pass 1 only feeds the tracker and publishes it; then the warp waits (bounded: ~50 us, the
        // other CTAs are co-resident and do the same) until the level warps have turned every list's
        // first tile into the first cross-list level, and pass 2 re-reads the accumulator - still in
        // tensor memory - against it.  Lists stay ~4x shorter, which pass 2 of the search reads.
#pragma unroll 1
        for (int c = 0; c < kGroupChunks; ++c) {
          uint32_t r[kChunkCols];
          ptx::tmem_ld_cols(tile_addr + c * kChunkCols, r);
          ptx::tc_wait_ld();
          track_chunk(r, tk, two_slots);
        }
        publish();
        const unsigned long long t0 = global_timer_ns();
        while (!__all_sync(0xffffffffu, !live || vthr[t] > -CUDART_INF_F) && global_timer_ns() - t0 < 50000ull)
          __nanosleep(100);
        ls.thr = fmaxf(ls.thr, vthr[t]);
#pragma unroll 1
        for (int c = 0; c < kGroupChunks; ++c) {
          uint32_t r[kChunkCols];
          ptx::tmem_ld_cols(tile_addr + c * kChunkCols, r);
          ptx::tc_wait_ld();
          scan_chunk<false>(r, ls, tk, row0 + c * kChunkCols, two_slots);
        }
      } else if (p.perf_mode == 1) {
      } else if (dbg || (tile_begin + it + 1) * N_T > n) {  // warp-uniform
#pragma unroll 1
        for (int c = 0; c < kGroupChunks; ++c) {
          const int base_row = row0 + c * kChunkCols;
          if (base_row < n)
            scan_chunk_careful<kChunkCols>(tile_addr + c * kChunkCols, ls, tk, base_row, min(kChunkCols, n - base_row),
                                           dbg_row, two_slots);
        }
      } else if constexpr (kGroupChunks == 1) {
        uint32_t r[kChunkCols];
        ptx::tmem_ld_cols(tile_addr, r);
        ptx::tc_wait_ld();
        if (p.perf_mode != 3) scan_chunk<true>(r, ls, tk, row0, two_slots);
      } else {
        // software pipeline over the group's chunks: chunk c+1 is in flight while c is compared
        uint32_t r[2][kChunkCols];
        ptx::tmem_ld_cols(tile_addr, r[0]);
#pragma unroll
        for (int c = 0; c < kGroupChunks; ++c) {
          ptx::tc_wait_ld();
          if (c + 1 < kGroupChunks) ptx::tmem_ld_cols(tile_addr + (c + 1) * kChunkCols, r[(c + 1) & 1]);
          if (p.perf_mode != 3) scan_chunk<true>(r[c & 1], ls, tk, row0 + c * kChunkCols, two_slots);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // one arrival per warp (in pair mode the odd CTA's arrivals are remote)
        if constexpr (kPair) ptx::mbar_arrive_cluster(&tmem_empty_bar[acc], 0); else ptx::mbar_arrive(&tmem_empty_bar[acc]);
      }
      publish();
      // a tile appends at most kGroupCols entries per list: keep room for the next one
      compact_full_lists(ls, cap - N_T, kp, keep_max, lane);
    }
    // ---- without a level the pass-2 merge expects at most kp entries per list
    if (lvl_r == 0) compact_full_lists(ls, kp, kp, kp, lane);
    if (live) {
      p.cand_count[list] = ls.count();
      p.slice_thr[list] = ls.thr;
    }
    __syncwarp();
    if (lane == 0) atomicAdd(&epi_done, 1);
  }

  ptx::tc_fence_before();
  if constexpr (kPair) ptx::cluster_sync(); else __syncthreads();
  if (warp == kTmaWarp) {
    ptx::tc_fence_after();
    if constexpr (kPair) ptx::tmem_dealloc_pair(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace lxg
