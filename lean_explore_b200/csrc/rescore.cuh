// Pass 2 of the flat inner-product search: per query, merge the per-slice candidate lists of
// pass 1, re-score the best kp exactly (fp64 accumulation of exact fp32 x fp16/fp32 products),
// order them (score desc, row asc - the tie rule of FAISS' CMin heap, see oracle/faiss_flat.py),
// write the (D float32[nq,k], I int64[nq,k]) result that faiss.Index.search returns
// (reference call site src/lean_explore/search/engine.py:250) and CERTIFY it: the result is
// provably the exact top-k iff the k-th exact score beats every possible score of a dropped
// row (a_min + eps).  Uncertified queries (near-duplicate rows around rank k) are re-done by
// the exact collectors at the end of this file, so ids are exact in every case.
#pragma once
#include <cfloat>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include "scan_topk.cuh"

namespace lxg {

// merge_rescore_kernel is one CTA per query; its size is picked by the batch: 128 threads when there
// are many queries (a 1024-query batch is resident in one wave), up to 1024 threads for a handful of
// queries, where the gather / re-score of one query is all the parallelism there is.
constexpr int kExactListCap = 16384;  // rows an uncertified query may collect before we give up

struct CorpusView {
  const void* rows;   // original corpus rows (fp32 or fp16), used for exact scores
  long long pitch;    // elements between rows
  int dtype;          // 0 = fp32, 1 = fp16
  int n, d;
  long long row_offset;  // added to every output id (row-sharded indexes)
  float max_row_norm;    // upper bound of ||row||_2 over the corpus
  float scan_scale;      // power-of-two scale baked into the fp16 scan copy
  float rel_err;         // relative bound of |tensor-core score - exact score| / (||c|| ||q||)
};

struct MergeParams {
  const uint2* cand;
  const int* cand_count;
  const float* slice_thr;
  const uint32_t* lvl;  // [nq, lvl_slots] published levels of pass 1 (nullptr: no cross-list level)
  const float* xn;
  const float* qscale;
  const float* qnorm;
  float* out_d;        // [nq, k]
  long long* out_i;    // [nq, k]
  double* out_d64;     // optional [nq, k] exact scores (sharded merge), else nullptr
  int* flag_count;     // number of uncertified queries
  int* flag_list;      // [nq] their ids
  double* flag_theta;  // [nq] lower bound of the true k-th best exact score
  int nq, k, kp, cap, lists, lvl_slots;
  int max_items;       // capacity of the shared-memory candidate pool
  int sort_n;          // kp rounded up to a power of two (bitonic ranking of the re-scored rows)
  // A handful of queries with a large k (the engine's own request: one query, faiss_k = 1000): gathering,
  // re-scoring (~k' rows of 2-4 KB) and ranking inside one CTA takes ~0.2 ms, so the merge runs as three
  // kernels instead - this one with select_only (one CTA per query, returns after the selection),
  // rescore_rows_kernel (every SM), rank_rows_kernel (rank by counting, k'/64 CTAs per query) - handing over through:
  int select_only;
  unsigned* sel_row_g;   // [nq, sort_n]
  double* sel_score_g;   // [nq, sort_n]
  int* sel_n_g;          // [nq] rows selected, or -1 when stage 1 already finished the query
  float* sel_amin_g;     // [nq] largest tensor-core score a dropped row can have (stage 1 -> 3)
};

// Exact inner products of R corpus rows with a query held in shared memory, computed by a full
// warp with all R row gathers in flight.  Products of an fp32 by an fp16/fp32 value are exact in
// fp64; the summation order is fixed - lane L owns the element quads 4*(L + 32*j)..+3 in
// increasing j, then an xor tree - so a row's score has the same bits wherever it is computed
// (merge kernel, exact collectors, any shard).  Rows are read with 8 / 16-byte loads when the
// layout allows it.
template <int R>
__device__ __forceinline__ void warp_exact_dots(const CorpusView& cv, const long long (&row)[R],
                                                const float* __restrict__ xq, int lane, double (&acc)[R]) {
#pragma unroll
  for (int u = 0; u < R; ++u) acc[u] = 0.0;
  const int d = cv.d;
  if (cv.dtype == 1) {
    const __half* base = reinterpret_cast<const __half*>(cv.rows);
    const bool vec = (cv.pitch % 4 == 0) && (reinterpret_cast<uintptr_t>(base) % 8 == 0);
#pragma unroll 2
    for (int i = 4 * lane; i < d; i += 128) {
      float v[R][4];
      if (vec && i + 3 < d) {
#pragma unroll
        for (int u = 0; u < R; ++u) {
          const uint2 w = __ldg(reinterpret_cast<const uint2*>(base + row[u] * cv.pitch + i));
          const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&w.x));
          const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
          v[u][0] = lo.x, v[u][1] = lo.y, v[u][2] = hi.x, v[u][3] = hi.y;
        }
      } else {
#pragma unroll
        for (int u = 0; u < R; ++u)
#pragma unroll
          for (int e = 0; e < 4; ++e) v[u][e] = i + e < d ? __half2float(base[row[u] * cv.pitch + i + e]) : 0.0f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const double x = i + e < d ? static_cast<double>(xq[i + e]) : 0.0;
#pragma unroll
        for (int u = 0; u < R; ++u) acc[u] = fma(static_cast<double>(v[u][e]), x, acc[u]);
      }
    }
  } else {
    const float* base = reinterpret_cast<const float*>(cv.rows);
    const bool vec = (cv.pitch % 4 == 0) && (reinterpret_cast<uintptr_t>(base) % 16 == 0);
#pragma unroll 2
    for (int i = 4 * lane; i < d; i += 128) {
      float v[R][4];
      if (vec && i + 3 < d) {
#pragma unroll
        for (int u = 0; u < R; ++u) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(base + row[u] * cv.pitch + i));
          v[u][0] = w.x, v[u][1] = w.y, v[u][2] = w.z, v[u][3] = w.w;
        }
      } else {
#pragma unroll
        for (int u = 0; u < R; ++u)
#pragma unroll
          for (int e = 0; e < 4; ++e) v[u][e] = i + e < d ? base[row[u] * cv.pitch + i + e] : 0.0f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const double x = i + e < d ? static_cast<double>(xq[i + e]) : 0.0;
#pragma unroll
        for (int u = 0; u < R; ++u) acc[u] = fma(static_cast<double>(v[u][e]), x, acc[u]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < R; ++u)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
}
__device__ __forceinline__ double warp_exact_dot(const CorpusView& cv, long long row,
                                                 const float* __restrict__ xq, int lane) {
  const long long rows[1] = {row};
  double acc[1];
  warp_exact_dots<1>(cv, rows, xq, lane, acc);
  return acc[0];
}

__device__ __forceinline__ bool better(double sa, unsigned ia, double sb, unsigned ib) {
  return sa > sb || (sa == sb && ia < ib);
}
// The rankings compare scores as order-preserving 64-bit integer keys: fp64 compares issue at a small
// fraction of the integer rate on this part, and a ranking is nothing but compares.  -0.0 is folded
// into +0.0 first so that equal scores have equal keys.
__device__ __forceinline__ unsigned long long score_key(double x) {
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(x + 0.0));
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_score(unsigned long long k) {
  return __longlong_as_double(static_cast<long long>((k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k));
}
__device__ __forceinline__ bool better_key(unsigned long long ka, unsigned ia, unsigned long long kb, unsigned ib) {
  return ka > kb || (ka == kb && ia < ib);
}

// Walks every entry of every candidate list of query q with a warp per PAIR of lists and two 32-entry
// chunks of each in flight (four independent L2 loads per lane): the lists are short (tens of entries),
// so a walk is a chain of dependent round trips and their number is what it costs.  f(entry, valid) is
// called warp-uniformly (it may use ballots).
template <int kThreads, class F>
__device__ __forceinline__ void for_each_list_entry(const MergeParams& p, int q, const int* s_len, int warp, int lane, F&& f) {
  constexpr int kWarps = kThreads / 32;
  for (int s0 = warp; s0 < p.lists; s0 += 2 * kWarps) {
    const int s1 = s0 + kWarps;
    const bool two = s1 < p.lists;
    const uint2* l0 = p.cand + (static_cast<size_t>(s0) * p.nq + q) * p.cap;
    const uint2* l1 = p.cand + (static_cast<size_t>(two ? s1 : s0) * p.nq + q) * p.cap;
    const int c0 = s_len[s0], c1 = two ? s_len[s1] : 0;
    for (int i0 = 0; i0 < max(c0, c1); i0 += 64) {
      uint2 e[4];
      bool v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + (u & 1) * 32 + lane;
        v[u] = i < (u < 2 ? c0 : c1);
        e[u] = v[u] ? __ldcg((u < 2 ? l0 : l1) + i) : make_uint2(0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) f(e[u], v[u]);
    }
  }
}

// Histogram increment for a whole warp (call warp-uniformly): lanes that hit the same bin are
// combined first.  Score keys of one query share their leading bits, so the first radix pass puts
// thousands of entries into a handful of bins - one shared-memory atomic per entry would serialise.
__device__ __forceinline__ void warp_hist_add(int* hist, uint32_t bin, bool pred) {
  const uint32_t peers = __match_any_sync(0xffffffffu, pred ? bin : 0xFFFFFFFFu);
  if (pred && static_cast<int>(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
}

// Warp helper of the radix selections below: the highest bin b of hist[0, nbins) with
// sum(hist[b ..]) >= need.  out[0] = b (-1: the whole histogram holds fewer than `need`), out[1] =
// entries in the bins above b (or the total when b = -1), out[2] = hist[b].  nbins is a multiple of 32.
__device__ __forceinline__ void radix_boundary_bin(const int* hist, int nbins, int need, int lane, int* out) {
  const int per = nbins / 32;
  int mine = 0;
  // lane owns bins [lane * per, (lane + 1) * per); read rotated by the lane so that the 32 lanes hit 32
  // different banks (per is a multiple of 32: unrotated, every lane would read the same bank)
  for (int i = 0; i < per; ++i) mine += hist[lane * per + ((i + lane) & (per - 1))];
  int suffix = mine;  // inclusive suffix sum over the lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_down_sync(0xffffffffu, suffix, o);
    if (lane + o < 32) suffix += v;
  }
  const int all = __shfl_sync(0xffffffffu, suffix, 0);
  const uint32_t ok = __ballot_sync(0xffffffffu, suffix >= need);
  if (ok == 0u) {
    if (lane == 0) {
      out[0] = -1;
      out[1] = all;
      out[2] = 0;
    }
    return;
  }
  if (lane == 31 - __clz(ok)) {
    int run = suffix - mine;  // entries in the lanes above this one
    int b = per - 1;
    for (; b > 0; --b) {
      if (run + hist[lane * per + b] >= need) break;
      run += hist[lane * per + b];
    }
    out[0] = lane * per + b;
    out[1] = run;
    out[2] = hist[lane * per + b];
  }
}

// One CTA (kMergeThreads threads) per query.  Dynamic shared memory: max_items (score key, row)
// pairs, then kp (double,uint) pairs, then d floats.
template <int kMergeThreads>
__global__ void __launch_bounds__(kMergeThreads, kMergeThreads <= 128 ? 8 : (kMergeThreads <= 256 ? 4 : 1))
merge_rescore_kernel(const MergeParams p, const CorpusView cv) {
  extern __shared__ __align__(16) uint8_t msm[];
  const int q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint2* items = reinterpret_cast<uint2*>(msm);  // .x = ordered score key, .y = row
  double* sel_score = reinterpret_cast<double*>(items + p.max_items);  // [sort_n] (kp rounded up to a power of two)
  unsigned* sel_row = reinterpret_cast<unsigned*>(sel_score + p.sort_n);
  float* xq = reinterpret_cast<float*>(sel_row + p.sort_n);
  __shared__ int s_cnt[3], s_sel, s_m;
  __shared__ int s_len[kGroups * 148];  // list lengths (kGroups lists per slice, at most 148 slices)
  __shared__ uint32_t s_kmin, s_kmax, s_tkey, s_lvl;
  __shared__ double s_kth;

  if (tid == 0) {
    s_sel = 0;
    s_m = 0;
    s_cnt[0] = s_cnt[1] = s_cnt[2] = 0;
    s_kth = 0.0;
    s_kmin = 0xFFFFFFFFu;
    s_kmax = 0u;
    s_tkey = 0u;
    s_lvl = kLvlSkip;
  }
  for (int i = tid; i < cv.d; i += kMergeThreads) xq[i] = p.xn[static_cast<size_t>(q) * cv.d + i];
  __syncthreads();
  int nsel = 0;
  float a_min = -CUDART_INF_F;
  if (p.select_only && tid == 0) p.sel_n_g[q] = -1;  // "finished here" until the selection below completes
  if (p.qnorm[q] == 0.0f) {
    // all-zero query (normalize_L2 leaves it untouched): every inner product is exactly 0, so the
    // answer is the k lowest row ids.  Pass 1 keeps no candidates for it (every score ties with
    // its threshold), hence the closed form here.
    for (int r = tid; r < p.k; r += kMergeThreads) {
      const bool have = r < cv.n;
      p.out_d[static_cast<size_t>(q) * p.k + r] = have ? 0.0f : -FLT_MAX;
      p.out_i[static_cast<size_t>(q) * p.k + r] = have ? static_cast<long long>(r) + cv.row_offset : -1;
      if (p.out_d64) p.out_d64[static_cast<size_t>(q) * p.k + r] = have ? 0.0 : -static_cast<double>(FLT_MAX);
    }
    return;
  }
  // largest final list threshold (no dropped row scored above it) and the final cross-list level
  // (at least kp rows score >= it: entries below it cannot be among the best kp)
  {
    uint32_t tk = 0u, lo = kLvlSkip;
    for (int s = tid; s < p.lists; s += kMergeThreads) {
      s_len[s] = p.cand_count[static_cast<size_t>(s) * p.nq + q];
      tk = max(tk, float_to_key(__float_as_uint(p.slice_thr[static_cast<size_t>(s) * p.nq + q])));
    }
    if (p.lvl != nullptr)
      for (int s = tid; s < p.lvl_slots; s += kMergeThreads) lo = min(lo, __ldcg(p.lvl + static_cast<size_t>(q) * p.lvl_slots + s));
    tk = __reduce_max_sync(0xffffffffu, tk);
    lo = __reduce_min_sync(0xffffffffu, lo);
    if (lane == 0) {
      atomicMax(&s_tkey, tk);
      atomicMin(&s_lvl, lo);
    }
  }
  __syncthreads();
  uint32_t tau_key = (p.lvl != nullptr && s_lvl != kLvlNone && s_lvl != kLvlSkip) ? s_lvl : 0u;
  // gather the entries at or above the level into shared memory (a warp per list, four 32-entry
  // chunks in flight).  If they do not fit the pool (row orders that defeat the running thresholds:
  // a sorted or strongly clustered corpus), the level is first tightened by bisection over the
  // lists in global memory - slow, but exact and only taken then.
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
  int m = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    kmin = 0xFFFFFFFFu;
    kmax = 0u;
    for_each_list_entry<kMergeThreads>(p, q, s_len, warp, lane, [&](const uint2& e, bool valid) {
      const uint32_t key = float_to_key(e.x);
      const bool keep = valid && key >= tau_key;
      const uint32_t bal = __ballot_sync(0xffffffffu, keep);
      if (bal == 0u) return;
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_m, __popc(bal));
      base = __shfl_sync(0xffffffffu, base, 0);
      const int pos = base + __popc(bal & ((1u << lane) - 1u));
      if (keep && pos < p.max_items) {
        items[pos] = make_uint2(key, e.y);
        kmin = min(kmin, key);
        kmax = max(kmax, key);
      }
    });
    __syncthreads();
    m = s_m;
    if (m <= p.max_items || attempt == 1) break;
    // ---- tighten: a prefix with count(key >= prefix) >= kp whose count fits the pool, by radix
    // histograms over the lists in global memory: 12 + 12 + 8 key bits per pass, stopping at the first
    // pass whose boundary bin fits (normally the first or second).  The histogram aliases the pool,
    // which is refilled afterwards.
    uint32_t prefix = 0u;
    {
      int* hist = reinterpret_cast<int*>(items);
      int above = 0;  // keys above the bin range still being resolved
      int shift = 20, nbits = 12;
      for (int pass = 0; pass < 3; ++pass) {
        const int nbins = 1 << nbits;
        for (int i = tid; i < nbins; i += kMergeThreads) hist[i] = 0;
        __syncthreads();
        const int hs = shift + nbits;  // bits above the pass's digit must equal the prefix
        for_each_list_entry<kMergeThreads>(p, q, s_len, warp, lane, [&](const uint2& e, bool valid) {
          const uint32_t key = float_to_key(e.x);
          warp_hist_add(hist, (key >> shift) & (nbins - 1), valid && (hs >= 32 || (key >> hs) == (prefix >> hs)));
        });
        __syncthreads();
        if (warp == 0) radix_boundary_bin(hist, nbins, p.kp - above, lane, s_cnt);
        __syncthreads();
        const int bin = s_cnt[0], over = s_cnt[1], inbin = s_cnt[2];
        __syncthreads();
        if (bin < 0) {
          prefix = 0u;
          break;
        }
        prefix |= static_cast<uint32_t>(bin) << shift;
        if (above + over + inbin <= p.max_items) break;  // count(key >= prefix) fits the pool
        above += over;
        nbits = pass == 0 ? 12 : 8;
        shift -= nbits;
      }
    }
    __syncthreads();
    if (tid == 0) {
      s_cnt[0] = s_cnt[1] = s_cnt[2] = 0;
      s_m = 0;
    }
    tau_key = max(tau_key, prefix);
    __syncthreads();
  }
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  if (lane == 0) {
    atomicMin(&s_kmin, kmin);
    atomicMax(&s_kmax, kmax);
  }
  __syncthreads();
  const float tau = tau_key ? __uint_as_float(key_to_float_bits(tau_key)) : -CUDART_INF_F;
  if (m > p.max_items) {
    // still more survivors than the pool holds: more than max_items - kp rows tie (in tensor-core
    // score) around the kp-th best.  Hand the query to the exact path; at least kp >= k rows have
    // a tensor-core score >= tau, which bounds the true k-th best.
    if (tid == 0) {
      const double unscale = 1.0 / (static_cast<double>(p.qscale[q]) * cv.scan_scale);
      const double eps = static_cast<double>(cv.max_row_norm) * p.qnorm[q] * cv.rel_err;
      const int slot = atomicAdd(p.flag_count, 1);
      p.flag_list[slot] = q;
      p.flag_theta[slot] = static_cast<double>(tau) * unscale - eps;
    }
    return;
  }

  // cut = kp-th largest score key (bisection from the highest bit in which the keys differ);
  // entries above the cut are selected, ties with the cut by ascending row
  uint32_t cut = 0;
  int need_eq = 0x7fffffff;
  // approx score no dropped row can exceed (scaled units)
  a_min = fmaxf(__uint_as_float(key_to_float_bits(s_tkey)), tau);
  if (m > p.kp) {
    kmin = s_kmin;
    kmax = s_kmax;
    cut = kmax;
    const int hist_ints = (p.sort_n * 12) / 4;  // sel_score / sel_row are not in use yet: room for the histogram
    if (kmin != kmax && hist_ints >= 2048) {
      // exact k'-th largest key by radix selection over the pool (12 + 12 + 8 or 11 + 11 + 10 key bits)
      int* hist = reinterpret_cast<int*>(sel_score);
      const int b1 = hist_ints >= 4096 ? 12 : 11;
      int shift = 32 - b1, nbits = b1, above = 0;
      uint32_t pre = 0u;
      for (int pass = 0; pass < 3; ++pass) {
        const int nbins = 1 << nbits;
        for (int i = tid; i < nbins; i += kMergeThreads) hist[i] = 0;
        __syncthreads();
        const int hs = shift + nbits;
        for (int i = tid; (i & ~31) < m; i += kMergeThreads) {  // warp-uniform trip count
          const uint32_t key = i < m ? items[i].x : 0u;
          warp_hist_add(hist, (key >> shift) & (nbins - 1), i < m && (hs >= 32 || (key >> hs) == (pre >> hs)));
        }
        __syncthreads();
        if (warp == 0) radix_boundary_bin(hist, nbins, p.kp - above, lane, s_cnt);
        __syncthreads();
        pre |= static_cast<uint32_t>(s_cnt[0]) << shift;  // m > kp: a boundary bin always exists
        above += s_cnt[1];
        __syncthreads();
        nbits = pass == 0 ? b1 : 32 - 2 * b1;
        shift -= nbits;
      }
      cut = pre;
    } else if (kmin != kmax) {
      const int hb = 31 - __clz(kmin ^ kmax);
      cut = (hb == 31) ? 0u : (kmax & ~((2u << hb) - 1u));
      int cur = 0;  // three rotating counters: one barrier per step instead of three
      for (int bit = hb; bit >= 0; --bit) {
        const uint32_t cnd = cut | (1u << bit);
        int mine = 0;
        for (int i = tid; i < m; i += kMergeThreads) mine += (items[i].x >= cnd) ? 1 : 0;
        mine = __reduce_add_sync(0xffffffffu, mine);
        if (lane == 0 && mine) atomicAdd(&s_cnt[cur], mine);
        __syncthreads();
        if (s_cnt[cur] >= p.kp) cut = cnd;
        // counter cur+2 was last read before this barrier and is next written after the following one
        if (tid == 0) s_cnt[(cur + 2) % 3] = 0;
        cur = (cur + 1) % 3;
      }
    }
    __syncthreads();
    if (tid == 0) s_cnt[0] = s_cnt[1] = 0;
    __syncthreads();
    int gt = 0, eq = 0;
    for (int i = tid; i < m; i += kMergeThreads) {
      gt += (items[i].x > cut) ? 1 : 0;
      eq += (items[i].x == cut) ? 1 : 0;
    }
    gt = __reduce_add_sync(0xffffffffu, gt);
    eq = __reduce_add_sync(0xffffffffu, eq);
    if (lane == 0 && gt) atomicAdd(&s_cnt[0], gt);
    if (lane == 0 && eq) atomicAdd(&s_cnt[1], eq);
    __syncthreads();
    need_eq = p.kp - s_cnt[0];
    if (s_cnt[1] == need_eq) need_eq = 0x7fffffff;  // every tie is selected: no ranking needed
    a_min = fmaxf(a_min, __uint_as_float(key_to_float_bits(cut)));
  }
  // gather the selected rows
  for (int i = tid; i < m; i += kMergeThreads) {
    const uint2 e = items[i];
    bool take = e.x > cut || m <= p.kp;
    if (!take && e.x == cut) {
      take = need_eq == 0x7fffffff;
    }
    if (!take && e.x == cut) {
      // exact ties straddle the cut: rank them by row id (rare)
      int rank = 0;
      for (int u = 0; u < m; ++u) rank += (items[u].x == cut && items[u].y < e.y) ? 1 : 0;
      take = rank < need_eq;
    }
    if (take) sel_row[atomicAdd(&s_sel, 1)] = e.y;
  }
  __syncthreads();
  nsel = s_sel;  // == min(m, kp)
  if (p.select_only) {  // three-kernel merge: the re-score and the ranking run on every SM
    for (int j = tid; j < nsel; j += kMergeThreads) p.sel_row_g[static_cast<size_t>(q) * p.sort_n + j] = sel_row[j];
    if (tid == 0) {
      p.sel_n_g[q] = nsel;
      p.sel_amin_g[q] = a_min;
    }
    return;
  }
  // exact re-score: four rows in flight per warp (the rows are random gathers from HBM)
  for (int j0 = warp * 4; j0 < nsel; j0 += (kMergeThreads / 32) * 4) {
    long long rows[4];
    double acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) rows[u] = static_cast<long long>(sel_row[min(j0 + u, nsel - 1)]);
    warp_exact_dots<4>(cv, rows, xq, lane, acc);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (lane == 0 && j0 + u < nsel) sel_score[j0 + u] = acc[u];
  }
  __syncthreads();
  // rank the re-scored rows by (score desc, row asc): bitonic sort in shared memory (padding
  // entries sort last), then the first k are the answer
  const int k = p.k;
  unsigned long long* sel_key = reinterpret_cast<unsigned long long*>(sel_score);  // in place
  for (int j = tid; j < nsel; j += kMergeThreads) sel_key[j] = score_key(sel_score[j]);
  for (int j = nsel + tid; j < p.sort_n; j += kMergeThreads) {
    sel_key[j] = 0ull;  // below every real score
    sel_row[j] = 0xFFFFFFFFu;
  }
  __syncthreads();
  for (int size = 2; size <= p.sort_n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < (p.sort_n >> 1); i += kMergeThreads) {
        const int lo = 2 * i - (i & (stride - 1));  // element whose `stride` bit is clear
        const int hi = lo + stride;
        const bool descending = (lo & size) == 0;
        const unsigned long long sl = sel_key[lo], sh = sel_key[hi];
        const unsigned rl = sel_row[lo], rh = sel_row[hi];
        if (better_key(sh, rh, sl, rl) == descending) {
          sel_key[lo] = sh;
          sel_key[hi] = sl;
          sel_row[lo] = rh;
          sel_row[hi] = rl;
        }
      }
      __syncthreads();
    }
  }
  for (int r = tid; r < min(k, nsel); r += kMergeThreads) {
    const double sc = key_score(sel_key[r]);
    p.out_d[static_cast<size_t>(q) * k + r] = static_cast<float>(sc);
    p.out_i[static_cast<size_t>(q) * k + r] = static_cast<long long>(sel_row[r]) + cv.row_offset;
    if (p.out_d64) p.out_d64[static_cast<size_t>(q) * k + r] = sc;
  }
  if (tid == 0 && nsel >= k) s_kth = key_score(sel_key[k - 1]);
  for (int r = nsel + tid; r < k; r += kMergeThreads) {  // fewer than k rows: FAISS pads -FLT_MAX / -1
    p.out_d[static_cast<size_t>(q) * k + r] = -FLT_MAX;
    p.out_i[static_cast<size_t>(q) * k + r] = -1;
    if (p.out_d64) p.out_d64[static_cast<size_t>(q) * k + r] = -static_cast<double>(FLT_MAX);
  }
  __syncthreads();
  if (tid == 0) {
    bool certified;
    double theta = 0.0;
    if (a_min == -CUDART_INF_F) {
      certified = true;  // nothing was ever dropped: every row of the corpus was re-scored
    } else {
      const double unscale = 1.0 / (static_cast<double>(p.qscale[q]) * cv.scan_scale);
      const double eps = static_cast<double>(cv.max_row_norm) * p.qnorm[q] * cv.rel_err;
      const double bound = static_cast<double>(a_min) * unscale + eps;  // >= any dropped row's exact score
      if (nsel >= k) {
        certified = s_kth > bound;
        theta = s_kth;
      } else {
        certified = false;
        theta = static_cast<double>(a_min) * unscale - eps;
      }
    }
    if (!certified) {
      const int slot = atomicAdd(p.flag_count, 1);
      p.flag_list[slot] = q;
      p.flag_theta[slot] = theta;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Three-kernel merge for a handful of queries with a large k: merge_rescore_kernel with
// MergeParams.select_only (returns after writing the selected rows), rescore_rows_kernel, rank_rows_kernel.
//
// rank_rows_kernel: ranks the re-scored rows of a query by (score desc, row asc) by COUNTING - eight
// threads per row each compare it with an eighth of the others (all in shared memory, broadcast reads),
// 32 rows per CTA, k'/32 CTAs per query - writes the first k in order, the FAISS padding when there are
// fewer than k, and certifies the answer exactly as merge_rescore_kernel does (the thread that owns the
// k-th best does it).
static __global__ void __launch_bounds__(256) rank_rows_kernel(const MergeParams p, const CorpusView cv) {
  extern __shared__ __align__(16) uint8_t ksm[];
  const int q = blockIdx.y;
  const int nsel = p.sel_n_g[q];
  if (nsel < 0) return;  // zero query / handed to the exact path by the selection kernel
  if (blockIdx.x != 0 && blockIdx.x * 32 >= nsel) return;  // no row of this CTA (CTA 0 also pads / certifies)
  unsigned long long* sk = reinterpret_cast<unsigned long long*>(ksm);  // order-preserving score keys
  unsigned* rw = reinterpret_cast<unsigned*>(sk + p.sort_n);
  const int tid = threadIdx.x, k = p.k;
  for (int j = tid; j < nsel; j += 256) {
    sk[j] = score_key(p.sel_score_g[static_cast<size_t>(q) * p.sort_n + j]);
    rw[j] = p.sel_row_g[static_cast<size_t>(q) * p.sort_n + j];
  }
  __syncthreads();
  const int i = blockIdx.x * 32 + (tid >> 3), part = tid & 7;
  const bool have = i < nsel;
  const unsigned long long ki = have ? sk[i] : 0ull;
  const double si = key_score(ki);
  const unsigned ri = have ? rw[i] : 0u;
  const int chunk = (nsel + 7) >> 3;
  int rank = 0;
  if (have) {
    const int j1 = min(nsel, (part + 1) * chunk);
#pragma unroll 4
    for (int j = part * chunk; j < j1; ++j) rank += better_key(sk[j], rw[j], ki, ri) ? 1 : 0;
  }
  rank += __shfl_xor_sync(0xffffffffu, rank, 1);
  rank += __shfl_xor_sync(0xffffffffu, rank, 2);
  rank += __shfl_xor_sync(0xffffffffu, rank, 4);
  if (have && part == 0 && rank < k) {
    p.out_d[static_cast<size_t>(q) * k + rank] = static_cast<float>(si);
    p.out_i[static_cast<size_t>(q) * k + rank] = static_cast<long long>(ri) + cv.row_offset;
    if (p.out_d64) p.out_d64[static_cast<size_t>(q) * k + rank] = si;
  }
  if (blockIdx.x == 0)
    for (int r = nsel + tid; r < k; r += 256) {  // fewer than k rows: FAISS pads -FLT_MAX / -1
      p.out_d[static_cast<size_t>(q) * k + r] = -FLT_MAX;
      p.out_i[static_cast<size_t>(q) * k + r] = -1;
      if (p.out_d64) p.out_d64[static_cast<size_t>(q) * k + r] = -static_cast<double>(FLT_MAX);
    }
  // certificate: by the owner of the k-th best, or by the first thread when there are fewer than k rows
  const bool owner = nsel >= k ? (have && part == 0 && rank == k - 1) : (blockIdx.x == 0 && tid == 0);
  if (owner) {
    const float a_min = p.sel_amin_g[q];
    bool certified;
    double theta = 0.0;
    if (a_min == -CUDART_INF_F) {
      certified = true;  // nothing was ever dropped: every row of the corpus was re-scored
    } else {
      const double unscale = 1.0 / (static_cast<double>(p.qscale[q]) * cv.scan_scale);
      const double eps = static_cast<double>(cv.max_row_norm) * p.qnorm[q] * cv.rel_err;
      const double bound = static_cast<double>(a_min) * unscale + eps;  // >= any dropped row's exact score
      if (nsel >= k) {
        certified = si > bound;
        theta = si;
      } else {
        certified = false;
        theta = static_cast<double>(a_min) * unscale - eps;
      }
    }
    if (!certified) {
      const int slot = atomicAdd(p.flag_count, 1);
      p.flag_list[slot] = q;
      p.flag_theta[slot] = theta;
    }
  }
}

// rescore_rows_kernel: exact scores of the selected rows, 32 rows per CTA (8 warps x 4
// rows in flight), grid = (sort_n / 32, nq) - the gathers of one query are spread over every SM.  Same
// warp_exact_dots as the one-kernel path: identical bits.
static __global__ void __launch_bounds__(256) rescore_rows_kernel(const MergeParams p, const CorpusView cv) {
  extern __shared__ __align__(16) uint8_t rsm[];
  float* xq = reinterpret_cast<float*>(rsm);
  const int q = blockIdx.y;
  const int nsel = p.sel_n_g[q];
  const int j0 = blockIdx.x * 32 + (threadIdx.x >> 5) * 4;
  if (blockIdx.x * 32 >= nsel) return;
  for (int i = threadIdx.x; i < cv.d; i += 256) xq[i] = p.xn[static_cast<size_t>(q) * cv.d + i];
  __syncthreads();
  if (j0 >= nsel) return;
  const unsigned* rows_g = p.sel_row_g + static_cast<size_t>(q) * p.sort_n;
  long long rows[4];
  double acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) rows[u] = static_cast<long long>(rows_g[min(j0 + u, nsel - 1)]);
  warp_exact_dots<4>(cv, rows, xq, threadIdx.x & 31, acc);
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if ((threadIdx.x & 31) == 0 && j0 + u < nsel) p.sel_score_g[static_cast<size_t>(q) * p.sort_n + j0 + u] = acc[u];
}

// ---------------------------------------------------------------------------------------
// Exact path for uncertified queries: collect every row whose exact score is >= theta (theta
// is a proven lower bound of the true k-th best), then order that short list.
struct ExactParams {
  const float* xn;
  const int* flag_count;
  const int* flag_list;
  const double* flag_theta;
  double* list_score;  // [nflag_max, kExactListCap]
  unsigned* list_row;  // [nflag_max, kExactListCap]
  int* list_count;     // [nflag_max]
  float* out_d;
  long long* out_i;
  double* out_d64;
  int* overflow;       // set to 1 if some list overflowed (result would be wrong -> error)
  int nq, k;
  int nflag_max;       // lists allocated; queries flagged beyond this are reported as an error
};

__global__ void __launch_bounds__(256) exact_collect_kernel(const ExactParams p, const CorpusView cv) {
  const int nflag = min(*p.flag_count, p.nflag_max);
  if (nflag == 0) return;
  extern __shared__ __align__(16) float xq_s[];
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const long long gwarp = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * warps_per_block;
  for (int f = 0; f < nflag; ++f) {
    const int q = p.flag_list[f];
    const double theta = p.flag_theta[f];
    __syncthreads();
    for (int i = threadIdx.x; i < cv.d; i += blockDim.x) xq_s[i] = p.xn[static_cast<size_t>(q) * cv.d + i];
    __syncthreads();
    for (long long row = gwarp; row < cv.n; row += nwarps) {
      const double s = warp_exact_dot(cv, row, xq_s, lane);
      if (lane == 0 && s >= theta) {
        const int slot = atomicAdd(&p.list_count[f], 1);
        if (slot < kExactListCap) {
          p.list_score[static_cast<size_t>(f) * kExactListCap + slot] = s;
          p.list_row[static_cast<size_t>(f) * kExactListCap + slot] = static_cast<unsigned>(row);
        } else {
          *p.overflow = 1;
        }
      }
    }
  }
}

// One CTA per uncertified query: k rounds of block-wide arg-best over the collected list.
__global__ void __launch_bounds__(256) exact_finalize_kernel(const ExactParams p, const CorpusView cv) {
  const int nflag = min(*p.flag_count, p.nflag_max);
  const int f = blockIdx.x;
  if (f >= nflag) return;
  const int q = p.flag_list[f];
  const int cnt = min(p.list_count[f], kExactListCap);
  double* sc = p.list_score + static_cast<size_t>(f) * kExactListCap;
  unsigned* rw = p.list_row + static_cast<size_t>(f) * kExactListCap;
  __shared__ double w_s[8];
  __shared__ unsigned w_r[8];
  __shared__ int w_i[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int r = 0; r < p.k; ++r) {
    double bs = -CUDART_INF;
    unsigned br = 0xFFFFFFFFu;
    int bi = -1;
    for (int i = tid; i < cnt; i += 256) {
      const unsigned row = rw[i];
      if (row == 0xFFFFFFFFu) continue;  // already emitted
      if (bi < 0 || better(sc[i], row, bs, br)) {
        bs = sc[i];
        br = row;
        bi = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double os = __shfl_xor_sync(0xffffffffu, bs, o);
      const unsigned orr = __shfl_xor_sync(0xffffffffu, br, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (bi < 0 || better(os, orr, bs, br))) {
        bs = os;
        br = orr;
        bi = oi;
      }
    }
    if (lane == 0) {
      w_s[warp] = bs;
      w_r[warp] = br;
      w_i[warp] = bi;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w)
        if (w_i[w] >= 0 && (w_i[0] < 0 || better(w_s[w], w_r[w], w_s[0], w_r[0]))) {
          w_s[0] = w_s[w];
          w_r[0] = w_r[w];
          w_i[0] = w_i[w];
        }
      const size_t o = static_cast<size_t>(q) * p.k + r;
      if (w_i[0] >= 0) {
        p.out_d[o] = static_cast<float>(w_s[0]);
        p.out_i[o] = static_cast<long long>(w_r[0]) + cv.row_offset;
        if (p.out_d64) p.out_d64[o] = w_s[0];
        rw[w_i[0]] = 0xFFFFFFFFu;
      } else {
        p.out_d[o] = -FLT_MAX;
        p.out_i[o] = -1;
        if (p.out_d64) p.out_d64[o] = -static_cast<double>(FLT_MAX);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// Row-sharded indexes: merge `shards` per-shard results (exact fp64 scores + global int64 ids, as
// all-gathered over NCCL) into the global top-k.  Shard s's [nq, k] planes start at
// dg + s * shard_stride and ig + s * shard_stride (elements), so both the legacy [shards, nq, k]
// pair of arrays and the packed all-gather buffer [shards][2][nq][k] are read in place.
// One CTA per query, one thread per candidate.  Every per-shard list is ordered best first
// ((score desc, id asc), -1 padding at the end - what lxg_search_ex writes), so a candidate's
// global rank is its position in its own list plus, for every other shard, the number of entries
// that beat it: one binary search per (candidate, other shard) over shared memory (kSmem) or,
// for pools beyond the shared-memory budget, over the gathered buffer itself (L2).
template <bool kSmem>
__global__ void __launch_bounds__(256)
merge_shards_kernel(const double* __restrict__ dg, const long long* __restrict__ ig, long long shard_stride,
                    int nq, int k, int shards, float* __restrict__ out_d, long long* __restrict__ out_i) {
  extern __shared__ __align__(16) uint8_t msh[];
  __shared__ int s_valid;
  const int q = blockIdx.x;
  const int tid = threadIdx.x;
  const int total = shards * k;
  double* sd = reinterpret_cast<double*>(msh);
  long long* si = reinterpret_cast<long long*>(sd + (kSmem ? total : 0));
  const size_t qoff = static_cast<size_t>(q) * k;
  if (tid == 0) s_valid = 0;
  if (kSmem) {
    for (int j = tid; j < total; j += blockDim.x) {
      const size_t src = static_cast<size_t>(j / k) * shard_stride + qoff + (j % k);
      sd[j] = dg[src];
      si[j] = ig[src];
    }
  }
  __syncthreads();
  auto score_at = [&](int s, int r) { return kSmem ? sd[s * k + r] : dg[static_cast<size_t>(s) * shard_stride + qoff + r]; };
  auto id_at = [&](int s, int r) { return kSmem ? si[s * k + r] : ig[static_cast<size_t>(s) * shard_stride + qoff + r]; };
  int valid = 0;
  for (int j = tid; j < total; j += blockDim.x) {
    const int s = j / k, r = j % k;
    const long long id = id_at(s, r);
    if (id < 0) continue;
    ++valid;
    const double sc = score_at(s, r);
    int rank = r;
    for (int u = 0; u < shards; ++u) {
      if (u == s) continue;
      int lo = 0, hi = k;  // first entry of list u that does not beat (sc, id)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const long long uid = id_at(u, mid);
        const double usc = score_at(u, mid);
        if (uid >= 0 && (usc > sc || (usc == sc && uid < id))) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) {
      out_d[qoff + rank] = static_cast<float>(sc);
      out_i[qoff + rank] = id;
    }
  }
  valid = __reduce_add_sync(0xffffffffu, valid);
  if ((tid & 31) == 0 && valid) atomicAdd(&s_valid, valid);
  __syncthreads();
  for (int r = s_valid + tid; r < k; r += blockDim.x) {  // fewer than k rows in the whole corpus
    out_d[qoff + r] = -FLT_MAX;
    out_i[qoff + r] = -1;
  }
}

// ---------------------------------------------------------------------------------------
// Index preparation (what faiss.Index.add does for a flat index: store the rows; here also the
// fp16 scan copy and the statistics the certificate needs).
// Per-block partial max of row L2 norm^2 (fp64) and of |element|.
__global__ void __launch_bounds__(256)
corpus_stats_kernel(const void* rows, long long pitch, int dtype, long long n, int d,
                    double* __restrict__ blk_norm2, float* __restrict__ blk_amax) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gwarp = static_cast<long long>(blockIdx.x) * 8 + warp;
  const long long nwarps = static_cast<long long>(gridDim.x) * 8;
  double best = 0.0;
  float amax = 0.0f;
  for (long long row = gwarp; row < n; row += nwarps) {
    double acc = 0.0;
    if (dtype == 1) {
      const __half* r = reinterpret_cast<const __half*>(rows) + row * pitch;
      for (int i = lane; i < d; i += 32) {
        const float v = __half2float(r[i]);
        acc = fma(static_cast<double>(v), static_cast<double>(v), acc);
        amax = fmaxf(amax, fabsf(v));
      }
    } else {
      const float* r = reinterpret_cast<const float*>(rows) + row * pitch;
      for (int i = lane; i < d; i += 32) {
        const float v = r[i];
        acc = fma(static_cast<double>(v), static_cast<double>(v), acc);
        amax = fmaxf(amax, fabsf(v));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    best = fmax(best, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  __shared__ double sb[8];
  __shared__ float sa[8];
  if (lane == 0) {
    sb[warp] = best;
    sa[warp] = amax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      sb[0] = fmax(sb[0], sb[w]);
      sa[0] = fmaxf(sa[0], sa[w]);
    }
    blk_norm2[blockIdx.x] = sb[0];
    blk_amax[blockIdx.x] = sa[0];
  }
}

// fp32/fp16 rows -> fp16 scan copy with a 16-byte-multiple pitch, scaled by a power of two.
__global__ void __launch_bounds__(256)
make_scan_copy_kernel(const void* rows, long long pitch, int dtype, long long n, int d,
                      __half* __restrict__ out, int out_pitch, float scale) {
  const long long total = n * out_pitch;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / out_pitch;
    const int col = static_cast<int>(i - row * out_pitch);
    float v = 0.0f;
    if (col < d) {
      v = dtype == 1 ? __half2float(reinterpret_cast<const __half*>(rows)[row * pitch + col])
                     : reinterpret_cast<const float*>(rows)[row * pitch + col];
    }
    out[i] = __float2half_rn(v * scale);
  }
}

}  // namespace lxg
