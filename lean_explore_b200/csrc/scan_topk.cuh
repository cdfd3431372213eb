// Pass 1 of the flat inner-product search (replaces faiss.normalize_L2 + Index.search,
// reference call sites: src/lean_explore/search/engine.py:242 and :250).
//
// One persistent CTA per (corpus slice, block of 128 queries):
//   * epilogue warps 0-3: thread t owns query t of the block.  Prologue = the fused
//     L2-normalise (FAISS fvec_renorm_L2 semantics, engine.py:242): read the fp32 query row,
//     normalise, scale by a power of two, convert to fp16 and park it in TENSOR MEMORY as the
//     A operand (lane t, d/2 columns).  The queries never touch shared memory.
//   * warp 4: TMA producer. Streams the slice of the fp16 corpus ("scan copy") through a
//     ring of 128B-swizzled shared-memory stages (N_T rows x 64 dims each) - all ~192 KB of
//     shared memory is corpus pipeline.
//   * warp 5: one elected thread issues tcgen05.mma (M=128 queries, N=N_T corpus rows, K=16),
//     A from TMEM, B from the swizzled stage, fp32 accumulators in TMEM, double buffered.
//   * epilogue: tcgen05.ld 32 scores per thread at a time (lane = query, column = corpus
//     row), compare against the thread's private threshold; survivors are appended to the
//     query's candidate list (global memory, L2 resident).  When a list fills up the warp
//     compacts it cooperatively to the best kp entries (exact k-th-largest by bit bisection)
//     and raises the threshold.
// The output is, per (slice, query), the exact top-kp of the slice *by fp16-input score*.
// Pass 2 (rescore.cuh) merges slices, re-scores the survivors exactly and certifies that the
// answer equals the exact top-k.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include "ptx.cuh"

namespace lxg {

constexpr int kEpiThreads = 128;   // warps 0..3
constexpr int kScanThreads = 192;  // + warp 4 (TMA) + warp 5 (MMA)
constexpr int kKC = 64;            // fp16 elements per 128-byte swizzled row
constexpr int kStageRing = 196608; // bytes of shared memory used as corpus pipeline
constexpr int kStageBytes = 32768; // one pipeline stage: N_T rows x (32768 / (128 N_T)) k-chunks
constexpr int kQueryBlock = 128;   // queries per CTA == UMMA M

struct ScanParams {
  const float* x;     // [nq, d] fp32 queries as given by the caller
  float* xn;          // [nq, d] fp32 queries after normalize_L2 (written by slice 0)
  float* qscale;      // [nq] power-of-two scale applied before the fp16 conversion
  float* qnorm;       // [nq] ||xn||_2
  uint2* cand;        // [slices, nq, cap] (score bits, row)
  int* cand_count;    // [slices, nq]
  float* slice_thr;   // [slices, nq] kp-th best of the slice (scaled units) or -inf if nothing dropped
  float* dbg_scores;  // optional [nq, n] raw tensor-core scores (tests only), else nullptr
  uint32_t* lvl;      // [slices, nq] ordered key of the lvl_r-th best score each slice has seen (0 = none yet)
  int lvl_r;          // slices * lvl_r >= kp; 0 disables the cross-slice level
  int nq, d, num_kc;
  int n;
  int num_tiles, slices, tiles_per_slice;
  int kp, cap, keep_max;
  int normalize;
};

__device__ __forceinline__ uint32_t float_to_key(uint32_t b) {
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ uint32_t key_to_float_bits(uint32_t k) {
  return (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
}

// ||x||^2 of one fp32 row, accumulated in fp64 in a fixed order (four interleaved partial sums),
// rounded once to fp32: the `nr` of FAISS' fvec_renorm_L2.  Shared by the fused prologue of the
// scan kernel and by normalize_l2_kernel so both produce bit-identical normalised queries.
__device__ __forceinline__ float row_norm_sq(const float* xr, int d) {
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  int i = 0;
  for (; i + 3 < d; i += 4) {
    const double a = xr[i], b = xr[i + 1], c = xr[i + 2], e = xr[i + 3];
    s0 = fma(a, a, s0);
    s1 = fma(b, b, s1);
    s2 = fma(c, c, s2);
    s3 = fma(e, e, s3);
  }
  for (; i < d; ++i) {
    const double a = xr[i];
    s0 = fma(a, a, s0);
  }
  return static_cast<float>((s0 + s1) + (s2 + s3));
}
// inv_nr = 1.0 / sqrtf(nr) evaluated in double and rounded to float, as FAISS writes it.
__device__ __forceinline__ float inv_norm(float nr) {
  return nr > 0.0f ? static_cast<float>(1.0 / static_cast<double>(sqrtf(nr))) : 1.0f;
}

// In-place faiss.normalize_L2 (engine.py:242): one thread per row.
__global__ void __launch_bounds__(128) normalize_l2_kernel(float* __restrict__ x, int nq, int d) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  float* xr = x + static_cast<size_t>(q) * d;
  const float inv = inv_norm(row_norm_sq(xr, d));
  for (int i = 0; i < d; ++i) xr[i] = xr[i] * inv;
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
  float thr;  // scores <= thr are dropped
  int cnt;
};

// Compacts the lists of every lane whose list is fuller than `limit` (warp-cooperative).
__device__ __forceinline__ void compact_full_lists(ListState& ls, int limit, int kp, int keep_max, int lane) {
  uint32_t need = __ballot_sync(0xffffffffu, ls.cnt > limit);
  while (need) {
    const int src = __ffs(need) - 1;
    need &= need - 1;
    uint2* b = reinterpret_cast<uint2*>(
        __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(ls.buf), src));
    const int bc = __shfl_sync(0xffffffffu, ls.cnt, src);
    int kept;
    const float tnew = warp_compact(b, bc, kp, keep_max, lane, &kept);
    if (lane == src) {
      ls.thr = tnew;
      ls.cnt = kept;
    }
  }
}

// lvl_r-th largest score of a thread's own list (thread-private walk; lists are L2 resident).
// Returns -inf when the list is shorter than r.  r <= 8.
__device__ __forceinline__ float own_rth_best(const uint2* __restrict__ buf, int cnt, int r) {
  float t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) t[i] = -CUDART_INF_F;
  for (int i = 0; i < cnt; ++i) {
    const float v = __uint_as_float(__ldcg(buf + i).x);
    if (v > t[7]) {
      t[7] = v;
#pragma unroll
      for (int k = 7; k > 0; --k) {
        const float hi = fmaxf(t[k - 1], t[k]), lo = fminf(t[k - 1], t[k]);
        t[k - 1] = hi;
        t[k] = lo;
      }
    }
  }
  float out = t[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) out = (r - 1 == i) ? t[i] : out;
  return out;
}

// Cross-slice level: every slice publishes the r-th best score it has seen;
// with slices * r >= kp, at least kp corpus rows score >= the minimum of the published values, so
// no row below that minimum can be among the query's best kp - whichever slice it lives in.
__device__ __forceinline__ void publish_level(const ScanParams& p, int slice, int q, const ListState& ls) {
  const float v = own_rth_best(ls.buf, ls.cnt, p.lvl_r);
  if (v > -CUDART_INF_F) __stcg(p.lvl + static_cast<size_t>(slice) * p.nq + q, float_to_key(__float_as_uint(v)));
}
__device__ __forceinline__ void refresh_level(const ScanParams& p, int q, ListState& ls) {
  uint32_t lo = 0xFFFFFFFFu;
  for (int s = 0; s < p.slices; ++s) lo = min(lo, __ldcg(p.lvl + static_cast<size_t>(s) * p.nq + q));
  if (lo != 0u) ls.thr = fmaxf(ls.thr, __uint_as_float(key_to_float_bits(lo)));
}

// Cold path of the epilogue: (re)load one 32-column chunk of the accumulator from TMEM and append
// the scores above the thread's threshold to its list.  Lists are compacted once per tile, AFTER
// the accumulator has been handed back to the MMA warp, so a compaction overlaps the next MMAs.
__device__ __noinline__ void append_chunk(uint32_t taddr, ListState& ls, int base_row, int valid,
                                          float* __restrict__ dbg_row) {
  uint32_t r[32];
  ptx::tmem_ld_32x32b_x32(taddr, r);
  ptx::tc_wait_ld();
  if (dbg_row != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < valid) dbg_row[base_row + j] = __uint_as_float(r[j]);
  }
  const float thr = ls.thr;
  int cnt = ls.cnt;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < valid && __uint_as_float(r[j]) > thr) {
      __stcg(ls.buf + cnt, make_uint2(r[j], static_cast<uint32_t>(base_row + j)));
      ++cnt;
    }
  }
  ls.cnt = cnt;
}

// N_T: corpus rows per accumulator tile (UMMA N).  128 when the fp16 queries need <= 256 TMEM
// columns (d <= 512), 64 when they need up to 384 (d <= 768): A + 2 accumulators <= 512 columns.
// kPair: two CTAs of a cluster (two adjacent blocks of 128 queries) run ONE tcgen05.mma
// cta_group::2 stream (M = 256): each CTA stages only N_T/2 rows of every corpus tile, so a corpus
// byte crosses L2 -> SM once per 256 queries instead of once per 128.
template <int N_T, bool kPair>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_topk_kernel(const __grid_constant__ CUtensorMap tmap, const ScanParams p) {
  constexpr int kBoxRows = kPair ? N_T / 2 : N_T;       // corpus rows this CTA stages per tile
  constexpr int kBoxBytes = kBoxRows * 128;             // one TMA box: kBoxRows rows x 64 fp16
  constexpr int kStageBytesT = kPair ? kStageBytes / 2 : kStageBytes;
  constexpr int kKcPerStage = kStageBytesT / kBoxBytes;  // k-chunks (boxes) per pipeline stage
  constexpr int kStages = kStageRing / kStageBytesT;
  constexpr int kChunksPerTile = N_T / 32;
  constexpr uint32_t kACol0 = 2 * N_T;  // TMEM columns: [0,N_T) acc0, [N_T,2N_T) acc1, then A
  constexpr uint32_t kIdesc = ptx::make_idesc_f16(kPair ? 256 : 128, N_T);
  constexpr uint32_t kArrivals = (kPair ? 2 : 1) * (kEpiThreads / 32);  // one arrival per epilogue warp

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ __align__(8) uint64_t a_ready_bar;
  __shared__ uint32_t tmem_base_holder;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qblock = blockIdx.x;
  const int slice = blockIdx.y;
  const uint32_t rank = kPair ? ptx::cluster_ctarank() : 0u;  // position in the CTA pair

  // 1024-byte aligned stage ring (SWIZZLE_128B atoms are 1024 bytes).
  const uint32_t ring_u32 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* ring = smem_raw + (ring_u32 - ptx::smem_u32(smem_raw));

  const int tile_begin = slice * p.tiles_per_slice;
  const int tile_end = min(p.num_tiles, tile_begin + p.tiles_per_slice);
  const int my_tiles = max(0, tile_end - tile_begin);

  if (warp == 5 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], kArrivals);
    }
    ptx::mbar_init(&a_ready_bar, kArrivals);
    ptx::fence_barrier_init();
  }
  if (warp == 4) {
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

  if (warp == 4) {
    // ------------------------------------------------------------ TMA producer
    // The whole warp walks the loop (warp-uniform control flow); one elected lane issues.
    // Pair mode: both CTAs load their half of the tile; the bytes of both halves are counted on
    // the even CTA's full barrier (it issues the MMAs), which therefore expects 2x the bytes.
    uint32_t stage = 0, phase = 0;
    for (int t = tile_begin; t < tile_end; ++t) {
      const int row = t * N_T + static_cast<int>(rank) * kBoxRows;
      for (int kc0 = 0; kc0 < p.num_kc; kc0 += kKcPerStage) {
        const int nb = min(kKcPerStage, p.num_kc - kc0);
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (ptx::elect_one()) {
          if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], nb * kBoxBytes * (kPair ? 2 : 1));
          uint8_t* dst = ring + stage * kStageBytesT;
#pragma unroll
          for (int b = 0; b < kKcPerStage; ++b) {
            if (b < nb) {
              if constexpr (kPair)
                ptx::tma_load_2d_pair(dst + b * kBoxBytes, &tmap, (kc0 + b) * kKC, row, &full_bar[stage],
                                      ptx::kEvictNormal);
              else
                ptx::tma_load_2d(dst + b * kBoxBytes, &tmap, (kc0 + b) * kKC, row, &full_bar[stage],
                                 ptx::kEvictNormal);
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
  } else if (warp == 5) {
    // -------------------------------------------------------------- MMA issuer
    if (rank == 0) {
      ptx::mbar_wait(&a_ready_bar, 0);
      ptx::tc_fence_after();
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const uint32_t acc = it & 1;
        ptx::mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * N_T;
        for (int kc0 = 0; kc0 < p.num_kc; kc0 += kKcPerStage) {
          const int nb = min(kKcPerStage, p.num_kc - kc0);
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t stage_addr = ring_u32 + stage * kStageBytesT;
#pragma unroll
            for (int b = 0; b < kKcPerStage; ++b) {
              if (b < nb) {
                const uint64_t bdesc = ptx::make_kmajor_sw128_desc(stage_addr + b * kBoxBytes);
                const uint32_t a_tmem = tmem_base + kACol0 + (kc0 + b) * 32;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                  // K advances 16 fp16 = 32 bytes inside the swizzle atom (+2 in >>4 units) and
                  // 8 TMEM columns in A.
                  const uint32_t accum = (kc0 | b | k4) != 0 ? 1u : 0u;
                  if constexpr (kPair)
                    ptx::mma_f16_ts_pair(d_tmem, a_tmem + k4 * 8, bdesc + static_cast<uint64_t>(k4 * 2), kIdesc,
                                         accum);
                  else
                    ptx::mma_f16_ts(d_tmem, a_tmem + k4 * 8, bdesc + static_cast<uint64_t>(k4 * 2), kIdesc, accum);
                }
              }
            }
            // stage reusable (in both CTAs) once these MMAs retire
            const bool last = kc0 + kKcPerStage >= p.num_kc;
            if constexpr (kPair) {
              ptx::tc_commit_pair(&empty_bar[stage], 3);
              if (last) ptx::tc_commit_pair(&tmem_full_bar[acc], 3);
            } else {
              ptx::tc_commit(&empty_bar[stage]);
              if (last) ptx::tc_commit(&tmem_full_bar[acc]);
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
  } else {
    // ------------------------------------------- epilogue warps: one query per thread
    const int t = threadIdx.x;  // TMEM lane
    const int q = qblock * kQueryBlock + t;
    const bool live = q < p.nq;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const int d = p.d;

    // ---- fused normalize_L2 (FAISS fvec_renorm_L2: nr = ||x||^2; if nr > 0: x *= 1/sqrt(nr))
    float inv = 1.0f, sq = 1.0f, nrm = 0.0f;
    const float* xr = p.x + static_cast<size_t>(live ? q : 0) * d;
    if (live) {
      if (p.normalize) inv = inv_norm(row_norm_sq(xr, d));
      float amax = 0.0f;
      double n2 = 0;
      for (int j = 0; j < d; ++j) {
        const float v = __ldg(xr + j) * inv;
        amax = fmaxf(amax, fabsf(v));
        n2 = fma(static_cast<double>(v), static_cast<double>(v), n2);
      }
      nrm = static_cast<float>(sqrt(n2)) * 1.0000002f;  // rounded up: used only in the error bound
      // power-of-two scale so that max|x| lands in [1,2): exact, keeps fp16 out of the subnormals
      if (amax > 0.0f && amax < CUDART_INF_F) sq = ldexpf(1.0f, -ilogbf(amax));
      if (slice == 0) {
        p.qscale[q] = sq;
        p.qnorm[q] = nrm;
      }
    }
    for (int kc = 0; kc < p.num_kc; ++kc) {
      uint32_t r[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int i0 = kc * kKC + 2 * j;
        float v0 = 0.0f, v1 = 0.0f;
        if (live) {
          if (i0 < d) v0 = __ldg(xr + i0) * inv;
          if (i0 + 1 < d) v1 = __ldg(xr + i0 + 1) * inv;
          if (slice == 0) {
            if (i0 < d) p.xn[static_cast<size_t>(q) * d + i0] = v0;
            if (i0 + 1 < d) p.xn[static_cast<size_t>(q) * d + i0 + 1] = v1;
          }
        }
        const __half2 h = __floats2half2_rn(v0 * sq, v1 * sq);  // .x (low half) = even k
        r[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
      ptx::tmem_st_32x32b_x32(tmem_base + lane_base + kACol0 + kc * 32, r);
    }
    ptx::tc_wait_st();
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if constexpr (kPair) ptx::mbar_arrive_cluster(&a_ready_bar, 0); else ptx::mbar_arrive(&a_ready_bar);
    }

    // ---- threshold scan
    const size_t list = static_cast<size_t>(slice) * p.nq + (live ? q : 0);
    ListState ls;
    ls.buf = p.cand + list * p.cap;
    ls.thr = live ? -CUDART_INF_F : CUDART_INF_F;
    ls.cnt = 0;
    const int kp = p.kp, cap = p.cap, keep_max = p.keep_max, n = p.n;
    float* dbg_row = (p.dbg_scores != nullptr && live) ? p.dbg_scores + static_cast<size_t>(q) * n : nullptr;
    const bool dbg = p.dbg_scores != nullptr;

    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t acc = it & 1;
      ptx::mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
      ptx::tc_fence_after();
      const int row0 = (tile_begin + it) * N_T;
      const uint32_t tile_addr = tmem_base + lane_base + acc * N_T;
      // software pipeline over the tile's 32-column chunks: chunk c+1 is in flight while c is compared
      uint32_t r[2][32];
      ptx::tmem_ld_32x32b_x32(tile_addr, r[0]);
#pragma unroll
      for (int c = 0; c < kChunksPerTile; ++c) {
        ptx::tc_wait_ld();
        if (c + 1 < kChunksPerTile) ptx::tmem_ld_32x32b_x32(tile_addr + (c + 1) * 32, r[(c + 1) & 1]);
        const int base_row = row0 + c * 32;
        if (base_row < n) {  // warp-uniform: otherwise the chunk is TMA zero fill only
          const int valid = min(32, n - base_row);
          bool hit = false;
#pragma unroll
          for (int j = 0; j < 32; ++j) hit |= (__uint_as_float(r[c & 1][j]) > ls.thr);
          if (valid < 32 || dbg || __any_sync(0xffffffffu, hit)) {
            ptx::tc_wait_ld();  // settle the prefetched chunk before the call (registers may be saved)
            append_chunk(tile_addr + c * 32, ls, base_row, valid, dbg_row);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // one arrival per warp (in pair mode the odd CTA's arrivals are remote)
        if constexpr (kPair) ptx::mbar_arrive_cluster(&tmem_empty_bar[acc], 0); else ptx::mbar_arrive(&tmem_empty_bar[acc]);
      }
      // a tile appends at most N_T entries per list: keep that much room for the next one
      compact_full_lists(ls, cap - N_T, kp, keep_max, lane);
      if (p.lvl_r > 0 && live) {
        const int done = it + 1;
        if ((done & (done - 1)) == 0) publish_level(p, slice, q, ls);  // after tiles 1, 2, 4, 8, ...
        if ((done & (done - 1)) == 0 || (done & 7) == 0) refresh_level(p, q, ls);
      }
    }
    // ---- final compaction to exactly the slice's top-kp
    compact_full_lists(ls, kp, kp, kp, lane);
    if (live) {
      p.cand_count[list] = ls.cnt;
      p.slice_thr[list] = ls.thr;
    }
  }

  ptx::tc_fence_before();
  if constexpr (kPair) ptx::cluster_sync(); else __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    if constexpr (kPair) ptx::tmem_dealloc_pair(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace lxg
