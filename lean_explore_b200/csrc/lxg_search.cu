// C-ABI entry points of the flat inner-product index (include/lxg.h).  Host logic only:
// workspace sizing, slice/tile partitioning, tensor-map construction, launches.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lxg.h"
#include "common.h"
#include "rescore.cuh"
#include "scan_topk.cuh"

namespace lxg {

thread_local std::string g_last_error;
int set_error(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

static std::mutex g_init_mutex;
static bool g_inited = false;
constexpr int kMaxDevices = 64;
static int g_sms[kMaxDevices] = {0};  // SM count of every device lxg_init has brought up (0 = not initialised)
static bool g_no_level = false;      // LXG_SCAN_NOLEVEL=1: no cross-slice level (A/B measurements)
static int g_lvl_sleep_ns = 0;        // LXG_LVL_SLEEP=<ns>: pause between level-warp rounds (experiments; 0 = the built-in schedule)
static bool g_debug_counts = false;  // LXG_DEBUG_COUNTS=1: candidates per query after pass 1 -> stderr (synchronises)
static int g_perf_mode = 0;          // LXG_SCAN_PERF_MODE: pipeline measurements with a crippled epilogue (wrong results)
static bool g_asmem_768 = true;      // LXG_SCAN_ASMEM=0: 512 < d <= 768 falls back to 64-row tiles, all of A in tensor memory (A/B)
static int g_sync_mb = 28;           // LXG_SCAN_SYNC_MB: L2 megabytes the readers' spread may cover (all slices together)
static bool g_zero_copy = true;      // LXG_ZERO_COPY=0: pinned host queries / results go through staging copies
static bool g_sync_readers = true;   // LXG_SCAN_SYNC=0: the readers of a corpus slice are not kept in step (A/B)
static bool g_force_single = false;  // LXG_SCAN_SINGLE=1: never pair CTAs (A/B measurements, tests)
static bool g_no_split_merge = false;  // LXG_MERGE_SPLIT=0 / lxg_debug_config perf_mode 16: small batches keep the one-kernel merge
constexpr int kSplitMergeQueries = 8;  // up to this many queries with k' >= 256 run the merge in three stages (rescore.cuh)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;

int num_sms() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
  return g_sms[dev];
}
int device_of_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? a.device : -1;
}
cudaError_t ensure_dyn_smem(const void* func, size_t bytes) {
  if (bytes <= 48u * 1024u) return cudaSuccess;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> have;
  std::lock_guard<std::mutex> lock(mu);
  size_t& cur = have[std::make_pair(dev, func)];
  if (bytes <= cur) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e == cudaSuccess) cur = bytes;
  return e;
}
bool encode_tensor_map_ready() { return g_encode_tiled != nullptr; }
CUresult encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dt, cuuint32_t rank, void* base,
                           const cuuint64_t* gdim, const cuuint64_t* gstride, const cuuint32_t* box,
                           const cuuint32_t* estr, CUtensorMapInterleave il, CUtensorMapSwizzle sw,
                           CUtensorMapL2promotion l2, CUtensorMapFloatOOBfill oob) {
  return g_encode_tiled(map, dt, rank, base, gdim, gstride, box, estr, il, sw, l2, oob);
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Page-locked host memory (cudaHostAlloc / cudaHostRegister, e.g. a pinned torch tensor): DMA can
// read and write it directly, so the pinned staging bounce is skipped.
bool is_pinned_host_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// Device-side alias of page-locked, mapped host memory (what cudaHostAlloc / torch's pin_memory give
// under unified addressing), or nullptr.  Kernels can read the queries from it and write the results
// into it directly: no staging copies, and the flag read-back is the call's only synchronisation.
void* mapped_host_alias(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

// One growable device allocation.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    const size_t want = need + need / 4;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) bytes = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};
struct HostBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMallocHost(&p, need + need / 4);
    if (e == cudaSuccess) bytes = need + need / 4;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
  }
};

}  // namespace lxg

using namespace lxg;

struct lxg_index {
  CorpusView cv{};
  __half* scan = nullptr;  // fp16 rows the tensor cores stream (== cv.rows when it can alias)
  bool scan_owned = false;
  int scan_pitch = 0;      // elements
  int device = 0;          // the GPU that holds the corpus; made current by every entry point
  int sms = 0;
  cudaEvent_t done_ev = nullptr;  // end of the last search on this handle: the next one (any stream, any
  bool have_done = false;         // thread) waits for it before it touches the shared workspaces
  int* last_flags = nullptr;      // device flag_count / overflow words of the last search
  int tile_rows = 0;       // N_T
  int num_kc = 0;
  int a_smem_chunks = 0;    // k-chunks of the query block the scan keeps in shared memory (kASm)
  CUtensorMap tmap{};       // box = tile_rows rows   (one CTA per query block)
  CUtensorMap tmap_pair{};  // box = tile_rows/2 rows (CTA pairs, tcgen05 cta_group::2)
  std::mutex mu;
  DevBuf ws_cand, ws_small, ws_x, ws_out, ws_exact;
  HostBuf h_stage;
  lxg_search_stats stats{};
  bool timing = false;
  std::vector<cudaEvent_t> ev_pool;   // 5 events per timed call
  size_t ev_used = 0;
  ~lxg_index() {
    if (done_ev) cudaEventDestroy(done_ev);
  }
};

namespace {

int build_tensor_map(lxg_index* ix) {
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(ix->cv.d), static_cast<cuuint64_t>(ix->cv.n)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ix->scan_pitch) * sizeof(__half)};
  cuuint32_t estr[2] = {1, 1};
  for (int pair = 0; pair < 2; ++pair) {
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kKC), static_cast<cuuint32_t>(ix->tile_rows >> pair)};
    CUresult r = g_encode_tiled(pair ? &ix->tmap_pair : &ix->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ix->scan,
                                gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return set_error(LXG_ECUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(r));
  }
  return LXG_OK;
}

int candidates_per_slice(int k, float rel_err, int d, long long n) {
  // kp = k plus a margin (so that the exactness certificate almost always holds), rounded to 32.
  // The certificate needs exact score(rank k) - tensor score(rank kp) > eps = rel_err |c| |q|.  For
  // isotropic rows the scores near rank r sit ~ sigma ln(kp / k) / sqrt(2 ln(n / k)) apart with
  // sigma = |c| |q| / sqrt(d), so the margin that keeps that gap at 3 eps is
  // ln(kp / k) = 3 rel_err sqrt(d) sqrt(2 ln(n / k)).  It only exceeds the k / 4 floor for fp32
  // corpora at large d (rel_err doubles: the scan copy is rounded too) - the shipped 1024-d index,
  // where k' = 64 left ~20 % of a 1024-query batch to the exhaustive exact path (10 ms).
  int margin = std::max(14, k / 4);
  if (n > 4LL * k) {
    const double g = 3.0 * rel_err * std::sqrt(static_cast<double>(d)) * std::sqrt(2.0 * std::log(static_cast<double>(n) / k));
    margin = std::max(margin, static_cast<int>(std::ceil(k * (std::exp(std::min(g, 1.0)) - 1.0))));
  }
  return ((k + margin + 31) / 32) * 32;
}

struct Plan {
  int kp, cap, keep_max, qblocks, slices, tiles_per_slice, num_tiles;
  bool pair;      // CTA pairs (cta_group::2) when there are at least two query blocks
  int grid_x;     // query blocks launched (padded to even in pair mode)
  int lists;      // candidate lists per query: one per slice and epilogue group
  int lvl_r;      // cross-list level: tracker depth, every list publishes its lvl_r best (0 = disabled)
  int lvl_lg;     // log2 of the tracker ranks per list the level warps read
  uint32_t lvl_slot, lvl_w;  // per rank class: tracker slot, rows it stands for (4 bits each)
  int max_items;  // pass-2 candidate pool (entries)
};

// Lists that see at least 8*kTrack rows, i.e. that can fill a tracker (scan_topk.cuh).  List
// (slice, g) holds columns [g*gc, (g+1)*gc) of every tile of the slice, gc = tile_rows / kGroups.
int lists_with_level(int n, int tile_rows, int num_tiles, int slices, int tiles_per_slice, int stride) {
  const int gc = tile_rows / kGroups;
  int ok = 0;
  for (int s = 0; s < slices; ++s) {
    const int tb = s * tiles_per_slice;
    const int te = std::min(num_tiles, tb + tiles_per_slice);
    const int my = std::max(0, te - tb);
    for (int g = 0; g < kGroups; ++g) {
      if ((s * kGroups + g) % stride != 0) continue;
      long long rows = static_cast<long long>(my) * gc;
      if (my > 0 && te == num_tiles) {
        const long long first = static_cast<long long>(num_tiles - 1) * tile_rows + g * gc;
        rows -= gc - std::max(0ll, std::min(static_cast<long long>(gc), static_cast<long long>(n) - first));
      }
      if (rows >= kTrack * 8) ++ok;
    }
  }
  return ok;
}

// What a plan depends on (a handle's shape, or the arguments of lxg_debug_plan).
struct PlanInput {
  int tile_rows;
  float rel_err;
  int d;
  long long n;
  int sms;
};
// Tile rows and error bound of a corpus shape - shared by lxg_index_create and lxg_debug_plan.
int tile_rows_for(int num_kc) {
  const int a_smem = num_kc <= 8 ? 0 : (num_kc > 12 ? 8 : (g_asmem_768 ? 4 : 0));
  return (num_kc <= 8 || a_smem > 0) ? 128 : 64;
}
float rel_err_for(int d, int dtype) {
  // |tensor-core score - exact score| <= rel_err * ||row|| * ||query||  (certificate)
  const float u16 = 1.0f / 2048.0f;  // fp16 round-to-nearest unit
  float rel = u16 * 1.001f;          // query fp32 -> fp16
  if (dtype == LXG_F32) rel += u16 * 1.001f;  // corpus fp32 -> fp16 scan copy
  rel += 1.0f / 65536.0f;                     // fp32 accumulation inside the tensor core
  rel += std::sqrt(static_cast<float>(d)) * (1.0f / 8388608.0f);  // fp16 subnormal flush of tiny elements
  return rel;
}

Plan make_plan(const PlanInput& in, int nq, int k) {
  Plan pl;
  const int nt = in.tile_rows;
  pl.kp = candidates_per_slice(k, in.rel_err, in.d, in.n);
  pl.qblocks = (nq + kQueryBlock - 1) / kQueryBlock;
  pl.num_tiles = static_cast<int>((in.n + nt - 1) / nt);
  pl.pair = pl.qblocks >= 2 && !g_force_single;
  pl.grid_x = pl.pair ? (pl.qblocks + 1) / 2 * 2 : pl.qblocks;
  auto slice_up = [&](int s) {
    s = std::min(s, 148);
    s = std::min(s, std::max(1, pl.num_tiles));
    s = std::max(s, 1);
    pl.tiles_per_slice = (pl.num_tiles + s - 1) / s;
    pl.slices = std::max(1, (pl.num_tiles + pl.tiles_per_slice - 1) / std::max(1, pl.tiles_per_slice));
    pl.lists = kGroups * pl.slices;
  };
  slice_up(std::max(1, in.sms / pl.grid_x));
  // cross-list level = the kp-th largest of the union of the lists' trackers (scan_topk.cuh): needs
  // lists * depth >= kp with depth <= kTrack.  The depth is the smallest that gives the union ~2 kp
  // entries (few lists hold more than twice their share of a query's best kp).  With hundreds of lists
  // (one or two query blocks) and a deep tracker the level warps read four ranks per list - 1st, 2nd,
  // 4th, 8th best standing for 1, 1, 2, 4 rows - to keep a query's words within kLvlMaxWords.
  const int lv = g_no_level ? 0 : lists_with_level(in.n, nt, pl.num_tiles, pl.slices, pl.tiles_per_slice, 1);
  pl.lvl_r = 0;
  pl.lvl_lg = 0;
  pl.lvl_slot = pl.lvl_w = 0;
  if (lv >= 2 && static_cast<long long>(lv) * kTrack >= pl.kp) {
    int depth = 1;
    while (depth < kTrack && static_cast<long long>(lv) * depth < 2LL * pl.kp) depth *= 2;
    int ncls = depth;
    while (pl.lists * ncls > kLvlMaxWords) ncls /= 2;
    pl.lvl_r = depth;
    for (int c = 0; c < ncls; ++c) {
      // class c = rank (1-based): every rank when ncls == depth, else depth >> (ncls - 1 - c)
      const int rank = ncls == depth ? c + 1 : depth >> (ncls - 1 - c);
      const int prev = c == 0 ? 0 : (ncls == depth ? c : depth >> (ncls - c));
      pl.lvl_slot |= static_cast<uint32_t>(kTrack - depth + rank - 1) << (4 * c);
      pl.lvl_w |= static_cast<uint32_t>(rank - prev) << (4 * c);
    }
    while ((1 << pl.lvl_lg) < ncls) ++pl.lvl_lg;
  }
  if (pl.lvl_r > 0) {
    // lists only grow (a few hundred entries); a list that does fill up is compacted exactly
    pl.cap = std::max(1024, pl.kp + 2 * nt);
    pl.keep_max = pl.kp;
    // typical survivors: a few k' over all lists; overflow is tightened exactly
    pl.max_items = std::max(2048, 4 * pl.kp);
    // a handful of queries (the 1024-thread merge CTA has an SM's shared memory to itself): room for
    // 12 k' entries - with k' = 1408 (faiss_k = 1000) and ~290 lists publishing their 5th best, 4 k'
    // overflowed and a third of the merge went into tightening the level over global memory
    if (nq <= 64) pl.max_items = std::min(20480, std::max(pl.max_items, 12 * pl.kp));
  } else {
    // thresholds come from compacting full lists: pass 2 holds lists * kp keys in shared memory
    slice_up(std::min(pl.slices, std::max(1, 24576 / kGroups / pl.kp)));
    // list capacity: room for many appends between two compactions (each one costs a warp ~1-2k
    // cycles); a mid-scan compaction may keep up to keep_max entries (cheaper inexact cut)
    pl.cap = pl.kp <= 64 ? 4 * pl.kp : 2 * pl.kp;
    pl.cap = std::max(pl.cap, pl.kp + 2 * nt);  // a whole tile is appended between compactions
    pl.keep_max = pl.kp + std::max(16, pl.kp / 2);
    if (pl.keep_max > pl.cap - nt - 32) pl.keep_max = pl.kp;
    pl.max_items = pl.lists * pl.kp;
  }
  return pl;
}

template <int N_T, bool kPair, int kASm = 0>
cudaError_t launch_scan(const lxg_index* ix, const ScanParams& sp, int grid_x, cudaStream_t st) {
  const int smem = (kASm == 0 ? kStageRing : kScanSmemMax) + 1024;
  {
    cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(&scan_topk_kernel<N_T, kPair, kASm>), smem);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid_x, sp.slices, 1);
  cfg.blockDim = dim3(kScanThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, scan_topk_kernel<N_T, kPair, kASm>, kPair ? ix->tmap_pair : ix->tmap, sp);
}

template <int THREADS>
cudaError_t launch_merge(int nq, size_t smem, const MergeParams& mp, const CorpusView& cv, cudaStream_t st) {
  {
    cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(&merge_rescore_kernel<THREADS>), smem);
    if (e != cudaSuccess) return e;
  }
  merge_rescore_kernel<THREADS><<<nq, THREADS, smem, st>>>(mp, cv);
  return cudaGetLastError();
}

}  // namespace

extern "C" {

int lxg_abi_version(void) { return 3; }

const char* lxg_last_error(void) { return g_last_error.c_str(); }

int lxg_init(int device) {
  std::lock_guard<std::mutex> lock(g_init_mutex);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return set_error(LXG_ENODEVICE,
                     std::string("no CUDA device visible (") + cudaGetErrorString(e) +
                         "); the lxg kernels are sm_100a only and there is no CPU fallback");
  }
  if (device < 0 || device >= count) return set_error(LXG_EINVAL, "device index out of range");
  cudaDeviceProp prop;
  LXG_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return set_error(LXG_ENODEVICE, std::string("device is ") + prop.name + " (sm_" +
                                        std::to_string(prop.major) + std::to_string(prop.minor) +
                                        "); lxg needs sm_100 (B200)");
  LXG_CUDA(cudaSetDevice(device));
  LXG_CUDA(cudaFree(nullptr));
  if (!g_encode_tiled) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    LXG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
      return set_error(LXG_ECUDA, "driver does not export cuTensorMapEncodeTiled");
    g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  }
  if (device >= kMaxDevices) return set_error(LXG_EINVAL, "device index out of range");
  g_sms[device] = prop.multiProcessorCount;
  const char* fs = std::getenv("LXG_SCAN_SINGLE");
  g_force_single = fs && fs[0] == '1';
  const char* zc = std::getenv("LXG_ZERO_COPY");
  g_zero_copy = !(zc && zc[0] == '0');
  const char* sr = std::getenv("LXG_SCAN_SYNC");
  g_sync_readers = !(sr && sr[0] == '0');
  const char* smb = std::getenv("LXG_SCAN_SYNC_MB");
  if (smb && std::atoi(smb) > 0) g_sync_mb = std::atoi(smb);
  const char* am = std::getenv("LXG_SCAN_ASMEM");
  g_asmem_768 = !(am && am[0] == '0');
  const char* nl = std::getenv("LXG_SCAN_NOLEVEL");
  g_no_level = nl && nl[0] == '1';
  const char* ms = std::getenv("LXG_MERGE_SPLIT");
  g_no_split_merge = ms && ms[0] == '0';
  const char* pm = std::getenv("LXG_SCAN_PERF_MODE");
  g_perf_mode = pm ? std::atoi(pm) : 0;
  const char* ls = std::getenv("LXG_LVL_SLEEP");
  g_lvl_sleep_ns = ls ? std::atoi(ls) : 0;
  const char* dc = std::getenv("LXG_DEBUG_COUNTS");
  g_debug_counts = dc && dc[0] == '1';
  g_inited = true;
  return LXG_OK;
}

int lxg_debug_config(int no_level, int force_single, int perf_mode) {
  if (no_level >= 0) g_no_level = no_level != 0;
  if (force_single >= 0) g_force_single = force_single != 0;
  if (perf_mode == 16 || perf_mode == 17) {
    g_no_split_merge = perf_mode == 16;  // 16: one-kernel merge for every batch size, 17: back to the default
    return LXG_OK;
  }
  if (perf_mode >= 0) g_perf_mode = perf_mode;
  return LXG_OK;
}

int lxg_debug_plan(int64_t n, int32_t d, int dtype, int32_t sms, int32_t nq, int32_t k, lxg_plan_info* out) {
  if (!out || n < 0 || d <= 0 || d > 1024 || sms <= 0 || nq <= 0 || k <= 0 || (dtype != LXG_F32 && dtype != LXG_F16))
    return set_error(LXG_EINVAL, "bad argument");
  const int num_kc = (d + kKC - 1) / kKC;
  const Plan pl = make_plan(PlanInput{tile_rows_for(num_kc), rel_err_for(d, dtype), d, static_cast<long long>(n), sms}, nq, k);
  *out = lxg_plan_info{};
  out->kp = pl.kp;
  out->query_blocks = pl.qblocks;
  out->slices = pl.slices;
  out->lists = pl.lists;
  out->tile_rows = tile_rows_for(num_kc);
  out->pair = pl.pair ? 1 : 0;
  out->level_depth = pl.lvl_r;
  out->level_classes = pl.lvl_r > 0 ? 1 << pl.lvl_lg : 0;
  for (int c = 0; c < out->level_classes; ++c) {
    out->level_rank[c] = static_cast<int32_t>((pl.lvl_slot >> (4 * c)) & 15u) - (kTrack - pl.lvl_r) + 1;
    out->level_weight[c] = static_cast<int32_t>((pl.lvl_w >> (4 * c)) & 15u);
  }
  out->list_capacity = pl.cap;
  out->merge_pool = pl.max_items;
  return LXG_OK;
}

int lxg_index_create(lxg_index** out, const void* corpus_dev, int64_t n, int32_t d, int dtype,
                     int64_t row_offset) {
  if (!out) return set_error(LXG_EINVAL, "out is NULL");
  *out = nullptr;
  if (!g_inited) return set_error(LXG_EINVAL, "lxg_init has not been called");
  if (n < 0 || d <= 0 || (dtype != LXG_F32 && dtype != LXG_F16))
    return set_error(LXG_EINVAL, "bad n / d / dtype");
  if (n >= (1ll << 31) - 256) return set_error(LXG_EUNSUPPORTED, "more than 2^31 rows per shard");
  if (d > 1024)
    return set_error(LXG_EUNSUPPORTED,
                     "d > 1024: the fp16 query block must fit tensor memory + shared memory next to the corpus pipeline");
  if (n > 0 && (!corpus_dev || !is_device_ptr(corpus_dev)))
    return set_error(LXG_EINVAL, "corpus_dev must be device memory");
  // the index lives on the device that holds the corpus (an empty index: on the current device)
  int device = n > 0 ? device_of_ptr(corpus_dev) : -1;
  if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return set_error(LXG_ECUDA, "cudaGetDevice failed");
  if (device >= kMaxDevices || g_sms[device] == 0)
    return set_error(LXG_EINVAL, "lxg_init has not been called for device " + std::to_string(device) + " (the one that holds the corpus)");
  DeviceGuard guard(device);
  lxg_index* ix = new lxg_index();
  ix->device = device;
  ix->sms = g_sms[device];
  if (cudaEventCreateWithFlags(&ix->done_ev, cudaEventDisableTiming) != cudaSuccess) {
    delete ix;
    return set_error(LXG_ECUDA, "cudaEventCreate failed");
  }
  ix->cv.rows = corpus_dev;
  ix->cv.pitch = d;
  ix->cv.dtype = dtype;
  ix->cv.n = static_cast<int>(n);
  ix->cv.d = d;
  ix->cv.row_offset = row_offset;
  ix->cv.max_row_norm = 0.f;
  ix->cv.scan_scale = 1.f;
  ix->num_kc = (d + kKC - 1) / kKC;
  // 128-row tiles throughout; the first 8 k-chunks (512 dims) of the query block live in tensor
  // memory, the rest in shared memory: 4 chunks for d <= 768 (measured on cfg3, same box: 2.21 ms
  // against 2.62 ms for the 64-row-tile variant that keeps all of A in tensor memory, which
  // LXG_SCAN_ASMEM=0 still selects), 8 chunks for d <= 1024
  ix->a_smem_chunks = ix->num_kc <= 8 ? 0 : (ix->num_kc > 12 ? 8 : (g_asmem_768 ? 4 : 0));
  ix->tile_rows = tile_rows_for(ix->num_kc);
  const float sqrt_d = std::sqrt(static_cast<float>(d));
  const bool alias = dtype == LXG_F16 && d % 8 == 0 && (reinterpret_cast<uintptr_t>(corpus_dev) % 16 == 0);
  if (n > 0) {
    // statistics for the certificate: max row norm, max |element|
    const int blocks = std::min<int64_t>(4 * ix->sms, (n + 7) / 8);
    DevBuf tmp;
    cudaError_t e = tmp.reserve(blocks * (sizeof(double) + sizeof(float)));
    if (e != cudaSuccess) {
      delete ix;
      return set_error(LXG_ECUDA, cudaGetErrorString(e));
    }
    double* bn = reinterpret_cast<double*>(tmp.p);
    float* ba = reinterpret_cast<float*>(bn + blocks);
    corpus_stats_kernel<<<blocks, 256>>>(corpus_dev, d, dtype, n, d, bn, ba);
    std::vector<double> hn(blocks);
    std::vector<float> ha(blocks);
    e = cudaMemcpy(hn.data(), bn, blocks * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(ha.data(), ba, blocks * sizeof(float), cudaMemcpyDeviceToHost);
    tmp.release();
    if (e != cudaSuccess) {
      delete ix;
      return set_error(LXG_ECUDA, std::string("corpus statistics: ") + cudaGetErrorString(e));
    }
    double n2 = 0;
    float amax = 0;
    for (int i = 0; i < blocks; ++i) {
      n2 = std::max(n2, hn[i]);
      amax = std::max(amax, ha[i]);
    }
    if (!std::isfinite(n2) || !std::isfinite(amax)) {
      delete ix;
      return set_error(LXG_EINVAL, "corpus contains non-finite values");
    }
    ix->cv.max_row_norm = static_cast<float>(std::sqrt(n2) * 1.000001);
    if (alias) {
      ix->scan = const_cast<__half*>(reinterpret_cast<const __half*>(corpus_dev));
      ix->scan_pitch = d;
      ix->cv.scan_scale = 1.f;
    } else {
      ix->scan_pitch = ((d + 7) / 8) * 8;
      float scale = 1.f;
      if (dtype == LXG_F32 && amax > 0.f) scale = std::ldexp(1.f, -std::ilogb(amax));
      ix->cv.scan_scale = scale;
      e = cudaMalloc(&ix->scan, static_cast<size_t>(n) * ix->scan_pitch * sizeof(__half));
      if (e != cudaSuccess) {
        delete ix;
        return set_error(LXG_ECUDA, std::string("scan copy: ") + cudaGetErrorString(e));
      }
      ix->scan_owned = true;
      make_scan_copy_kernel<<<8 * ix->sms, 256>>>(corpus_dev, d, dtype, n, d, ix->scan,
                                                    ix->scan_pitch, scale);
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        cudaFree(ix->scan);
        delete ix;
        return set_error(LXG_ECUDA, std::string("scan copy kernel: ") + cudaGetErrorString(e));
      }
    }
    int rc = build_tensor_map(ix);
    if (rc != LXG_OK) {
      if (ix->scan_owned) cudaFree(ix->scan);
      delete ix;
      return rc;
    }
  }
  ix->cv.rel_err = rel_err_for(d, dtype);
  *out = ix;
  return LXG_OK;
}

int lxg_index_destroy(lxg_index* ix) {
  if (!ix) return LXG_OK;
  DeviceGuard guard(ix->device);
  if (ix->have_done) cudaEventSynchronize(ix->done_ev);
  if (ix->scan_owned && ix->scan) cudaFree(ix->scan);
  for (cudaEvent_t e : ix->ev_pool) cudaEventDestroy(e);
  ix->ws_cand.release();
  ix->ws_small.release();
  ix->ws_x.release();
  ix->ws_out.release();
  ix->ws_exact.release();
  ix->h_stage.release();
  delete ix;
  return LXG_OK;
}

int64_t lxg_index_ntotal(const lxg_index* ix) { return ix ? ix->cv.n : -1; }
int32_t lxg_index_d(const lxg_index* ix) { return ix ? ix->cv.d : -1; }

int lxg_index_set_timing(lxg_index* ix, int enable) {
  if (!ix) return set_error(LXG_EINVAL, "NULL argument");
  std::lock_guard<std::mutex> lock(ix->mu);
  ix->timing = enable != 0;
  ix->ev_used = 0;
  return LXG_OK;
}

int lxg_index_get_timing(lxg_index* ix, lxg_timing* out) {
  if (!ix || !out) return set_error(LXG_EINVAL, "NULL argument");
  DeviceGuard guard(ix->device);
  std::lock_guard<std::mutex> lock(ix->mu);
  *out = lxg_timing{};
  for (size_t i = 0; i + 5 <= ix->ev_used; i += 5) {
    LXG_CUDA(cudaEventSynchronize(ix->ev_pool[i + 4]));
    float p = 0, a = 0, b = 0, c = 0;
    LXG_CUDA(cudaEventElapsedTime(&p, ix->ev_pool[i], ix->ev_pool[i + 1]));
    LXG_CUDA(cudaEventElapsedTime(&a, ix->ev_pool[i + 1], ix->ev_pool[i + 2]));
    LXG_CUDA(cudaEventElapsedTime(&b, ix->ev_pool[i + 2], ix->ev_pool[i + 3]));
    LXG_CUDA(cudaEventElapsedTime(&c, ix->ev_pool[i + 3], ix->ev_pool[i + 4]));
    out->prep_ms += p;
    out->scan_ms += a;
    out->merge_ms += b;
    out->exact_ms += c;
    out->calls += 1;
  }
  ix->ev_used = 0;
  return LXG_OK;
}

int lxg_index_last_stats(const lxg_index* ix, lxg_search_stats* out) {
  if (!ix || !out) return set_error(LXG_EINVAL, "NULL argument");
  *out = ix->stats;
  return LXG_OK;
}

}  // extern "C"

namespace {

// Device-side search of up to 148*128 queries.  All pointers are device memory.
int search_device(lxg_index* ix, const float* x, int nq, int k, int normalize, float* D, long long* I,
                  double* D64, float* dbg_scores, float* dbg_qscale, cudaStream_t st, bool read_flags) {
  const int d = ix->cv.d;
  const Plan pl = make_plan(PlanInput{ix->tile_rows, ix->cv.rel_err, ix->cv.d, ix->cv.n, ix->sms}, nq, k);
  const size_t lists = static_cast<size_t>(pl.lists) * nq;
  const int nq_pad = pl.grid_x * kQueryBlock;
  const int dpad = ix->num_kc * kKC;
  LXG_CUDA(ix->ws_cand.reserve(lists * pl.cap * sizeof(uint2)));
  // small arrays: cand_count, slice_thr, lvl [lists]; qscale, qnorm [nq]; flag_count(1)+overflow(1)+pad,
  // flag_list [nq]; flag_theta [nq] (double)
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) / 256 * 256;
    return o;
  };
  const size_t o_count = take(lists * sizeof(int));
  const size_t o_thr = take(lists * sizeof(float));
  const size_t o_qscale = take(nq * sizeof(float));
  const size_t o_qnorm = take(nq * sizeof(float));
  const size_t o_flist = take(nq * sizeof(int));
  const size_t o_theta = take(nq * sizeof(double));
  // flag_count, overflow, padding (64 words), then the scan's reader-progress counters (1024 words);
  // cleared together with the levels right behind them (prep kernel)
  constexpr int kProgWords = 1024;
  const size_t o_flags = take(256 + kProgWords * sizeof(int));
  // cross-list level: [nq] levels, then [nq, lists, kTrack] published trackers
  const size_t lvl_words = pl.lvl_r > 0 ? (static_cast<size_t>(nq) + 63) / 64 * 64 + lists * kTrack : 0;
  const size_t o_lvl = take(lvl_words * sizeof(uint32_t));
  // split merge (a handful of queries, large k): selected rows and their exact scores between the stages
  int sort_n = 1;
  while (sort_n < pl.kp) sort_n <<= 1;
  const bool split_merge = nq <= kSplitMergeQueries && pl.kp >= 256 && !g_no_split_merge;
  const size_t o_selrow = take(split_merge ? static_cast<size_t>(nq) * sort_n * sizeof(unsigned) : 0);
  const size_t o_selscore = take(split_merge ? static_cast<size_t>(nq) * sort_n * sizeof(double) : 0);
  const size_t o_seln = take(split_merge ? nq * sizeof(int) : 0);
  const size_t o_selamin = take(split_merge ? nq * sizeof(float) : 0);
  LXG_CUDA(ix->ws_small.reserve(off));
  // normalised fp32 queries, then the prepared fp16 query blocks
  const size_t xn_bytes = (static_cast<size_t>(nq) * d * sizeof(float) + 255) / 256 * 256;
  LXG_CUDA(ix->ws_x.reserve(xn_bytes + static_cast<size_t>(nq_pad) * dpad * sizeof(__half)));
  uint8_t* sm = reinterpret_cast<uint8_t*>(ix->ws_small.p);
  int* flag_count = reinterpret_cast<int*>(sm + o_flags);
  int* overflow = flag_count + 1;
  ix->last_flags = flag_count;
  float* xn = reinterpret_cast<float*>(ix->ws_x.p);
  __half* xh = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(ix->ws_x.p) + xn_bytes);
  float* qscale = dbg_qscale ? dbg_qscale : reinterpret_cast<float*>(sm + o_qscale);
  float* qnorm = reinterpret_cast<float*>(sm + o_qnorm);

  ScanParams sp{};
  sp.xh = xh;
  sp.cand = reinterpret_cast<uint2*>(ix->ws_cand.p);
  sp.cand_count = reinterpret_cast<int*>(sm + o_count);
  sp.slice_thr = reinterpret_cast<float*>(sm + o_thr);
  sp.dbg_scores = dbg_scores;
  sp.nq = nq;
  sp.dpad = dpad;
  sp.num_kc = ix->num_kc;
  sp.n = ix->cv.n;
  sp.num_tiles = pl.num_tiles;
  sp.slices = pl.slices;
  sp.tiles_per_slice = pl.tiles_per_slice;
  sp.kp = pl.kp;
  sp.cap = pl.cap;
  sp.keep_max = pl.keep_max;
  sp.lvl = reinterpret_cast<uint32_t*>(sm + o_lvl);
  sp.lvl_r = pl.lvl_r;
  sp.trk = sp.lvl + (static_cast<size_t>(nq) + 63) / 64 * 64;
  sp.lvl_lg = pl.lvl_lg;
  sp.lvl_slot = pl.lvl_slot;
  sp.lvl_w = pl.lvl_w;
  sp.lists = pl.lists;
  sp.lvl_sleep_ns = g_lvl_sleep_ns;
  sp.lvl_dbg = g_debug_counts ? reinterpret_cast<unsigned long long*>(flag_count + 8) : nullptr;  // cleared by the prep kernel
  sp.perf_mode = g_perf_mode;
  {
    // readers of a slice are kept within ~half an L2 share of each other (see the TMA producer)
    const int readers = pl.grid_x / (pl.pair ? 2 : 1);
    const size_t tile_bytes = static_cast<size_t>(ix->tile_rows) * dpad * sizeof(__half);
    const size_t corpus_bytes = static_cast<size_t>(pl.num_tiles) * tile_bytes;
    sp.progress = nullptr;
    if (g_sync_readers && readers > 1 && readers <= 32 && pl.slices * readers <= kProgWords && corpus_bytes > (96u << 20)) {
      sp.progress = flag_count + 64;
      const size_t share = (static_cast<size_t>(g_sync_mb) << 20) / static_cast<size_t>(pl.slices);
      sp.sync_window = static_cast<int>(std::min<size_t>(64, std::max<size_t>(2, share / tile_bytes)));
    }
  }
  // exact-path workspace (its counters are cleared by the prep kernel as well)
  // one list per query: however many queries of a batch end up uncertified (a corpus full of exact
  // duplicates), none can be left without its exact pass - also when the caller reads the results
  // asynchronously and nobody looks at the flag count.  192 KB of workspace per query, reserved once.
  const int nflag_max = nq;
  const size_t ex_bytes = static_cast<size_t>(nflag_max) * kExactListCap * (sizeof(double) + sizeof(unsigned)) +
                          static_cast<size_t>(nflag_max) * sizeof(int) + 256;
  LXG_CUDA(ix->ws_exact.reserve(ex_bytes));
  double* ex_score = reinterpret_cast<double*>(ix->ws_exact.p);
  unsigned* ex_row = reinterpret_cast<unsigned*>(ex_score + static_cast<size_t>(nflag_max) * kExactListCap);
  int* ex_count = reinterpret_cast<int*>(ex_row + static_cast<size_t>(nflag_max) * kExactListCap);

  int launches = 0;
  cudaEvent_t* ev = nullptr;
  if (ix->timing && !dbg_scores) {
    if (ix->ev_used + 5 > ix->ev_pool.size()) {
      for (int i = 0; i < 5; ++i) {
        cudaEvent_t e;
        LXG_CUDA(cudaEventCreate(&e));
        ix->ev_pool.push_back(e);
      }
    }
    ev = &ix->ev_pool[ix->ev_used];
    ix->ev_used += 5;
  }
  if (ev) LXG_CUDA(cudaEventRecord(ev[0], st));
  {
  NvtxRange nvtx_scan("lxg_search: prep + scan (TMA + tcgen05 + level warps)");
  prep_queries_kernel<<<(nq_pad + 7) / 8, 256, 0, st>>>(
      x, xn, xh, qscale, qnorm, nq, nq_pad, d, dpad, normalize, reinterpret_cast<uint32_t*>(flag_count),
      64 + kProgWords + static_cast<int>(lvl_words), reinterpret_cast<uint32_t*>(ex_count), nflag_max);
  LXG_CUDA(cudaGetLastError());
  ++launches;
  if (ev) LXG_CUDA(cudaEventRecord(ev[1], st));
  if (ix->a_smem_chunks == 8) {
    if (pl.pair) LXG_CUDA((launch_scan<128, true, 8>(ix, sp, pl.grid_x, st)));
    else LXG_CUDA((launch_scan<128, false, 8>(ix, sp, pl.grid_x, st)));
  } else if (ix->a_smem_chunks == 4) {
    if (pl.pair) LXG_CUDA((launch_scan<128, true, 4>(ix, sp, pl.grid_x, st)));
    else LXG_CUDA((launch_scan<128, false, 4>(ix, sp, pl.grid_x, st)));
  } else if (ix->tile_rows == 128) {
    if (pl.pair) LXG_CUDA((launch_scan<128, true>(ix, sp, pl.grid_x, st)));
    else LXG_CUDA((launch_scan<128, false>(ix, sp, pl.grid_x, st)));
  } else {
    if (pl.pair) LXG_CUDA((launch_scan<64, true>(ix, sp, pl.grid_x, st)));
    else LXG_CUDA((launch_scan<64, false>(ix, sp, pl.grid_x, st)));
  }
  ++launches;
  }
  if (ev) LXG_CUDA(cudaEventRecord(ev[2], st));
  if (g_debug_counts) {
    std::vector<int> hc(lists);
    std::vector<uint32_t> hl(sp.lvl_r > 0 ? nq : 0);
    LXG_CUDA(cudaStreamSynchronize(st));
    LXG_CUDA(cudaMemcpy(hc.data(), sp.cand_count, lists * sizeof(int), cudaMemcpyDeviceToHost));
    if (!hl.empty()) LXG_CUDA(cudaMemcpy(hl.data(), sp.lvl, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    long long tot = 0;
    int mx = 0, none = 0;
    for (int c : hc) {
      tot += c;
      mx = std::max(mx, c);
    }
    for (uint32_t l : hl) none += l == 0u;
    unsigned long long dbg[3] = {0, 0, 0};
    LXG_CUDA(cudaMemcpy(dbg, flag_count + 8, sizeof(dbg), cudaMemcpyDeviceToHost));
    const double ctas = static_cast<double>(pl.grid_x) * pl.slices * kLvlWarps;
    std::fprintf(stderr, "[lxg] level warps: %.1f rounds per warp, %.0f clocks per round selecting, %.0f clocks alive\n",
                 dbg[0] / ctas, dbg[0] ? static_cast<double>(dbg[1]) / dbg[0] : 0.0, dbg[2] / ctas);
    std::fprintf(stderr, "[lxg] pass 1: nq=%d kp=%d lists=%d depth=%d classes=%d  candidates/query=%.1f  longest list=%d  queries without a level=%d\n",
                 nq, pl.kp, pl.lists, pl.lvl_r, 1 << pl.lvl_lg, static_cast<double>(tot) / nq, mx, none);
  }
  ix->stats.slices = pl.slices;
  ix->stats.query_blocks = pl.qblocks;
  ix->stats.kp = pl.kp;
  ix->stats.tile_rows = ix->tile_rows;
  ix->stats.uncertified = -1;
  if (dbg_scores || g_perf_mode) {  // perf mode: pass 1 only, results are not produced
    if (ev) {
      LXG_CUDA(cudaEventRecord(ev[3], st));
      LXG_CUDA(cudaEventRecord(ev[4], st));
    }
    ix->stats.kernel_launches = launches;
    return LXG_OK;
  }

  NvtxRange nvtx_merge("lxg_search: merge + exact re-score + certificate");
  MergeParams mp{};
  mp.cand = sp.cand;
  mp.cand_count = sp.cand_count;
  mp.slice_thr = sp.slice_thr;
  mp.lvl = sp.lvl_r > 0 ? sp.lvl : nullptr;
  mp.xn = xn;
  mp.qscale = qscale;
  mp.qnorm = qnorm;
  mp.out_d = D;
  mp.out_i = I;
  mp.out_d64 = D64;
  mp.flag_count = flag_count;
  mp.flag_list = reinterpret_cast<int*>(sm + o_flist);
  mp.flag_theta = reinterpret_cast<double*>(sm + o_theta);
  mp.nq = nq;
  mp.k = k;
  mp.kp = pl.kp;
  mp.cap = pl.cap;
  mp.lists = pl.lists;
  mp.lvl_slots = 1;
  mp.max_items = pl.max_items;
  mp.sort_n = sort_n;
  mp.select_only = 0;
  mp.sel_row_g = reinterpret_cast<unsigned*>(sm + o_selrow);
  mp.sel_score_g = reinterpret_cast<double*>(sm + o_selscore);
  mp.sel_n_g = reinterpret_cast<int*>(sm + o_seln);
  mp.sel_amin_g = reinterpret_cast<float*>(sm + o_selamin);
  const size_t msmem = static_cast<size_t>(pl.max_items) * 8 + static_cast<size_t>(sort_n) * 12 +
                       static_cast<size_t>(d) * 4 + 64;
  // one CTA per query, sized by the batch (see rescore.cuh)
  if (split_merge) {
    // select (one CTA per query) -> exact re-score on every SM -> rank + certify (rescore.cuh)
    mp.select_only = 1;
    LXG_CUDA((launch_merge<1024>(nq, msmem, mp, ix->cv, st)));
    rescore_rows_kernel<<<dim3(sort_n / 32, nq), 256, d * sizeof(float), st>>>(mp, ix->cv);
    LXG_CUDA(cudaGetLastError());
    rank_rows_kernel<<<dim3((sort_n + 31) / 32, nq), 256, static_cast<size_t>(sort_n) * 12, st>>>(mp, ix->cv);
    LXG_CUDA(cudaGetLastError());
    launches += 2;
  } else if (nq <= 64) LXG_CUDA((launch_merge<1024>(nq, msmem, mp, ix->cv, st)));
  else if (nq <= 512) LXG_CUDA((launch_merge<256>(nq, msmem, mp, ix->cv, st)));
  else LXG_CUDA((launch_merge<128>(nq, msmem, mp, ix->cv, st)));
  LXG_CUDA(cudaGetLastError());
  ++launches;
  if (ev) LXG_CUDA(cudaEventRecord(ev[3], st));

  // exact path for uncertified queries (normally zero of them: both kernels exit at once)
  ExactParams ep{};
  ep.xn = xn;
  ep.flag_count = flag_count;
  ep.flag_list = mp.flag_list;
  ep.flag_theta = mp.flag_theta;
  ep.list_score = ex_score;
  ep.list_row = ex_row;
  ep.list_count = ex_count;
  ep.out_d = D;
  ep.out_i = I;
  ep.out_d64 = D64;
  ep.overflow = overflow;
  ep.nq = nq;
  ep.k = k;
  ep.nflag_max = nflag_max;
  exact_collect_kernel<<<2 * ix->sms, 256, d * sizeof(float), st>>>(ep, ix->cv);
  LXG_CUDA(cudaGetLastError());
  exact_finalize_kernel<<<nflag_max, 256, 0, st>>>(ep, ix->cv);
  LXG_CUDA(cudaGetLastError());
  launches += 2;
  if (ev) LXG_CUDA(cudaEventRecord(ev[4], st));
  ix->stats.kernel_launches = launches;

  if (read_flags) {
    int h[2] = {0, 0};
    LXG_CUDA(cudaMemcpyAsync(h, flag_count, sizeof(h), cudaMemcpyDeviceToHost, st));
    LXG_CUDA(cudaStreamSynchronize(st));
    ix->stats.uncertified = h[0];
    if (h[0] > nflag_max) return set_error(LXG_ECUDA, "internal: more flagged queries than queries");
    if (h[1])
      return set_error(LXG_ETIES, "more than 16384 corpus rows tie with the k-th best score of a query");
  }
  return LXG_OK;
}

}  // namespace

extern "C" {

int lxg_search_ex(lxg_index* ix, const float* x, int32_t nq, int32_t k, int normalize, float* D_out,
                  int64_t* I_out, double* D64_out, void* stream) {
  if (!ix || !D_out || !I_out) return set_error(LXG_EINVAL, "NULL argument");
  if (nq < 0 || k <= 0) return set_error(LXG_EINVAL, "nq must be >= 0 and k >= 1");
  if (k > 2048) return set_error(LXG_EUNSUPPORTED, "k > 2048");
  if (nq == 0) return LXG_OK;
  if (!x) return set_error(LXG_EINVAL, "x is NULL");
  DeviceGuard guard(ix->device);
  NvtxRange nvtx("lxg_search_ex");
  std::lock_guard<std::mutex> lock(ix->mu);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int d = ix->cv.d;
  // Searches on one handle share its workspaces (candidate lists, counters, staging): order this one
  // behind the previous one ON THE DEVICE, whatever stream / host thread that was issued from (with
  // device outputs a call returns before its kernels have run).
  if (ix->have_done) LXG_CUDA(cudaStreamWaitEvent(st, ix->done_ev, 0));
  struct RecordDone {
    lxg_index* ix;
    cudaStream_t st;
    ~RecordDone() {
      if (cudaEventRecord(ix->done_ev, st) == cudaSuccess) ix->have_done = true;
    }
  } record_done{ix, st};
  const bool x_dev = is_device_ptr(x);
  const bool out_dev = is_device_ptr(D_out);
  if (out_dev != is_device_ptr(I_out))
    return set_error(LXG_EINVAL, "D_out and I_out must both be host or both be device memory");
  if (D64_out && !is_device_ptr(D64_out)) return set_error(LXG_EINVAL, "D64_out must be device memory");
  ix->stats = lxg_search_stats{};

  const size_t out_elems = static_cast<size_t>(nq) * k;
  if (ix->cv.n == 0) {  // empty index: FAISS returns all -1 / -FLT_MAX
    std::vector<float> hd(out_elems, -FLT_MAX);
    std::vector<long long> hi(out_elems, -1);
    if (out_dev) {
      LXG_CUDA(cudaMemcpyAsync(D_out, hd.data(), out_elems * 4, cudaMemcpyHostToDevice, st));
      LXG_CUDA(cudaMemcpyAsync(I_out, hi.data(), out_elems * 8, cudaMemcpyHostToDevice, st));
      LXG_CUDA(cudaStreamSynchronize(st));
    } else {
      std::memcpy(D_out, hd.data(), out_elems * 4);
      std::memcpy(I_out, hi.data(), out_elems * 8);
    }
    return LXG_OK;
  }

  // device staging for host-side arguments
  const size_t x_bytes = static_cast<size_t>(nq) * d * sizeof(float);
  const size_t stage_bytes = (x_dev ? 0 : x_bytes) + (out_dev ? 0 : out_elems * 12) + 256;
  LXG_CUDA(ix->ws_out.reserve(stage_bytes));
  uint8_t* stg = reinterpret_cast<uint8_t*>(ix->ws_out.p);
  const float* xd = x;
  void* const x_alias = (!x_dev && g_zero_copy) ? mapped_host_alias(x) : nullptr;
  if (x_alias != nullptr) {
    xd = reinterpret_cast<const float*>(x_alias);  // the preparation kernel reads the pinned queries over PCIe
  } else if (!x_dev) {
    const void* src = x;
    if (!is_pinned_host_ptr(x)) {  // pageable caller memory: bounce through the pinned stage
      LXG_CUDA(ix->h_stage.reserve(std::max(x_bytes, out_elems * 12)));
      std::memcpy(ix->h_stage.p, x, x_bytes);
      src = ix->h_stage.p;
    }
    LXG_CUDA(cudaMemcpyAsync(stg, src, x_bytes, cudaMemcpyHostToDevice, st));
    xd = reinterpret_cast<const float*>(stg);
    stg += (x_bytes + 255) / 256 * 256;
  }
  float* Dd = D_out;
  long long* Id = reinterpret_cast<long long*>(I_out);
  void* const d_alias = (!out_dev && g_zero_copy) ? mapped_host_alias(D_out) : nullptr;
  void* const i_alias = (!out_dev && g_zero_copy) ? mapped_host_alias(I_out) : nullptr;
  const bool out_zero_copy = d_alias != nullptr && i_alias != nullptr;
  if (out_zero_copy) {  // the merge / exact kernels write the pinned result buffers directly
    Dd = reinterpret_cast<float*>(d_alias);
    Id = reinterpret_cast<long long*>(i_alias);
  } else if (!out_dev) {
    Id = reinterpret_cast<long long*>(stg);
    Dd = reinterpret_cast<float*>(stg + out_elems * 8);
  }
  // at most 148 query blocks per launch
  const int max_q = 148 * kQueryBlock;
  int total_launches = 0, total_flag = 0;
  for (int q0 = 0; q0 < nq; q0 += max_q) {
    const int nb = std::min(max_q, nq - q0);
    int rc = search_device(ix, xd + static_cast<size_t>(q0) * d, nb, k, normalize,
                           Dd + static_cast<size_t>(q0) * k, Id + static_cast<size_t>(q0) * k,
                           D64_out ? D64_out + static_cast<size_t>(q0) * k : nullptr, nullptr, nullptr, st,
                           /*read_flags=*/!out_dev || nq > max_q);
    if (rc != LXG_OK) return rc;
    total_launches += ix->stats.kernel_launches;
    if (ix->stats.uncertified > 0) total_flag += ix->stats.uncertified;
  }
  ix->stats.kernel_launches = total_launches;
  if (!out_dev) {
    ix->stats.uncertified = total_flag;
    if (out_zero_copy) {
      // nothing to copy; every search_device call above ended with the flag read-back's synchronise
    } else if (is_pinned_host_ptr(D_out) && is_pinned_host_ptr(I_out)) {
      LXG_CUDA(cudaMemcpyAsync(I_out, Id, out_elems * 8, cudaMemcpyDeviceToHost, st));
      LXG_CUDA(cudaMemcpyAsync(D_out, Dd, out_elems * 4, cudaMemcpyDeviceToHost, st));
      LXG_CUDA(cudaStreamSynchronize(st));
    } else {
      LXG_CUDA(ix->h_stage.reserve(out_elems * 12));
      LXG_CUDA(cudaMemcpyAsync(ix->h_stage.p, Id, out_elems * 12, cudaMemcpyDeviceToHost, st));
      LXG_CUDA(cudaStreamSynchronize(st));
      std::memcpy(I_out, ix->h_stage.p, out_elems * 8);
      std::memcpy(D_out, reinterpret_cast<uint8_t*>(ix->h_stage.p) + out_elems * 8, out_elems * 4);
    }
  }
  return LXG_OK;
}

int lxg_search(lxg_index* ix, const float* x, int32_t nq, int32_t k, int normalize, float* D_out,
               int64_t* I_out, void* stream) {
  return lxg_search_ex(ix, x, nq, k, normalize, D_out, I_out, nullptr, stream);
}

int lxg_debug_scores(lxg_index* ix, const float* x_dev, int32_t nq, int normalize, float* scores_dev,
                     float* qscale_dev, float* scan_scale_host, void* stream) {
  if (!ix || !x_dev || !scores_dev || !qscale_dev) return set_error(LXG_EINVAL, "NULL argument");
  if (nq <= 0 || nq > 148 * kQueryBlock) return set_error(LXG_EINVAL, "nq out of range");
  DeviceGuard guard(ix->device);
  std::lock_guard<std::mutex> lock(ix->mu);
  if (scan_scale_host) *scan_scale_host = ix->cv.scan_scale;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (ix->have_done) LXG_CUDA(cudaStreamWaitEvent(st, ix->done_ev, 0));
  const int rc = search_device(ix, x_dev, nq, 1, normalize, nullptr, nullptr, nullptr, scores_dev, qscale_dev, st, false);
  if (cudaEventRecord(ix->done_ev, st) == cudaSuccess) ix->have_done = true;
  return rc;
}

int lxg_index_sync(lxg_index* ix, int32_t* uncertified) {
  if (!ix) return set_error(LXG_EINVAL, "NULL argument");
  DeviceGuard guard(ix->device);
  std::lock_guard<std::mutex> lock(ix->mu);
  if (uncertified) *uncertified = 0;
  if (!ix->have_done || !ix->last_flags) return LXG_OK;
  LXG_CUDA(cudaEventSynchronize(ix->done_ev));
  int h[2] = {0, 0};
  LXG_CUDA(cudaMemcpy(h, ix->last_flags, sizeof(h), cudaMemcpyDeviceToHost));
  ix->stats.uncertified = h[0];
  if (uncertified) *uncertified = h[0];
  if (h[1]) return set_error(LXG_ETIES, "more than 16384 corpus rows tie with the k-th best score of a query");
  return LXG_OK;
}

int lxg_normalize_l2(float* x, int32_t nq, int32_t d, void* stream) {
  if (!x || nq < 0 || d <= 0) return set_error(LXG_EINVAL, "bad argument");
  if (nq == 0) return LXG_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t bytes = static_cast<size_t>(nq) * d * sizeof(float);
  const int dev = device_of_ptr(x);
  if (dev >= 0) {
    DeviceGuard guard(dev);
    normalize_l2_kernel<<<(nq + 7) / 8, 256, 0, st>>>(x, nq, d);
    LXG_CUDA(cudaGetLastError());
    return LXG_OK;
  }
  // Host array (the reference's own call: one [1, d] query per request, engine.py:242).  Page-locked
  // memory is normalised in place through its device alias; pageable memory goes through a device
  // scratch buffer that is kept per device (no allocation on the request path).
  if (void* alias = g_zero_copy ? mapped_host_alias(x) : nullptr) {
    normalize_l2_kernel<<<(nq + 7) / 8, 256, 0, st>>>(reinterpret_cast<float*>(alias), nq, d);
    LXG_CUDA(cudaGetLastError());
    LXG_CUDA(cudaStreamSynchronize(st));
    return LXG_OK;
  }
  int cur = 0;
  LXG_CUDA(cudaGetDevice(&cur));
  if (cur < 0 || cur >= kMaxDevices) return set_error(LXG_EINVAL, "device index out of range");
  static std::mutex mu;
  static DevBuf scratch[kMaxDevices];
  std::lock_guard<std::mutex> lock(mu);
  LXG_CUDA(scratch[cur].reserve(bytes));
  float* tmp = reinterpret_cast<float*>(scratch[cur].p);
  LXG_CUDA(cudaMemcpyAsync(tmp, x, bytes, cudaMemcpyHostToDevice, st));
  normalize_l2_kernel<<<(nq + 7) / 8, 256, 0, st>>>(tmp, nq, d);
  LXG_CUDA(cudaGetLastError());
  LXG_CUDA(cudaMemcpyAsync(x, tmp, bytes, cudaMemcpyDeviceToHost, st));
  LXG_CUDA(cudaStreamSynchronize(st));
  return LXG_OK;
}

static int merge_shards(const double* dg, const long long* ig, long long shard_stride, int nq, int k, int shards,
                        float* D_out, int64_t* I_out, cudaStream_t st) {
  const int dev = device_of_ptr(D_out);
  if (dev < 0 || device_of_ptr(dg) != dev || device_of_ptr(ig) != dev || device_of_ptr(I_out) != dev)
    return set_error(LXG_EINVAL, "the gathered candidates and the outputs must be memory of one device");
  DeviceGuard guard(dev);
  const int total = shards * k;
  const int threads = std::min(256, (total + 31) / 32 * 32);
  const size_t smem = static_cast<size_t>(total) * 16;
  if (smem <= 160u * 1024u) {
    LXG_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(&merge_shards_kernel<true>), smem));
    merge_shards_kernel<true><<<nq, threads, smem, st>>>(dg, ig, shard_stride, nq, k, shards, D_out,
                                                          reinterpret_cast<long long*>(I_out));
  } else {
    merge_shards_kernel<false><<<nq, threads, 0, st>>>(dg, ig, shard_stride, nq, k, shards, D_out,
                                                        reinterpret_cast<long long*>(I_out));
  }
  LXG_CUDA(cudaGetLastError());
  return LXG_OK;
}

int lxg_merge_topk(const double* Dg, const int64_t* Ig, int32_t nq, int32_t k, int32_t shards, float* D_out,
                   int64_t* I_out, void* stream) {
  if (!Dg || !Ig || !D_out || !I_out) return set_error(LXG_EINVAL, "NULL argument");
  if (nq < 0 || k <= 0 || shards <= 0) return set_error(LXG_EINVAL, "bad nq / k / shards");
  if (nq == 0) return LXG_OK;
  return merge_shards(Dg, reinterpret_cast<const long long*>(Ig), static_cast<long long>(nq) * k, nq, k, shards, D_out,
                      I_out, reinterpret_cast<cudaStream_t>(stream));
}

int lxg_merge_topk_packed(const int64_t* gathered, int32_t nq, int32_t k, int32_t shards, float* D_out,
                          int64_t* I_out, void* stream) {
  if (!gathered || !D_out || !I_out) return set_error(LXG_EINVAL, "NULL argument");
  if (nq < 0 || k <= 0 || shards <= 0) return set_error(LXG_EINVAL, "bad nq / k / shards");
  if (nq == 0) return LXG_OK;
  const long long plane = static_cast<long long>(nq) * k;
  return merge_shards(reinterpret_cast<const double*>(gathered), reinterpret_cast<const long long*>(gathered) + plane,
                      2 * plane, nq, k, shards, D_out, I_out, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
