/* CPU ORACLE (C) - test infrastructure, not product code.
 *
 * Plain-C restatement of the FAISS routines behind the reference's k-NN call sites
 *   faiss.normalize_L2(q)            src/lean_explore/search/engine.py:242
 *   index.search(q, faiss_k)         src/lean_explore/search/engine.py:250
 * for a flat inner-product index (the oracle the north star names).  FAISS itself is a
 * third-party wheel (faiss-cpu>=1.7, reference pyproject.toml:36) absent from /root/reference
 * and not installable here, so this follows FAISS' published algorithm
 * (faiss/utils/distances.cpp: fvec_renorm_L2, exhaustive_inner_product_seq,
 * exhaustive_inner_product_blas; faiss/utils/Heap.h: CMin heap, heap_replace_top,
 * heap_reorder; faiss/impl/ResultHandler.h: HeapBlockResultHandler::add_results).
 *
 * Used by tests/ (cross-check of oracle/faiss_flat.py on tie-free data, KAT) and by bench.py's
 * cpu_baseline / --impl reference arm (numpy sgemm for the block products, these heaps for
 * the selection - the structure of FAISS' nq >= 20 path).  Never linked into liblxg.so.
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- faiss/utils/distances_simd.cpp: fvec_inner_product / fvec_norm_L2sqr (fp32 accumulate) */
static float fvec_inner_product(const float* x, const float* y, size_t d) {
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  size_t i = 0;
  for (; i + 8 <= d; i += 8)
    for (int j = 0; j < 8; ++j) acc[j] += x[i + j] * y[i + j];
  float r = ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
  for (; i < d; ++i) r += x[i] * y[i];
  return r;
}

/* faiss fvec_renorm_L2: nr = ||x||^2 (fp32); if nr > 0 { inv_nr = 1.0 / sqrtf(nr); x *= inv_nr } */
void lxo_renorm_l2(size_t d, size_t nx, float* x) {
#pragma omp parallel for schedule(static) if (nx > 10000)
  for (int64_t i = 0; i < (int64_t)nx; ++i) {
    float* xi = x + (size_t)i * d;
    float nr = fvec_inner_product(xi, xi, d);
    if (nr > 0) {
      const float inv_nr = 1.0 / sqrtf(nr);
      for (size_t j = 0; j < d; ++j) xi[j] *= inv_nr;
    }
  }
}

/* ---- faiss/utils/Heap.h, CMin<float,int64_t>: min-heap keeping the k largest values */
static inline int cmin_cmp2(float a1, float a2, int64_t b1, int64_t b2) {
  return (a1 < a2) || ((a1 == a2) && (b1 < b2));
}

static void heap_heapify(size_t k, float* val, int64_t* ids) {
  for (size_t i = 0; i < k; ++i) {
    val[i] = -FLT_MAX; /* CMin::neutral() == numeric_limits<float>::lowest() */
    ids[i] = -1;
  }
}

static void heap_replace_top(size_t k, float* bh_val, int64_t* bh_ids, float val, int64_t id) {
  bh_val--; /* 1-based indexing */
  bh_ids--;
  size_t i = 1, i1, i2;
  for (;;) {
    i1 = i << 1;
    i2 = i1 + 1;
    if (i1 > k) break;
    if (i2 == k + 1 || cmin_cmp2(bh_val[i1], bh_val[i2], bh_ids[i1], bh_ids[i2])) {
      if (cmin_cmp2(val, bh_val[i1], id, bh_ids[i1])) break;
      bh_val[i] = bh_val[i1];
      bh_ids[i] = bh_ids[i1];
      i = i1;
    } else {
      if (cmin_cmp2(val, bh_val[i2], id, bh_ids[i2])) break;
      bh_val[i] = bh_val[i2];
      bh_ids[i] = bh_ids[i2];
      i = i2;
    }
  }
  bh_val[i] = val;
  bh_ids[i] = id;
}

static void heap_pop(size_t k, float* bh_val, int64_t* bh_ids) {
  bh_val--;
  bh_ids--;
  float val = bh_val[k];
  int64_t id = bh_ids[k];
  size_t i = 1, i1, i2;
  for (;;) {
    i1 = i << 1;
    i2 = i1 + 1;
    if (i1 > k) break;
    if (i2 == k + 1 || cmin_cmp2(bh_val[i1], bh_val[i2], bh_ids[i1], bh_ids[i2])) {
      if (cmin_cmp2(val, bh_val[i1], id, bh_ids[i1])) break;
      bh_val[i] = bh_val[i1];
      bh_ids[i] = bh_ids[i1];
      i = i1;
    } else {
      if (cmin_cmp2(val, bh_val[i2], id, bh_ids[i2])) break;
      bh_val[i] = bh_val[i2];
      bh_ids[i] = bh_ids[i2];
      i = i2;
    }
  }
  bh_val[i] = bh_val[k];
  bh_ids[i] = bh_ids[k];
}

/* heap_reorder: best first, unfilled slots (-FLT_MAX, -1) at the end */
static void heap_reorder(size_t k, float* bh_val, int64_t* bh_ids) {
  size_t i, ii;
  for (i = 0, ii = 0; i < k; i++) {
    float val = bh_val[0];
    int64_t id = bh_ids[0];
    heap_pop(k - i, bh_val, bh_ids);
    bh_val[k - ii - 1] = val;
    bh_ids[k - ii - 1] = id;
    if (id != -1) ii++;
  }
  size_t nel = ii;
  memmove(bh_val, bh_val + k - ii, ii * sizeof(*bh_val));
  memmove(bh_ids, bh_ids + k - ii, ii * sizeof(*bh_ids));
  for (; ii < k; ii++) {
    bh_val[ii] = -FLT_MAX;
    bh_ids[ii] = -1;
  }
  (void)nel;
}

void lxo_heap_init(size_t nq, size_t k, float* D, int64_t* I) {
  for (size_t q = 0; q < nq; ++q) heap_heapify(k, D + q * k, I + q * k);
}

/* HeapBlockResultHandler::add_results: scores is the [nq, nb] block x . y[j0:j0+nb]^T */
void lxo_heap_add_block(size_t nq, size_t k, float* D, int64_t* I, const float* scores, size_t nb,
                        int64_t j0) {
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < (int64_t)nq; ++q) {
    float* simi = D + (size_t)q * k;
    int64_t* idxi = I + (size_t)q * k;
    const float* line = scores + (size_t)q * nb;
    float thresh = simi[0];
    for (size_t j = 0; j < nb; ++j) {
      const float ip = line[j];
      if (ip > thresh) { /* C::cmp(thresh, ip) for CMin */
        heap_replace_top(k, simi, idxi, ip, j0 + (int64_t)j);
        thresh = simi[0];
      }
    }
  }
}

void lxo_heap_finish(size_t nq, size_t k, float* D, int64_t* I) {
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < (int64_t)nq; ++q) heap_reorder(k, D + (size_t)q * k, I + (size_t)q * k);
}

/* exhaustive_inner_product_seq: the nq < 20 path (per-pair dot products, no BLAS) */
void lxo_knn_inner_product_seq(const float* x, const float* y, size_t d, size_t nx, size_t ny, size_t k,
                               float* D, int64_t* I) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t i = 0; i < (int64_t)nx; ++i) {
    const float* xi = x + (size_t)i * d;
    float* simi = D + (size_t)i * k;
    int64_t* idxi = I + (size_t)i * k;
    heap_heapify(k, simi, idxi);
    float thresh = simi[0];
    for (size_t j = 0; j < ny; ++j) {
      const float ip = fvec_inner_product(xi, y + j * d, d);
      if (ip > thresh) {
        heap_replace_top(k, simi, idxi, ip, (int64_t)j);
        thresh = simi[0];
      }
    }
    heap_reorder(k, simi, idxi);
  }
}

/* ---- exact-arithmetic ranking at full size (bench.py's ids_equal_exact, tests at the BASELINE
 * sizes).  Not a FAISS routine: it is the DEFINITION the FAISS fp32 result approximates, the same
 * one oracle/faiss_flat.py flat_ip_search_f64 evaluates with numpy, written so that a 16M x 768
 * fp16 corpus can be streamed through it block by block in seconds:
 *   x     [nq, d] fp32 queries as handed to index.search (engine.py:250; already normalised),
 *   rows  [m, d]  corpus rows, IEEE fp16 bit patterns (rows_f16 != 0) or fp32, global ids row0..,
 *   D/I   [nq, k] running result, best first, initialised by lxo_f64_topk_init.
 * Products of an fp32 by an fp16/fp32 value are exact in fp64; the fp64 sum over d carries a
 * relative error ~1e-13, nine orders of magnitude below the fp32 resolution of the scores.
 * Order: score descending, ties by ascending row id (the tie rule of faiss_flat.py). */
void lxo_f64_topk_init(size_t nq, size_t k, double* D, int64_t* I) {
  for (size_t i = 0; i < nq * k; ++i) {
    D[i] = -INFINITY;
    I[i] = -1;
  }
}

static inline int f64_better(double sa, int64_t ia, double sb, int64_t ib) {
  if (ib < 0) return 1;
  return sa > sb || (sa == sb && ia < ib);
}

static void f64_insert(size_t k, double* D, int64_t* I, double s, int64_t id) {
  if (!f64_better(s, id, D[k - 1], I[k - 1])) return;
  size_t j = k - 1;
  while (j > 0 && f64_better(s, id, D[j - 1], I[j - 1])) {
    D[j] = D[j - 1];
    I[j] = I[j - 1];
    --j;
  }
  D[j] = s;
  I[j] = id;
}

void lxo_f64_topk_add_rows(size_t nq, size_t d, size_t k, const float* x, const void* rows, int rows_f16,
                           size_t m, int64_t row0, double* D, int64_t* I) {
  double* xd = (double*)malloc(nq * d * sizeof(double));
  for (size_t i = 0; i < nq * d; ++i) xd[i] = (double)x[i];
#pragma omp parallel
  {
    enum { R = 4 }; /* rows per pass over a query: each query element is loaded once per R products */
    double* r = (double*)malloc(R * d * sizeof(double));
    double* Dl = (double*)malloc(nq * k * sizeof(double));
    int64_t* Il = (int64_t*)malloc(nq * k * sizeof(int64_t));
    lxo_f64_topk_init(nq, k, Dl, Il);
    const int64_t groups = ((int64_t)m + R - 1) / R;
#pragma omp for schedule(static)
    for (int64_t g = 0; g < groups; ++g) {
      const int64_t j0 = g * R;
      const int nr = (int)((int64_t)m - j0 < R ? (int64_t)m - j0 : R);
      for (int u = 0; u < R; ++u) {
        double* ru = r + (size_t)u * d;
        if (u >= nr) {
          for (size_t i = 0; i < d; ++i) ru[i] = 0.0;
        } else if (rows_f16) {
          const _Float16* src = (const _Float16*)rows + (size_t)(j0 + u) * d;
          for (size_t i = 0; i < d; ++i) ru[i] = (double)src[i];
        } else {
          const float* src = (const float*)rows + (size_t)(j0 + u) * d;
          for (size_t i = 0; i < d; ++i) ru[i] = (double)src[i];
        }
      }
      const double *r0 = r, *r1 = r + d, *r2 = r + 2 * d, *r3 = r + 3 * d;
      for (size_t q = 0; q < nq; ++q) {
        const double* xq = xd + q * d;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma omp simd reduction(+ : a0, a1, a2, a3)
        for (size_t i = 0; i < d; ++i) {
          const double xv = xq[i];
          a0 += r0[i] * xv;
          a1 += r1[i] * xv;
          a2 += r2[i] * xv;
          a3 += r3[i] * xv;
        }
        const double acc[R] = {a0, a1, a2, a3};
        for (int u = 0; u < nr; ++u)
          if (acc[u] == acc[u]) f64_insert(k, Dl + q * k, Il + q * k, acc[u], row0 + j0 + u); /* NaN never enters */
      }
    }
#pragma omp critical
    {
      for (size_t q = 0; q < nq; ++q)
        for (size_t t = 0; t < k && Il[q * k + t] >= 0; ++t)
          f64_insert(k, D + q * k, I + q * k, Dl[q * k + t], Il[q * k + t]);
    }
    free(r);
    free(Dl);
    free(Il);
  }
  free(xd);
}

int lxo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
