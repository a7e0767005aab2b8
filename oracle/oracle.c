/*
 * oracle.c — CPU restatement of the cuVS scan + top-k hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this file's library.  Nothing under cuvs_b200/ links, imports or executes it.
 *
 * Each function restates one reference routine (paths relative to /root/reference):
 *   oracle_knn                 cpp/tests/neighbors/naive_knn.cuh:21-138 (unexpanded metrics) and
 *                              cpp/src/distance/detail/distance_ops/l2_exp.cuh:68-128 +
 *                              cpp/src/neighbors/detail/knn_brute_force.cuh:353-539 (expanded L2,
 *                              clamp, post-sqrt), cosine :194-225
 *   oracle_select_k            cpp/include/cuvs/selection/select_k.hpp:70-198 (semantics; the
 *                              arithmetic lives in RAFT which is not vendored — tie order is
 *                              UNPINNED upstream, pinned here to "smaller position first")
 *   oracle_fp8_*               cpp/src/neighbors/ivf_pq/ivf_pq_fp_8bit.cuh:31-100
 *   oracle_ivf_coarse          cpp/src/neighbors/ivf_pq/ivf_pq_search.cuh:60-168 (NOTE[qc_distances])
 *                              cpp/src/neighbors/ivf_flat/ivf_flat_search.cuh:105-187
 *   oracle_ivf_flat_search     ivf_flat_search.cuh:41-309 + ivf_flat/detail/jit_lto_kernels/
 *                              interleaved_scan_impl.cuh:127-186, metric_impl.cuh
 *   oracle_ivf_pq_search       ivf_pq_search.cuh:881-1048, detail/jit_lto_kernels/
 *                              create_lut_impl.cuh:16-79, compute_score_impl.cuh:53-79,
 *                              compute_distances_impl.cuh:15-105, ivf_common.cuh:175-252
 *   oracle_cagra_search        cpp/src/neighbors/detail/cagra/jit_lto_kernels/
 *                              search_single_cta_jit.cuh:55-451, device_common_jit.cuh:36-179,
 *                              hashmap.hpp:37-73, search_single_cta_device_helpers.cuh:97-137,
 *                              search_plan.cuh:199-372
 *   oracle_kmeans_assign       cpp/src/cluster/detail/minClusterDistanceCompute.cu:18-165,
 *                              cpp/src/distance/unfused_distance_nn.cuh:41-83
 *
 * Arithmetic contract (what "bit-exact" means in tests/): every dot product / norm / squared
 * difference is accumulated in fp32 with fmaf in ascending component order from 0.  The GPU
 * re-scoring kernels use the same order, so distances — and therefore indices, with ties broken
 * towards the smaller id — can be compared exactly.  Accumulation order inside the reference's own
 * kernels (cuBLAS / tiled FFMA) is not observable from its tests (they accept id-or-distance within
 * eps, cpp/tests/neighbors/knn_utils.cuh:19-70), so this pins one valid order.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* cuvsDistanceType values used on the path (include/cuvs/distance/distance.h) */
enum { M_L2Expanded = 0, M_L2SqrtExpanded = 1, M_Cosine = 2, M_L2Unexpanded = 4,
       M_L2SqrtUnexpanded = 5, M_InnerProduct = 6 };

static inline float dotf(const float* a, const float* b, int d)
{
  float acc = 0.0f;
  for (int k = 0; k < d; ++k) acc = fmaf(a[k], b[k], acc);
  return acc;
}
static inline float sqdiff(const float* a, const float* b, int d)
{
  float acc = 0.0f;
  for (int k = 0; k < d; ++k) { float t = a[k] - b[k]; acc = fmaf(t, t, acc); }
  return acc;
}

void oracle_row_norms(const float* x, int64_t n, int d, float* out)
{
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) out[i] = dotf(x + i * d, x + i * d, d);
}

/* l2_exp.cuh:100-118: val = xn + yn - 2 acc ; negative -> 0 ; self-neighbour clamp */
static inline float l2_expanded(float qn, float xn, float dot)
{
  float val = fmaf(-2.0f, dot, qn + xn);
  if (!(val > 0.0f)) val = 0.0f;
  if (val * val < 1e-6f && qn == xn) val = 0.0f;
  return val;
}

/* distance between one query and one row, in the metric's "search" form (smaller is closer
 * except InnerProduct).  Sqrt variants are selected on the squared value and rooted afterwards. */
static inline float pair_distance(const float* q, const float* x, int d, int metric, float qn, float xn)
{
  switch (metric) {
    case M_L2Expanded:
    case M_L2SqrtExpanded: return l2_expanded(qn, xn, dotf(q, x, d));
    case M_L2Unexpanded:
    case M_L2SqrtUnexpanded: return sqdiff(q, x, d);
    case M_InnerProduct: return dotf(q, x, d);
    case M_Cosine: return 1.0f - dotf(q, x, d) / (sqrtf(qn) * sqrtf(xn));
    default: return NAN;
  }
}

typedef struct { float v; int64_t i; } pair_t;

/* strict weak order: better first; ties -> smaller id */
static inline int better(float va, int64_t ia, float vb, int64_t ib, int select_min)
{
  if (va != vb) return select_min ? (va < vb) : (va > vb);
  return ia < ib;
}

/* insert into a sorted (best-first) list of capacity k; *cnt is current size */
static inline void topk_push(pair_t* lst, int* cnt, int k, float v, int64_t i, int select_min)
{
  int n = *cnt;
  if (n == k) {
    if (!better(v, i, lst[k - 1].v, lst[k - 1].i, select_min)) return;
    n = k - 1;
  }
  int p = n;
  while (p > 0 && better(v, i, lst[p - 1].v, lst[p - 1].i, select_min)) { lst[p] = lst[p - 1]; --p; }
  lst[p].v = v; lst[p].i = i;
  *cnt = n + 1;
}

static inline void topk_flush(const pair_t* lst, int cnt, int k, int select_min, int64_t* oi, float* ov,
                              int64_t pad_idx)
{
  for (int j = 0; j < k; ++j) {
    if (j < cnt) { oi[j] = lst[j].i; ov[j] = lst[j].v; }
    else { oi[j] = pad_idx; ov[j] = select_min ? 3.402823466e+38f : -3.402823466e+38f; }
  }
}

/* Exact kNN.  out_idx[nq,k] int64, out_dist[nq,k].  n < k pads with id -1 (knn_brute_force.cuh:159-167). */
void oracle_knn(const float* ds, int64_t n, const float* qs, int64_t nq, int d, int metric, int k,
                int64_t* out_idx, float* out_dist)
{
  float* xn = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  oracle_row_norms(ds, n, d, xn);
  const int select_min = metric != M_InnerProduct;
#pragma omp parallel
  {
    pair_t* lst = (pair_t*)malloc(sizeof(pair_t) * (size_t)k);
#pragma omp for schedule(dynamic, 4)
    for (int64_t qi = 0; qi < nq; ++qi) {
      const float* q = qs + qi * d;
      const float qn = dotf(q, q, d);
      int cnt = 0;
      for (int64_t j = 0; j < n; ++j) {
        float v = pair_distance(q, ds + j * d, d, metric, qn, xn[j]);
        topk_push(lst, &cnt, k, v, j, select_min);
      }
      topk_flush(lst, cnt, k, select_min, out_idx + qi * k, out_dist + qi * k, -1);
      if (metric == M_L2SqrtExpanded || metric == M_L2SqrtUnexpanded)
        for (int j = 0; j < k && j < cnt; ++j) out_dist[qi * k + j] = sqrtf(out_dist[qi * k + j]);
    }
    free(lst);
  }
  free(xn);
}

/* Batched top-k over [batch,len] rows; in_idx may be NULL (=> position). Ties -> smaller position. */
void oracle_select_k(const float* in_val, const int64_t* in_idx, int64_t batch, int64_t len, int k,
                     int select_min, int64_t* out_idx, float* out_val)
{
#pragma omp parallel
  {
    pair_t* lst = (pair_t*)malloc(sizeof(pair_t) * (size_t)k);
#pragma omp for schedule(static)
    for (int64_t r = 0; r < batch; ++r) {
      int cnt = 0;
      for (int64_t c = 0; c < len; ++c) topk_push(lst, &cnt, k, in_val[r * len + c], c, select_min);
      for (int j = 0; j < k; ++j) {
        if (j < cnt) {
          out_val[r * k + j] = lst[j].v;
          out_idx[r * k + j] = in_idx ? in_idx[r * len + lst[j].i] : lst[j].i;
        } else {
          out_val[r * k + j] = select_min ? 3.402823466e+38f : -3.402823466e+38f;
          out_idx[r * k + j] = -1;
        }
      }
    }
    free(lst);
  }
}

/* ---------------------------------------------------------------- fp_8bit<5, Signed> */
/* ivf_pq_fp_8bit.cuh:31-84 with ExpBits = 5: ExpMask = 15, ValBits = 3 */
uint8_t oracle_fp8_encode(float v, int is_signed)
{
  const uint32_t ExpMask = 15u, ExpBits = 5u, ValBits = 3u;
  const float kMin = 1.0f / (float)(1u << ExpMask);
  const float kMax = (float)(1u << (ExpMask + 1)) * (2.0f - 1.0f / (float)(1u << ValBits));
  float a = is_signed ? fabsf(v) : v;
  uint8_t u;
  if (a < kMin) u = 0;
  else if (a >= kMax) u = 0xffu;
  else {
    uint32_t bits; memcpy(&bits, &a, 4);
    u = (uint8_t)((bits + (ExpMask << 23u) - 0x3f800000u) >> (15u + ExpBits));
  }
  if (is_signed) u = (uint8_t)((u & 0xfeu) | (v < 0 ? 1u : 0u));
  return u;
}
float oracle_fp8_decode(uint8_t b, int is_signed)
{
  const uint32_t ExpMask = 15u, ExpBits = 5u, ValBits = 3u;
  uint32_t u = b;
  if (is_signed) u &= ~1u;
  const uint32_t kBase32 = (0x3f800000u | (0x00400000u >> ValBits)) - (ExpMask << 23);
  uint32_t bits = kBase32 + (u << (15u + ExpBits));
  float r; memcpy(&r, &bits, 4);
  if (is_signed && (b & 1)) r = -r;
  return r;
}

/* IEEE binary16 round-to-nearest-even conversion (for lut_dtype / internal_distance_dtype = half) */
static uint16_t f2h(float f)
{
  uint32_t x; memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  int32_t e = (int32_t)((x >> 23) & 0xff) - 127 + 15;
  uint32_t m = x & 0x7fffffu;
  if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(sign | 0x7c00u | (m ? 0x200u : 0));
  if (e >= 31) return (uint16_t)(sign | 0x7c00u);
  if (e <= 0) {
    if (e < -10) return (uint16_t)sign;
    m |= 0x800000u;
    uint32_t shift = (uint32_t)(14 - e);
    uint32_t hm = m >> shift;
    uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (hm & 1))) hm++;
    return (uint16_t)(sign | hm);
  }
  uint32_t hm = m >> 13, rem = m & 0x1fffu;
  uint16_t h = (uint16_t)(sign | ((uint32_t)e << 10) | hm);
  if (rem > 0x1000u || (rem == 0x1000u && (hm & 1))) h++;
  return h;
}
static float h2f(uint16_t h)
{
  uint32_t sign = ((uint32_t)h & 0x8000u) << 16, e = (h >> 10) & 0x1f, m = h & 0x3ffu, x;
  if (e == 0) {
    if (m == 0) x = sign;
    else { int s = 0; while (!(m & 0x400u)) { m <<= 1; ++s; } m &= 0x3ffu; x = sign | ((uint32_t)(113 - s) << 23) | (m << 13); }
  } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
  else x = sign | ((e + 112) << 23) | (m << 13);
  float f; memcpy(&f, &x, 4); return f;
}
float oracle_half_round(float f) { return h2f(f2h(f)); }

/* ---------------------------------------------------------------- IVF coarse search */
/* qc[i,j] = |c_j|^2 - 2 q_i.c_j (L2) or -(q_i.c_j) (IP/cos); choose n_probes smallest.
 * centers: [n_lists, ldc] with the first `dim` columns the center (the norm is recomputed here). */
void oracle_ivf_coarse(const float* centers, int64_t n_lists, int ldc, const float* qs, int64_t nq, int dim,
                       int metric, int n_probes, uint32_t* out_labels /*[nq,n_probes]*/)
{
  float* cn = (float*)malloc(sizeof(float) * (size_t)n_lists);
  for (int64_t j = 0; j < n_lists; ++j) cn[j] = dotf(centers + j * ldc, centers + j * ldc, dim);
#pragma omp parallel
  {
    pair_t* lst = (pair_t*)malloc(sizeof(pair_t) * (size_t)n_probes);
#pragma omp for schedule(static)
    for (int64_t qi = 0; qi < nq; ++qi) {
      int cnt = 0;
      for (int64_t j = 0; j < n_lists; ++j) {
        float dp = dotf(qs + qi * dim, centers + j * ldc, dim);
        float v = (metric == M_InnerProduct || metric == M_Cosine) ? -dp : fmaf(-2.0f, dp, cn[j]);
        topk_push(lst, &cnt, n_probes, v, j, 1);
      }
      for (int p = 0; p < n_probes; ++p) out_labels[qi * n_probes + p] = p < cnt ? (uint32_t)lst[p].i : 0xffffffffu;
    }
    free(lst);
  }
  free(cn);
}

/* ---------------------------------------------------------------- IVF-Flat search */
/* Lists are given flat: list_offsets[n_lists+1] into data[n_total, dim] / ids[n_total]. */
void oracle_ivf_flat_search(const float* centers, int64_t n_lists, const int64_t* list_offsets, const float* data,
                            const int64_t* ids, const float* qs, int64_t nq, int dim, int metric, int n_probes,
                            int k, int64_t* out_idx, float* out_dist)
{
  uint32_t* labels = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(nq * n_probes));
  oracle_ivf_coarse(centers, n_lists, dim, qs, nq, dim, metric, n_probes, labels);
  const int select_min = metric != M_InnerProduct;
#pragma omp parallel
  {
    pair_t* lst = (pair_t*)malloc(sizeof(pair_t) * (size_t)k);
#pragma omp for schedule(dynamic, 4)
    for (int64_t qi = 0; qi < nq; ++qi) {
      const float* q = qs + qi * dim;
      int cnt = 0;
      for (int p = 0; p < n_probes; ++p) {
        uint32_t l = labels[qi * n_probes + p];
        if (l == 0xffffffffu) continue;
        for (int64_t r = list_offsets[l]; r < list_offsets[l + 1]; ++r) {
          float v = (metric == M_InnerProduct) ? dotf(q, data + r * dim, dim) : sqdiff(q, data + r * dim, dim);
          topk_push(lst, &cnt, k, v, ids[r], select_min);
        }
      }
      /* kOutOfBoundsRecord = max(IdxT) for missing results (ivf_common.cuh:25-31) */
      topk_flush(lst, cnt, k, select_min, out_idx + qi * k, out_dist + qi * k, INT64_MAX);
      if (metric == M_L2SqrtExpanded || metric == M_L2SqrtUnexpanded)
        for (int j = 0; j < k && j < cnt; ++j) out_dist[qi * k + j] = sqrtf(out_dist[qi * k + j]);
    }
    free(lst);
  }
  free(labels);
}

/* ---------------------------------------------------------------- IVF-PQ search */
/*
 * centers      [n_lists, dim]         cluster centers (un-rotated)
 * centers_rot  [n_lists, rot_dim]
 * rotation     [rot_dim, dim]
 * pq_centers   PER_SUBSPACE: [pq_dim, pq_len, 2^bits]; PER_CLUSTER: [n_lists, pq_len, 2^bits]
 * codes        [n_total, pq_dim] uint8, one unpacked code per byte, lists flat via list_offsets
 * lut_dtype    0 = f32, 1 = f16, 2 = fp_8bit<5>; dist_dtype 0 = f32, 1 = f16
 */
void oracle_ivf_pq_search(const float* centers, const float* centers_rot, const float* rotation,
                          const float* pq_centers, int per_cluster, int64_t n_lists, int dim, int rot_dim,
                          int pq_dim, int pq_bits, const int64_t* list_offsets, const uint8_t* codes,
                          const int64_t* ids, const float* qs, int64_t nq, int metric, int n_probes, int k,
                          int lut_dtype, int dist_dtype, int64_t* out_idx, float* out_dist)
{
  const int pq_len = rot_dim / pq_dim;
  const int book = 1 << pq_bits;
  uint32_t* labels = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(nq * n_probes));
  oracle_ivf_coarse(centers, n_lists, dim, qs, nq, dim, metric, n_probes, labels);
  const int is_ip = (metric == M_InnerProduct || metric == M_Cosine);
#pragma omp parallel
  {
    pair_t* lst = (pair_t*)malloc(sizeof(pair_t) * (size_t)k);
    float* qrot = (float*)malloc(sizeof(float) * (size_t)rot_dim);
    float* lut = (float*)malloc(sizeof(float) * (size_t)(pq_dim * book));
#pragma omp for schedule(dynamic, 2)
    for (int64_t qi = 0; qi < nq; ++qi) {
      for (int r = 0; r < rot_dim; ++r) qrot[r] = dotf(rotation + (int64_t)r * dim, qs + qi * dim, dim);
      int cnt = 0;
      for (int p = 0; p < n_probes; ++p) {
        uint32_t l = labels[qi * n_probes + p];
        if (l == 0xffffffffu) continue;
        const float* crot = centers_rot + (int64_t)l * rot_dim;
        const float* pqc = per_cluster ? pq_centers + (int64_t)l * pq_len * book : pq_centers;
        /* create_lut_impl.cuh:40-77 */
        for (int i = 0; i < pq_dim; ++i) {
          for (int c = 0; c < book; ++c) {
            float score = 0.0f;
            for (int t = 0; t < pq_len; ++t) {
              int j = i * pq_len + t;
              float pq_c = per_cluster ? pqc[t * book + c] : pqc[((int64_t)i * pq_len + t) * book + c];
              if (!is_ip) {
                float diff = qrot[j] - crot[j];
                diff -= pq_c;
                score = fmaf(diff, diff, score);
              } else {
                score = fmaf(-qrot[j], crot[j], score);
                score = fmaf(-qrot[j], pq_c, score);
              }
            }
            if (lut_dtype == 1) score = oracle_half_round(score);
            else if (lut_dtype == 2) score = oracle_fp8_decode(oracle_fp8_encode(score, is_ip), is_ip);
            lut[i * book + c] = score;
          }
        }
        /* compute_score_impl.cuh:53-79: sequential accumulation in OutT */
        for (int64_t r = list_offsets[l]; r < list_offsets[l + 1]; ++r) {
          const uint8_t* code = codes + r * pq_dim;
          float score = 0.0f;
          for (int i = 0; i < pq_dim; ++i) {
            score += lut[i * book + code[i]];
            if (dist_dtype == 1) score = oracle_half_round(score);
          }
          if (metric == M_Cosine) score += 1.0f;
          topk_push(lst, &cnt, k, score, ids[r], 1);
        }
      }
      topk_flush(lst, cnt, k, 1, out_idx + qi * k, out_dist + qi * k, INT64_MAX);
      /* ivf_common.cuh:175-252 postprocess_distances */
      for (int j = 0; j < k && j < cnt; ++j) {
        float* v = out_dist + qi * k + j;
        if (metric == M_L2SqrtExpanded || metric == M_L2SqrtUnexpanded) *v = sqrtf(*v);
        else if (metric == M_InnerProduct) *v = -*v;
      }
    }
    free(lst); free(qrot); free(lut);
  }
  free(labels);
}

/* ---------------------------------------------------------------- CAGRA single-CTA walk */
static inline uint64_t xorshift64(uint64_t u)
{
  u ^= u >> 12; u ^= u << 25; u ^= u >> 27;
  return u * 0x2545F4914F6CDD1DULL;
}
/* hashmap.hpp:37-73, open addressing.  The reference compiles its linear-probing variant (HASHMAP_LINEAR_PROBING); this is
 * the double-hashing branch of the same file (:52-55).  Both are exact sets below capacity, which the search plan guarantees
 * (fill rate <= 50 %), so the walk does not depend on the probing scheme. */
static int hash_insert(uint32_t* table, uint32_t bitlen, uint32_t key)
{
  const uint32_t size = 1u << bitlen, mask = size - 1;
  uint32_t index = key & mask;
  const uint32_t stride = (key >> bitlen) * 2 + 1;
  for (uint32_t i = 0; i < size; ++i) {
    if (table[index] == 0xffffffffu) { table[index] = key; return 1; }
    if (table[index] == key) return 0;
    index = (index + stride) & mask;
  }
  return 0;
}

typedef struct { float d; uint32_t i; uint32_t pos; } cand_t;
static int cand_cmp(const void* a, const void* b)
{
  const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
  if (x->d < y->d) return -1;
  if (x->d > y->d) return 1;
  return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0);
}

/*
 * graph [n, degree] u32; dataset [n, dim]; metric L2Expanded (squared L2) or InnerProduct.
 * hash_bitlen / small_hash_bitlen / reset_interval follow search_plan.cuh:256-372
 * (small_hash_bitlen != 0 => table of that size, cleared and re-seeded with the itopk every
 * reset_interval iterations).  out_idx [nq,k] u32, out_dist [nq,k]; n_iters (may be NULL) [nq].
 */
void oracle_cagra_search(const uint32_t* graph, const float* dataset, int64_t n, int dim, int degree,
                         const float* qs, int64_t nq, int metric, int k, int itopk, int search_width,
                         int min_iter, int max_iter, uint32_t hash_bitlen, uint32_t small_hash_bitlen,
                         uint32_t reset_interval, int num_random_samplings, uint64_t rand_xor_mask,
                         uint32_t* out_idx, float* out_dist, uint32_t* n_iters)
{
  const uint32_t MSB = 0x80000000u, INVALID = 0xffffffffu;
  const int n_cand = search_width * degree;
  const int buf = itopk + n_cand;
  const uint32_t bitlen = small_hash_bitlen ? small_hash_bitlen : hash_bitlen;
#pragma omp parallel
  {
    cand_t* rb = (cand_t*)malloc(sizeof(cand_t) * (size_t)buf);
    uint32_t* table = (uint32_t*)malloc(sizeof(uint32_t) << bitlen);
    uint32_t* parents = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)search_width);
#pragma omp for schedule(dynamic, 8)
    for (int64_t qi = 0; qi < nq; ++qi) {
      const float* q = qs + qi * dim;
      memset(table, 0xff, sizeof(uint32_t) << bitlen);
      /* device_common_jit.cuh:36-112: random seeds, best of num_random_samplings, dedup by hash */
      for (int i = 0; i < buf; ++i) {
        float best = INFINITY; uint32_t bi = INVALID;
        for (int j = 0; j < num_random_samplings; ++j) {
          uint64_t gid = (uint64_t)i + (uint64_t)buf * (uint64_t)j;
          uint32_t s = (uint32_t)(xorshift64(gid ^ rand_xor_mask) % (uint64_t)n);
          float dd = (metric == M_InnerProduct) ? -dotf(q, dataset + (int64_t)s * dim, dim)
                                                : sqdiff(q, dataset + (int64_t)s * dim, dim);
          if (dd < best) { best = dd; bi = s; }
        }
        if (bi != INVALID && !hash_insert(table, bitlen, bi)) { best = INFINITY; bi = INVALID; }
        rb[i].d = bi == INVALID ? INFINITY : best; rb[i].i = bi; rb[i].pos = (uint32_t)i;
      }
      uint32_t iter = 0;
      for (;;) {
        if (small_hash_bitlen && (iter + 1) % reset_interval == 0) memset(table, 0xff, sizeof(uint32_t) << bitlen);
        /* topk_by_bitonic_sort_and_merge: keep the itopk best of itopk ∪ candidates, sorted */
        for (int i = 0; i < buf; ++i) rb[i].pos = (uint32_t)i;
        qsort(rb, (size_t)buf, sizeof(cand_t), cand_cmp);
        if ((int)(iter + 1) == max_iter) break;
        /* pickup_next_parents: first search_width entries without the MSB flag */
        int np = 0;
        for (int j = 0; j < itopk && np < search_width; ++j) {
          if (rb[j].i != INVALID && (rb[j].i & MSB) == 0) { parents[np++] = (uint32_t)j; rb[j].i |= MSB; }
          else if (rb[j].i == INVALID) { /* INVALID has the MSB set: never a parent */ }
        }
        if (small_hash_bitlen && (iter + 1) % reset_interval == 0)
          for (int j = 0; j < itopk; ++j) if (rb[j].i != INVALID) hash_insert(table, bitlen, rb[j].i & ~MSB);
        if (np == 0 && (int)iter >= min_iter) break;
        /* compute_distance_to_child_nodes */
        for (int c = 0; c < n_cand; ++c) {
          uint32_t child = INVALID;
          int pi = c / degree;
          if (pi < np) {
            uint32_t parent = rb[parents[pi]].i & ~MSB;
            child = graph[(int64_t)parent * degree + (c % degree)];
          }
          if (child != INVALID && !hash_insert(table, bitlen, child)) child = INVALID;
          rb[itopk + c].i = child;
          rb[itopk + c].d = child == INVALID ? INFINITY
                            : ((metric == M_InnerProduct) ? -dotf(q, dataset + (int64_t)child * dim, dim)
                                                          : sqdiff(q, dataset + (int64_t)child * dim, dim));
        }
        ++iter;
      }
      for (int j = 0; j < k; ++j) {
        uint32_t id = rb[j].i;
        out_idx[qi * k + j] = id == INVALID ? INVALID : (id & ~MSB);
        float dd = rb[j].d;
        out_dist[qi * k + j] = (id == INVALID) ? 3.402823466e+38f : (metric == M_InnerProduct ? -dd : dd);
      }
      if (n_iters) n_iters[qi] = iter + 1;
    }
    free(rb); free(table); free(parents);
  }
}

/* ---------------------------------------------------------------- k-means assignment */
/* argmin_j of expanded L2 (clamped at 0); ties -> smaller j (unfused_distance_nn.cuh:41-47,79-83) */
void oracle_kmeans_assign(const float* x, int64_t n, const float* c, int kc, int d, int32_t* labels, float* mind)
{
  float* cn = (float*)malloc(sizeof(float) * (size_t)kc);
  for (int j = 0; j < kc; ++j) cn[j] = dotf(c + (int64_t)j * d, c + (int64_t)j * d, d);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const float xn = dotf(x + i * d, x + i * d, d);
    float best = INFINITY; int bj = 0;
    for (int j = 0; j < kc; ++j) {
      float v = fmaf(-2.0f, dotf(x + i * d, c + (int64_t)j * d, d), xn + cn[j]);
      if (!(v > 0.0f)) v = 0.0f;
      if (v < best) { best = v; bj = j; }
    }
    labels[i] = bj; if (mind) mind[i] = best;
  }
  free(cn);
}

void oracle_set_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
