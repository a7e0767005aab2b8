// Exact fp32 distance kernels in the oracle's pinned arithmetic (host API).
//
// Every dot product / norm / squared difference is accumulated with fmaf in ascending component
// order, exactly as oracle/oracle.c does, so results can be compared bit for bit.  These kernels
// are (a) the re-scoring stage behind every tensor-core candidate scan, (b) the certified fallback
// of brute force, (c) the filtered / exotic-shape path.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include <cuvs/distance/distance.h>

namespace b200 {

/** out[i] = sum_k x[i,k]^2 (ascending k, fmaf). */
void row_norms(cudaStream_t stream, const float* x, int64_t n, int d, int64_t ld, float* out);

struct filter_view {
  const uint32_t* bits = nullptr;  // 1 = keep
  int kind             = 0;        // 0 none, 1 bitset (per sample), 2 bitmap (per query x sample)
  int64_t n_samples    = 0;
};

/**
 * Dense distance tile: out[i, j] = dist(q_i, x_j) for i < nq, j < n, in the metric's *selection*
 * form (squared for the L2Sqrt variants; plain dot for InnerProduct).  qn/xn: squared row norms
 * (needed for L2*Expanded and Cosine, may be null otherwise).  Filtered-out pairs get the worst
 * value.  `q_row0` is the global query index of row 0 (bitmap addressing).
 */
void exact_distance_tile(cudaStream_t stream, const float* q, int64_t nq, int64_t ldq, const float* x, int64_t n,
                         int64_t ldx, int d, const float* qn, const float* xn, cuvsDistanceType metric, float* out,
                         int64_t ldo, filter_view filt, int64_t q_row0);

/** How a raw scan score s of the tensor-core engine maps to the metric's selection form, and the
 *  error budget of that approximation:  A = sa*s + sb*|q|^2 + sc,  eps = eps_rel * (eq*|q|^2 + ec). */
struct approx_map {
  float sa = 1.f, sb = 0.f, sc = 0.f;
  float eps_rel = 0.f, eq = 0.f, ec = 0.f;
};

/**
 * Re-score candidates exactly and keep the best k per query.
 *   cand_pos [nq, kc]  uint32 row positions into x (0xffffffff = empty slot)
 *   src_ids            optional int64 map position -> source id (null: identity)
 * Output rows are sorted best-first, ties towards the smaller *source id*; missing entries get
 * id `pad_id` and +/-FLT_MAX.  When `cand_score` (approximate selection-form scores of the
 * candidates = raw engine scores mapped through `amap`, same layout, sorted or not) and `flags` are given, flags[i] is set to 1 unless the
 * k-th exact distance is strictly better than (worst candidate approx score -/+ eps_i), where
 * eps_i comes from `amap` (for L2: eps_rel * (|q_i|^2 + max_j |x_j|^2)): the certificate that no non-candidate can
 * belong to the exact top-k.
 */
void rescore_topk(cudaStream_t stream, const float* q, int64_t nq, int64_t ldq, const float* x, int64_t ldx, int d,
                  const float* qn, const float* xn, cuvsDistanceType metric, const uint32_t* cand_pos,
                  const float* cand_score, int kc, const int64_t* src_ids, int k, int64_t* out_idx, float* out_dist,
                  int64_t pad_id, const struct approx_map& amap, int* flags, int* n_flagged,
                  const float* approx_floor = nullptr /* [nq] raw engine score no row outside the candidates can beat */);

/** In-place sqrt (L2Sqrt*) or sign/offset fix-ups applied after selection. */
void postprocess_distances(cudaStream_t stream, float* dist, int64_t count, cuvsDistanceType metric);

}  // namespace b200
