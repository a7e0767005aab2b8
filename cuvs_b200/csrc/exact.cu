// Exact fp32 kernels (oracle arithmetic): norms, dense distance tiles, candidate re-scoring.
// See exact.cuh.  Reference semantics restated:
//   expanded L2 + clamps      cpp/src/distance/detail/distance_ops/l2_exp.cuh:100-118
//   unexpanded / IP / cosine  cpp/tests/neighbors/naive_knn.cuh:34-86
//   post-selection sqrt       cpp/src/neighbors/detail/knn_brute_force.cuh:468-479
#include "common.hpp"
#include "exact.cuh"
#include "timing.hpp"

#include <cfloat>

namespace b200 {
namespace {

__global__ void row_norms_kernel(const float* __restrict__ x, int64_t n, int d, int64_t ld, float* __restrict__ out)
{
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float* r = x + i * ld;
  float acc      = 0.f;
  if ((d & 3) == 0 && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const float4* r4 = reinterpret_cast<const float4*>(r);
    for (int k = 0; k < d / 4; ++k) {
      float4 v = r4[k];
      acc      = fmaf(v.x, v.x, acc);
      acc      = fmaf(v.y, v.y, acc);
      acc      = fmaf(v.z, v.z, acc);
      acc      = fmaf(v.w, v.w, acc);
    }
  } else {
    for (int k = 0; k < d; ++k) acc = fmaf(r[k], r[k], acc);
  }
  out[i] = acc;
}

__device__ __forceinline__ bool keep_sample(const filter_view& f, int64_t q, int64_t j)
{
  if (f.kind == 0) return true;
  int64_t bit = f.kind == 1 ? j : q * f.n_samples + j;
  return (f.bits[bit >> 5] >> (bit & 31)) & 1u;
}

__device__ __forceinline__ float finish_distance(float acc, float qn, float xn, int metric)
{
  switch (metric) {
    case L2Expanded:
    case L2SqrtExpanded: {
      float val = fmaf(-2.0f, acc, qn + xn);
      if (!(val > 0.0f)) val = 0.0f;
      if (val * val < 1e-6f && qn == xn) val = 0.0f;
      return val;
    }
    case CosineExpanded: return 1.0f - acc / (sqrtf(qn) * sqrtf(xn));
    default: return acc;  // InnerProduct, L2Unexpanded (acc already holds the sum of squares)
  }
}

constexpr int TM = 64, TN = 64, TK = 16;

// 64x64 output tile per CTA, 256 threads, 4x4 micro-tile per thread; every output accumulates over
// k = 0..d-1 in order (one fmaf per k), which is the oracle's order.
template <bool SqDiff>
__global__ void __launch_bounds__(256) exact_tile_kernel(const float* __restrict__ q, int64_t nq, int64_t ldq,
                                                           const float* __restrict__ x, int64_t n, int64_t ldx, int d,
                                                           const float* __restrict__ qn, const float* __restrict__ xn,
                                                           int metric, float* __restrict__ out, int64_t ldo,
                                                           filter_view filt, int64_t q_row0)
{
  __shared__ float As[TK][TM + 1];
  __shared__ float Bs[TK][TN + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = static_cast<int64_t>(blockIdx.y) * TM, col0 = static_cast<int64_t>(blockIdx.x) * TN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < d; k0 += TK) {
    // 64 rows x 16 k = 1024 elements per operand, 4 per thread
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = tid + e * 256;
      int r = idx >> 4, kk = idx & 15;
      int64_t gr = row0 + r, gc = col0 + r;
      As[kk][r] = (gr < nq && k0 + kk < d) ? q[gr * ldq + k0 + kk] : 0.f;
      Bs[kk][r] = (gc < n && k0 + kk < d) ? x[gc * ldx + k0 + kk] : 0.f;
    }
    __syncthreads();
    const int kmax = min(TK, d - k0);
    for (int kk = 0; kk < kmax; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if constexpr (SqDiff) {
            float t   = a[i] - b[j];
            acc[i][j] = fmaf(t, t, acc[i][j]);
          } else {
            acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
          }
        }
    }
    __syncthreads();
  }
  const bool select_min = metric != InnerProduct;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t gr = row0 + ty * 4 + i;
    if (gr >= nq) continue;
    float qnv = qn ? qn[gr] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t gc = col0 + tx * 4 + j;
      if (gc >= n) continue;
      float v = finish_distance(acc[i][j], qnv, xn ? xn[gc] : 0.f, metric);
      if (!keep_sample(filt, q_row0 + gr, gc)) v = select_min ? FLT_MAX : -FLT_MAX;
      out[gr * ldo + gc] = v;
    }
  }
}

constexpr int kMaxCand = 256;

// one warp per query
__global__ void __launch_bounds__(128) rescore_kernel(const float* __restrict__ q, int64_t nq, int64_t ldq,
                                                       const float* __restrict__ x, int64_t ldx, int d,
                                                       const float* __restrict__ qn, const float* __restrict__ xn,
                                                       int metric, const uint32_t* __restrict__ cand_pos,
                                                       const float* __restrict__ cand_score, int kc,
                                                       const int64_t* __restrict__ src_ids, int k,
                                                       int64_t* __restrict__ out_idx, float* __restrict__ out_dist,
                                                       int64_t pad_id, approx_map amap,
                                                       int* __restrict__ flags, int* __restrict__ n_flagged,
                                                       const float* __restrict__ approx_floor)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sd       = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(wid) * kc;
  int64_t* si     = reinterpret_cast<int64_t*>(reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(warps) * kc) +
                static_cast<size_t>(wid) * kc;
  const int64_t qi = static_cast<int64_t>(blockIdx.x) * warps + wid;
  if (qi >= nq) return;
  const bool select_min = metric != InnerProduct;
  const float* qr       = q + qi * ldq;
  const float qnv       = qn ? qn[qi] : 0.f;
  const bool sqdiff     = (metric == L2Unexpanded || metric == L2SqrtUnexpanded);
  const bool vec4       = ((d & 3) == 0) && ((ldx & 3) == 0) && ((ldq & 3) == 0) &&
                    ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(q) & 15) == 0);

  float worst_approx = select_min ? -FLT_MAX : FLT_MAX;
  int n_valid        = 0;
  for (int c = lane; c < kc; c += 32) {
    uint32_t pos = cand_pos[qi * kc + c];
    float dist   = select_min ? FLT_MAX : -FLT_MAX;
    int64_t id   = INT64_MAX;
    if (pos != 0xffffffffu) {
      const float* xr = x + static_cast<int64_t>(pos) * ldx;
      float acc       = 0.f;
      if (vec4) {
        const float4* x4 = reinterpret_cast<const float4*>(xr);
        const float4* q4 = reinterpret_cast<const float4*>(qr);
        for (int kk = 0; kk < d / 4; ++kk) {
          float4 a = q4[kk], b = x4[kk];
          if (sqdiff) {
            float t;
            t = a.x - b.x; acc = fmaf(t, t, acc);
            t = a.y - b.y; acc = fmaf(t, t, acc);
            t = a.z - b.z; acc = fmaf(t, t, acc);
            t = a.w - b.w; acc = fmaf(t, t, acc);
          } else {
            acc = fmaf(a.x, b.x, acc);
            acc = fmaf(a.y, b.y, acc);
            acc = fmaf(a.z, b.z, acc);
            acc = fmaf(a.w, b.w, acc);
          }
        }
      } else {
        for (int kk = 0; kk < d; ++kk) {
          if (sqdiff) { float t = qr[kk] - xr[kk]; acc = fmaf(t, t, acc); }
          else acc = fmaf(qr[kk], xr[kk], acc);
        }
      }
      dist = finish_distance(acc, qnv, xn ? xn[pos] : 0.f, metric);
      id   = src_ids ? src_ids[pos] : static_cast<int64_t>(pos);
      ++n_valid;
      if (cand_score) {
        float a      = amap.sa * cand_score[qi * kc + c] + amap.sb * qnv + amap.sc;
        worst_approx = select_min ? fmaxf(worst_approx, a) : fminf(worst_approx, a);
      }
    }
    sd[c] = dist;
    si[c] = id;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
    float t = __shfl_xor_sync(0xffffffffu, worst_approx, o);
    worst_approx = select_min ? fmaxf(worst_approx, t) : fminf(worst_approx, t);
  }
  __syncwarp();
  // rank by (distance, id, slot): unique ranks even among empty slots
  float kth = select_min ? FLT_MAX : -FLT_MAX;
  for (int c = lane; c < kc; c += 32) {
    float dv = sd[c];
    int64_t iv = si[c];
    int rank = 0;
    for (int o = 0; o < kc; ++o) {
      float d2 = sd[o];
      int64_t i2 = si[o];
      bool better = (d2 != dv) ? (select_min ? d2 < dv : d2 > dv) : ((i2 != iv) ? (i2 < iv) : (o < c));
      rank += better ? 1 : 0;
    }
    if (rank < k) {
      bool valid = iv != INT64_MAX;
      out_idx[qi * k + rank]  = valid ? iv : pad_id;
      out_dist[qi * k + rank] = dv;
      if (rank == k - 1) kth = dv;
    }
  }
  for (int j = kc + lane; j < k; j += 32) {  // k > kc: pad
    out_idx[qi * k + j]  = pad_id;
    out_dist[qi * k + j] = select_min ? FLT_MAX : -FLT_MAX;
  }
  if (flags) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float t = __shfl_xor_sync(0xffffffffu, kth, o);
      kth     = select_min ? fminf(kth, t) : fmaxf(kth, t);
    }
    if (lane == 0) {
      int flag = 0;
      bool outside = n_valid == kc;  // candidate list is full: rows outside it exist (dropped by the merge to kc)
      if (approx_floor != nullptr && cand_score) {
        // a second bound, in raw engine units, on the rows that never entered a candidate list (brute force with k above the
        // fused list length: such rows are bounded by the lists' own worst entries, not by the merged set's); +inf = none
        const float fl = approx_floor[qi];
        if (fl < FLT_MAX) {
          const float a = amap.sa * fl + amap.sb * qnv + amap.sc;
          worst_approx  = !outside ? a : (select_min ? fminf(worst_approx, a) : fmaxf(worst_approx, a));
          outside       = true;
        }
      }
      if (outside && cand_score) {
        float eps = amap.eps_rel * (amap.eq * qnv + amap.ec);
        bool ok   = select_min ? (kth < worst_approx - eps) : (kth > worst_approx + eps);
        flag      = ok ? 0 : 1;
      }
      flags[qi] = flag;
      if (flag && n_flagged) atomicAdd(n_flagged, 1);
    }
  }
}

__global__ void postprocess_kernel(float* dist, int64_t count, int metric)
{
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= count) return;
  float v = dist[i];
  if (metric == L2SqrtExpanded || metric == L2SqrtUnexpanded) {
    if (v != FLT_MAX) dist[i] = sqrtf(fmaxf(v, 0.f));
  }
}

}  // namespace

void row_norms(cudaStream_t stream, const float* x, int64_t n, int d, int64_t ld, float* out)
{
  if (n == 0) return;
  count_launch();
  row_norms_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, stream>>>(x, n, d, ld, out);
  B2_CUDA(cudaGetLastError());
}

void exact_distance_tile(cudaStream_t stream, const float* q, int64_t nq, int64_t ldq, const float* x, int64_t n,
                         int64_t ldx, int d, const float* qn, const float* xn, cuvsDistanceType metric, float* out,
                         int64_t ldo, filter_view filt, int64_t q_row0)
{
  if (nq == 0 || n == 0) return;
  dim3 grid(static_cast<unsigned>((n + TN - 1) / TN), static_cast<unsigned>((nq + TM - 1) / TM));
  B2_EXPECTS(grid.y <= 65535, "exact_distance_tile: too many query rows per call");
  const bool sq = (metric == L2Unexpanded || metric == L2SqrtUnexpanded);
  count_launch();
  timed_section ts("exact_tile", stream);
  if (sq)
    exact_tile_kernel<true><<<grid, 256, 0, stream>>>(q, nq, ldq, x, n, ldx, d, qn, xn, int(metric), out, ldo, filt, q_row0);
  else
    exact_tile_kernel<false><<<grid, 256, 0, stream>>>(q, nq, ldq, x, n, ldx, d, qn, xn, int(metric), out, ldo, filt, q_row0);
  B2_CUDA(cudaGetLastError());
}

void rescore_topk(cudaStream_t stream, const float* q, int64_t nq, int64_t ldq, const float* x, int64_t ldx, int d,
                  const float* qn, const float* xn, cuvsDistanceType metric, const uint32_t* cand_pos,
                  const float* cand_score, int kc, const int64_t* src_ids, int k, int64_t* out_idx, float* out_dist,
                  int64_t pad_id, const approx_map& amap, int* flags, int* n_flagged, const float* approx_floor)
{
  if (nq == 0) return;
  B2_EXPECTS(kc >= 1 && kc <= kMaxCand, "rescore_topk: candidate count %d out of range", kc);
  const int warps = 4;
  size_t smem     = static_cast<size_t>(warps) * kc * (sizeof(float) + sizeof(int64_t));
  count_launch();
  rescore_kernel<<<static_cast<unsigned>((nq + warps - 1) / warps), warps * 32, smem, stream>>>(
    q, nq, ldq, x, ldx, d, qn, xn, int(metric), cand_pos, cand_score, kc, src_ids, k, out_idx, out_dist, pad_id, amap,
    flags, n_flagged, approx_floor);
  B2_CUDA(cudaGetLastError());
}

void postprocess_distances(cudaStream_t stream, float* dist, int64_t count, cuvsDistanceType metric)
{
  if (count == 0) return;
  if (metric != L2SqrtExpanded && metric != L2SqrtUnexpanded) return;
  count_launch();
  postprocess_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, stream>>>(dist, count, int(metric));
  B2_CUDA(cudaGetLastError());
}

}  // namespace b200
