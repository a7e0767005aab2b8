// Single-process multi-GPU IVF-Flat (cuvsMultiGpuIvfFlat*) and dense pairwise distance.
//
// Reference: cpp/src/neighbors/mg/snmg.cuh:92-707 (replicated / row-sharded index, OpenMP thread per GPU, NCCL
// send/recv to a root + knn_merge_parts), c/src/neighbors/mg_ivf_flat.cpp.  This entry point keeps the reference's
// contract (HOST queries / outputs, row-sharded or replicated) for bindings that use it; the B200-native sharding
// (by IVF list, one process per GPU, NCCL all-gather) lives in cuvs_b200/distributed.py — see DESIGN.md §7.
// Here every device runs the ordinary per-device C entry points on its own stream; the per-shard partials are
// merged on the host (the payload is nq*k*12 bytes per shard).
#include "common.hpp"
#include "exact.cuh"
#include "timing.hpp"

#include <cuvs/distance/pairwise_distance.h>
#include <cuvs/neighbors/mg_ivf_flat.h>

#include <algorithm>
#include <cfloat>
#include <memory>
#include <vector>

namespace b200 {
namespace {

struct mg_shard {
  int device = 0;
  cuvsResources_t res = 0;
  cuvsIvfFlatIndex_t index = nullptr;
  int64_t row0 = 0, rows = 0;
};

struct mg_index {
  cuvsMultiGpuDistributionMode mode = CUVS_NEIGHBORS_MG_SHARDED;
  cuvsDistanceType metric = L2Expanded;
  int dim = 0;
  std::vector<mg_shard> shards;
  ~mg_index()
  {
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& s : shards) {
      cudaSetDevice(s.device);
      if (s.index) cuvsIvfFlatIndexDestroy(s.index);
      if (s.res) cuvsResourcesDestroy(s.res);
    }
    cudaSetDevice(cur);
  }
};

struct device_guard {
  int prev = 0;
  explicit device_guard(int d) { cudaGetDevice(&prev); cudaSetDevice(d); }
  ~device_guard() { cudaSetDevice(prev); }
};

void check_c(cuvsError_t e, const char* what)
{
  if (e != CUVS_SUCCESS) {
    const char* t = cuvsGetLastErrorText();
    std::string msg = t ? t : "unknown error";
    B2_FAIL("%s: %s", what, msg.c_str());
  }
}

DLManagedTensor make_dl(void* p, DLDeviceType dt, int dev, DLDataType ty, int ndim, int64_t* shape)
{
  DLManagedTensor m{};
  m.dl_tensor.data = p; m.dl_tensor.device = DLDevice{dt, dev}; m.dl_tensor.ndim = ndim; m.dl_tensor.dtype = ty; m.dl_tensor.shape = shape;
  return m;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

cuvsError_t cuvsMultiGpuIvfFlatIndexParamsCreate(cuvsMultiGpuIvfFlatIndexParams_t* index_params)
{
  return guarded([=] {
    cuvsIvfFlatIndexParams_t base;
    check_c(cuvsIvfFlatIndexParamsCreate(&base), "cuvsIvfFlatIndexParamsCreate");
    *index_params = new cuvsMultiGpuIvfFlatIndexParams{base, CUVS_NEIGHBORS_MG_SHARDED};
  });
}
cuvsError_t cuvsMultiGpuIvfFlatIndexParamsDestroy(cuvsMultiGpuIvfFlatIndexParams_t p)
{
  return guarded([=] { if (p) { cuvsIvfFlatIndexParamsDestroy(p->base_params); delete p; } });
}
cuvsError_t cuvsMultiGpuIvfFlatSearchParamsCreate(cuvsMultiGpuIvfFlatSearchParams_t* params)
{
  return guarded([=] {
    cuvsIvfFlatSearchParams_t base;
    check_c(cuvsIvfFlatSearchParamsCreate(&base), "cuvsIvfFlatSearchParamsCreate");
    *params = new cuvsMultiGpuIvfFlatSearchParams{base, CUVS_NEIGHBORS_MG_LOAD_BALANCER, CUVS_NEIGHBORS_MG_TREE_MERGE, 1LL << 20};
  });
}
cuvsError_t cuvsMultiGpuIvfFlatSearchParamsDestroy(cuvsMultiGpuIvfFlatSearchParams_t p)
{
  return guarded([=] { if (p) { cuvsIvfFlatSearchParamsDestroy(p->base_params); delete p; } });
}
cuvsError_t cuvsMultiGpuIvfFlatIndexCreate(cuvsMultiGpuIvfFlatIndex_t* index)
{
  return guarded([=] { *index = new cuvsMultiGpuIvfFlatIndex{}; });
}
cuvsError_t cuvsMultiGpuIvfFlatIndexDestroy(cuvsMultiGpuIvfFlatIndex_t index)
{
  return guarded([=] {
    if (!index) return;
    delete reinterpret_cast<mg_index*>(index->addr);
    delete index;
  });
}

cuvsError_t cuvsMultiGpuIvfFlatBuild(cuvsResources_t res, cuvsMultiGpuIvfFlatIndexParams_t params, DLManagedTensor* dataset_tensor,
                                     cuvsMultiGpuIvfFlatIndex_t index)
{
  return guarded([=] {
    auto r = as_res(res);
    B2_EXPECTS(params && dataset_tensor && index, "null argument");
    B2_EXPECTS(!r->mg_devices.empty(), "cuvsMultiGpuIvfFlatBuild needs a handle from cuvsMultiGpuResourcesCreate");
    const DLTensor& ds = dataset_tensor->dl_tensor;
    B2_EXPECTS(dl_is(ds, kDLFloat, 32) && ds.ndim == 2 && dl_is_c_contiguous(ds), "dataset must be a row-major float32 matrix");
    B2_EXPECTS(dl_is_host(ds), "multi-GPU build requires the dataset in host memory");
    auto mg     = std::make_unique<mg_index>();
    mg->mode    = params->mode;
    mg->metric  = params->base_params->metric;
    mg->dim     = static_cast<int>(ds.shape[1]);
    const int64_t n = ds.shape[0];
    const int nd    = static_cast<int>(r->mg_devices.size());
    for (int i = 0; i < nd; ++i) {
      mg_shard sh;
      sh.device = r->mg_devices[i];
      sh.row0   = mg->mode == CUVS_NEIGHBORS_MG_SHARDED ? n * i / nd : 0;
      sh.rows   = mg->mode == CUVS_NEIGHBORS_MG_SHARDED ? n * (i + 1) / nd - sh.row0 : n;
      device_guard g(sh.device);
      check_c(cuvsResourcesCreate(&sh.res), "cuvsResourcesCreate");
      check_c(cuvsStreamSet(sh.res, r->mg_streams[i]), "cuvsStreamSet");
      check_c(cuvsIvfFlatIndexCreate(&sh.index), "cuvsIvfFlatIndexCreate");
      cuvsIvfFlatIndexParams p = *params->base_params;
      p.n_lists = static_cast<uint32_t>(std::max<int64_t>(1, std::min<int64_t>(p.n_lists, sh.rows)));
      int64_t shape[2] = {sh.rows, mg->dim};
      DLManagedTensor part = make_dl(dl_ptr<float>(ds) + sh.row0 * mg->dim, kDLCPU, 0, ds.dtype, 2, shape);
      mg->shards.push_back(sh);
      check_c(cuvsIvfFlatBuild(sh.res, &p, &part, sh.index), "cuvsIvfFlatBuild (shard)");
    }
    if (index->addr) delete reinterpret_cast<mg_index*>(index->addr);
    index->addr  = reinterpret_cast<uintptr_t>(mg.release());
    index->dtype = ds.dtype;
  });
}

cuvsError_t cuvsMultiGpuIvfFlatSearch(cuvsResources_t res, cuvsMultiGpuIvfFlatSearchParams_t params, cuvsMultiGpuIvfFlatIndex_t index,
                                      DLManagedTensor* queries_tensor, DLManagedTensor* neighbors_tensor, DLManagedTensor* distances_tensor)
{
  return guarded([=] {
    as_res(res);
    B2_EXPECTS(params && index && index->addr && queries_tensor && neighbors_tensor && distances_tensor, "null argument");
    auto& mg = *reinterpret_cast<mg_index*>(index->addr);
    const DLTensor& q  = queries_tensor->dl_tensor;
    const DLTensor& nb = neighbors_tensor->dl_tensor;
    const DLTensor& dd = distances_tensor->dl_tensor;
    B2_EXPECTS(dl_is_host(q) && dl_is_host(nb) && dl_is_host(dd), "multi-GPU search takes host queries / neighbors / distances");
    B2_EXPECTS(dl_is(q, kDLFloat, 32) && dl_is(nb, kDLInt, 64) && dl_is(dd, kDLFloat, 32), "queries f32, neighbors int64, distances f32 expected");
    const int64_t nq = q.shape[0];
    const int k      = static_cast<int>(nb.shape[1]);
    const int ns     = static_cast<int>(mg.shards.size());
    const bool select_min = mg.metric != InnerProduct;
    struct part { float* dq = nullptr; int64_t* di = nullptr; float* dv = nullptr; std::vector<int64_t> hi; std::vector<float> hv; int64_t q0 = 0, qn = 0; };
    std::vector<part> parts(ns);
    // launch everything asynchronously on each device's stream, then collect
    for (int s = 0; s < ns; ++s) {
      auto& sh = mg.shards[s];
      auto& pt = parts[s];
      if (mg.mode == CUVS_NEIGHBORS_MG_SHARDED) { pt.q0 = 0; pt.qn = nq; }
      else { pt.q0 = nq * s / ns; pt.qn = nq * (s + 1) / ns - pt.q0; }
      if (pt.qn == 0) continue;
      device_guard g(sh.device);
      cudaStream_t st;
      check_c(cuvsStreamGet(sh.res, &st), "cuvsStreamGet");
      B2_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&pt.dq), sizeof(float) * pt.qn * mg.dim, st));
      B2_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&pt.di), sizeof(int64_t) * pt.qn * k, st));
      B2_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&pt.dv), sizeof(float) * pt.qn * k, st));
      B2_CUDA(cudaMemcpyAsync(pt.dq, dl_ptr<float>(q) + pt.q0 * mg.dim, sizeof(float) * pt.qn * mg.dim, cudaMemcpyHostToDevice, st));
      int64_t qs[2] = {pt.qn, mg.dim}, os[2] = {pt.qn, k};
      DLManagedTensor tq = make_dl(pt.dq, kDLCUDA, sh.device, DLDataType{kDLFloat, 32, 1}, 2, qs);
      DLManagedTensor ti = make_dl(pt.di, kDLCUDA, sh.device, DLDataType{kDLInt, 64, 1}, 2, os);
      DLManagedTensor tv = make_dl(pt.dv, kDLCUDA, sh.device, DLDataType{kDLFloat, 32, 1}, 2, os);
      check_c(cuvsIvfFlatSearch(sh.res, params->base_params, sh.index, &tq, &ti, &tv, cuvsFilter{0, NO_FILTER}), "cuvsIvfFlatSearch (shard)");
      pt.hi.resize(static_cast<size_t>(pt.qn) * k);
      pt.hv.resize(static_cast<size_t>(pt.qn) * k);
      B2_CUDA(cudaMemcpyAsync(pt.hi.data(), pt.di, sizeof(int64_t) * pt.qn * k, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaMemcpyAsync(pt.hv.data(), pt.dv, sizeof(float) * pt.qn * k, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaFreeAsync(pt.dq, st));
      B2_CUDA(cudaFreeAsync(pt.di, st));
      B2_CUDA(cudaFreeAsync(pt.dv, st));
    }
    for (int s = 0; s < ns; ++s) {
      device_guard g(mg.shards[s].device);
      check_c(cuvsStreamSync(mg.shards[s].res), "cuvsStreamSync");
    }
    int64_t* out_i = dl_ptr<int64_t>(nb);
    float* out_v   = dl_ptr<float>(dd);
    if (mg.mode != CUVS_NEIGHBORS_MG_SHARDED) {
      for (int s = 0; s < ns; ++s) {
        auto& pt = parts[s];
        if (!pt.qn) continue;
        std::copy(pt.hi.begin(), pt.hi.end(), out_i + pt.q0 * k);
        std::copy(pt.hv.begin(), pt.hv.end(), out_v + pt.q0 * k);
      }
      return;
    }
    // k-way merge of the sorted per-shard lists, ids translated by the shard's row offset (snmg.cuh:346-356)
    std::vector<int> cur(ns);
    for (int64_t qi = 0; qi < nq; ++qi) {
      std::fill(cur.begin(), cur.end(), 0);
      for (int j = 0; j < k; ++j) {
        int best = -1;
        float bv = 0;
        for (int s = 0; s < ns; ++s) {
          if (cur[s] >= k) continue;
          const int64_t id = parts[s].hi[qi * k + cur[s]];
          if (id == INT64_MAX || id < 0) { cur[s] = k; continue; }
          const float v = parts[s].hv[qi * k + cur[s]];
          if (best < 0 || (select_min ? v < bv : v > bv)) { best = s; bv = v; }
        }
        if (best < 0) { out_i[qi * k + j] = INT64_MAX; out_v[qi * k + j] = select_min ? FLT_MAX : -FLT_MAX; continue; }
        out_i[qi * k + j] = parts[best].hi[qi * k + cur[best]] + mg.shards[best].row0;
        out_v[qi * k + j] = bv;
        ++cur[best];
      }
    }
  });
}

cuvsError_t cuvsMultiGpuIvfFlatExtend(cuvsResources_t, cuvsMultiGpuIvfFlatIndex_t, DLManagedTensor*, DLManagedTensor*)
{
  return guarded([=] { B2_FAIL("cuvsMultiGpuIvfFlatExtend is not implemented in this build (use the per-shard cuvsIvfFlatExtend through cuvs_b200.distributed)"); });
}
cuvsError_t cuvsMultiGpuIvfFlatSerialize(cuvsResources_t, cuvsMultiGpuIvfFlatIndex_t, const char*)
{
  return guarded([=] { B2_FAIL("cuvsMultiGpuIvfFlatSerialize is not implemented in this build"); });
}
cuvsError_t cuvsMultiGpuIvfFlatDeserialize(cuvsResources_t, const char*, cuvsMultiGpuIvfFlatIndex_t)
{
  return guarded([=] { B2_FAIL("cuvsMultiGpuIvfFlatDeserialize is not implemented in this build"); });
}
cuvsError_t cuvsMultiGpuIvfFlatDistribute(cuvsResources_t, const char*, cuvsMultiGpuIvfFlatIndex_t)
{
  return guarded([=] { B2_FAIL("cuvsMultiGpuIvfFlatDistribute is not implemented in this build"); });
}

// x [m,k], y [n,k] -> dist [m,n]; exact fp32 in the oracle's arithmetic (reference: c/src/distance/pairwise_distance.cpp,
// cpp/src/distance/detail/distance.cuh:308-333).  L2 / inner product / cosine.
cuvsError_t cuvsPairwiseDistance(cuvsResources_t res, DLManagedTensor* x_t, DLManagedTensor* y_t, DLManagedTensor* dist_t,
                                 cuvsDistanceType metric, float /*metric_arg*/)
{
  return guarded([=] {
    auto r = as_res(res);
    B2_EXPECTS(x_t && y_t && dist_t, "null argument");
    const DLTensor& x = x_t->dl_tensor;
    const DLTensor& y = y_t->dl_tensor;
    const DLTensor& d = dist_t->dl_tensor;
    B2_EXPECTS(dl_is_device(x) && dl_is_device(y) && dl_is_device(d), "x, y and dist should have device compatible memory");
    B2_EXPECTS(dl_is(x, kDLFloat, 32) && dl_is(y, kDLFloat, 32) && dl_is(d, kDLFloat, 32), "pairwise_distance: float32 tensors expected");
    B2_EXPECTS(x.ndim == 2 && y.ndim == 2 && d.ndim == 2 && x.shape[1] == y.shape[1] && d.shape[0] == x.shape[0] && d.shape[1] == y.shape[0], "shape mismatch");
    B2_EXPECTS(dl_is_c_contiguous(x) && dl_is_c_contiguous(y) && dl_is_c_contiguous(d), "row-major contiguous tensors expected");
    B2_EXPECTS(metric == L2Expanded || metric == L2SqrtExpanded || metric == L2Unexpanded || metric == L2SqrtUnexpanded || metric == InnerProduct || metric == CosineExpanded,
               "pairwise_distance: metric %d is outside the scan+top-k hot path of this library", int(metric));
    const int64_t m = x.shape[0], n = y.shape[0];
    const int k = static_cast<int>(x.shape[1]);
    auto s = r->stream;
    const bool norms = metric == L2Expanded || metric == L2SqrtExpanded || metric == CosineExpanded;
    dbuf<float> xn, yn;
    if (norms) {
      xn.alloc(static_cast<size_t>(m), s); yn.alloc(static_cast<size_t>(n), s);
      row_norms(s, dl_ptr<float>(x), m, k, k, xn.data());
      row_norms(s, dl_ptr<float>(y), n, k, k, yn.data());
    }
    const int64_t chunk = 65535 * 64;
    for (int64_t r0 = 0; r0 < m; r0 += chunk) {
      int64_t rows = std::min(chunk, m - r0);
      exact_distance_tile(s, dl_ptr<float>(x) + r0 * k, rows, k, dl_ptr<float>(y), n, k, k, norms ? xn.data() + r0 : nullptr,
                          norms ? yn.data() : nullptr, metric, dl_ptr<float>(d) + r0 * n, n, filter_view{}, 0);
    }
    postprocess_distances(s, dl_ptr<float>(d), m * n, metric);
  });
}

}  // extern "C"
