/*
 * Header-only C++ surface over the C ABI: cuvs::neighbors::{brute_force, ivf_flat, ivf_pq, cagra}::{build, search} with the
 * argument order and parameter structs of the reference's C++ API
 *   cpp/include/cuvs/neighbors/brute_force.hpp:195-445, ivf_flat.hpp, ivf_pq.hpp:47-240 + :1821-1828, cagra.hpp:1552-1559
 * so that a C++ caller of the reference can be pointed at libcuvs_c.so without going through DLPack by hand.
 *
 * The reference's signatures take RAFT types (raft::resources, raft::device_matrix_view).  RAFT is a third-party
 * dependency that is not part of the reference tree, so this header does not include it; instead every function is a
 * template over the VIEW type and only uses the two members of std::mdspan / raft::mdspan it needs
 *     view.data_handle()      pointer to the first element (row-major, contiguous)
 *     view.extent(i)          size of dimension i
 * — a raft::device_matrix_view<T, int64_t, row_major> satisfies that as is; cuvs::b200::matrix_view below is a minimal
 * stand-in for code that has no mdspan.  `resources` wraps a cuvsResources_t (use .set_stream() to adopt the stream of a
 * raft::resources).  Errors become exceptions (cuvs::b200::error, text = cuvsGetLastErrorText()).
 */
#pragma once
#include <cuvs/core/c_api.h>
#include <cuvs/neighbors/brute_force.h>
#include <cuvs/neighbors/cagra.h>
#include <cuvs/neighbors/ivf_flat.h>
#include <cuvs/neighbors/ivf_pq.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>

namespace cuvs {
namespace b200 {

struct error : std::runtime_error {
  using std::runtime_error::runtime_error;
};
inline void check(cuvsError_t e, const char* what)
{
  if (e != CUVS_SUCCESS) {
    const char* t = cuvsGetLastErrorText();
    throw error(std::string(what) + ": " + (t ? t : "unknown error"));
  }
}

/** Minimal row-major matrix view (pointer + extents) for callers without mdspan. */
template <typename T>
struct matrix_view {
  T* ptr;
  int64_t rows, cols;
  T* data_handle() const { return ptr; }
  int64_t extent(int i) const { return i == 0 ? rows : cols; }
};

/** RAII cuvsResources_t (the role raft::resources plays in the reference's signatures). */
class resources {
 public:
  resources() { check(cuvsResourcesCreate(&res_), "cuvsResourcesCreate"); }
  ~resources() { if (res_) cuvsResourcesDestroy(res_); }
  resources(const resources&)            = delete;
  resources& operator=(const resources&) = delete;
  void set_stream(cudaStream_t s) { check(cuvsStreamSet(res_, s), "cuvsStreamSet"); }
  void sync() const { check(cuvsStreamSync(res_), "cuvsStreamSync"); }
  cuvsResources_t get() const { return res_; }

 private:
  cuvsResources_t res_ = 0;
};

namespace detail {
template <typename T>
constexpr DLDataType dl_type()
{
  using U = std::remove_cv_t<T>;
  if (std::is_same_v<U, float>) return DLDataType{kDLFloat, 32, 1};
  if (std::is_same_v<U, int64_t>) return DLDataType{kDLInt, 64, 1};
  if (std::is_same_v<U, uint32_t>) return DLDataType{kDLUInt, 32, 1};
  if (std::is_same_v<U, int32_t>) return DLDataType{kDLInt, 32, 1};
  if (std::is_same_v<U, uint8_t>) return DLDataType{kDLUInt, 8, 1};
  return DLDataType{kDLOpaqueHandle, 0, 0};
}
/** A DLManagedTensor describing a 2-D device view (no ownership). */
struct dl2 {
  DLManagedTensor m{};
  int64_t shape[2];
  template <typename View>
  explicit dl2(const View& v, DLDeviceType where = kDLCUDA)
  {
    using T  = std::remove_pointer_t<decltype(v.data_handle())>;
    shape[0] = static_cast<int64_t>(v.extent(0));
    shape[1] = static_cast<int64_t>(v.extent(1));
    m.dl_tensor.data    = const_cast<void*>(static_cast<const void*>(v.data_handle()));
    m.dl_tensor.device  = DLDevice{where, 0};
    m.dl_tensor.ndim    = 2;
    m.dl_tensor.dtype   = dl_type<T>();
    m.dl_tensor.shape   = shape;
    m.dl_tensor.strides = nullptr;
  }
  dl2(const dl2&) = delete;
  DLManagedTensor* ptr() { return &m; }
};
}  // namespace detail
}  // namespace b200

namespace neighbors {

namespace brute_force {
/** Owning handle of a brute-force index (cuvs::neighbors::brute_force::index<float, float>). */
class index {
 public:
  index() { b200::check(cuvsBruteForceIndexCreate(&h_), "cuvsBruteForceIndexCreate"); }
  ~index() { if (h_) cuvsBruteForceIndexDestroy(h_); }
  index(index&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
  index(const index&) = delete;
  cuvsBruteForceIndex_t get() const { return h_; }

 private:
  cuvsBruteForceIndex_t h_ = nullptr;
};
/** brute_force.hpp:195: build(handle, index_params{metric, metric_arg}, dataset) */
template <typename DatasetView>
index build(const b200::resources& res, const DatasetView& dataset, cuvsDistanceType metric = L2Expanded, float metric_arg = 2.0f)
{
  index idx;
  b200::detail::dl2 d(dataset);
  b200::check(cuvsBruteForceBuild(res.get(), d.ptr(), metric, metric_arg, idx.get()), "cuvsBruteForceBuild");
  return idx;
}
/** brute_force.hpp:371: search(handle, search_params, index, queries, neighbors [n_queries, k] int64, distances) */
template <typename QueryView, typename NeighborView, typename DistanceView>
void search(const b200::resources& res, const index& idx, const QueryView& queries, const NeighborView& neighbors, const DistanceView& distances)
{
  b200::detail::dl2 q(queries), n(neighbors), d(distances);
  b200::check(cuvsBruteForceSearch(res.get(), idx.get(), q.ptr(), n.ptr(), d.ptr(), cuvsFilter{0, NO_FILTER}), "cuvsBruteForceSearch");
}
}  // namespace brute_force

namespace ivf_flat {
/** ivf_flat.hpp index_params / search_params: same field names and defaults as the reference. */
struct index_params {
  cuvsDistanceType metric         = L2Expanded;
  float metric_arg                = 2.0f;
  bool add_data_on_build          = true;
  uint32_t n_lists                = 1024;
  uint32_t kmeans_n_iters         = 20;
  double kmeans_trainset_fraction = 0.5;
  bool adaptive_centers           = false;
  bool conservative_memory_allocation = false;
};
struct search_params {
  uint32_t n_probes = 20;
};
class index {
 public:
  index() { b200::check(cuvsIvfFlatIndexCreate(&h_), "cuvsIvfFlatIndexCreate"); }
  ~index() { if (h_) cuvsIvfFlatIndexDestroy(h_); }
  index(index&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
  index(const index&) = delete;
  cuvsIvfFlatIndex_t get() const { return h_; }

 private:
  cuvsIvfFlatIndex_t h_ = nullptr;
};
template <typename DatasetView>
index build(const b200::resources& res, const index_params& p, const DatasetView& dataset)
{
  cuvsIvfFlatIndexParams_t cp;
  b200::check(cuvsIvfFlatIndexParamsCreate(&cp), "cuvsIvfFlatIndexParamsCreate");
  cp->metric = p.metric; cp->metric_arg = p.metric_arg; cp->add_data_on_build = p.add_data_on_build; cp->n_lists = p.n_lists;
  cp->kmeans_n_iters = p.kmeans_n_iters; cp->kmeans_trainset_fraction = p.kmeans_trainset_fraction;
  cp->adaptive_centers = p.adaptive_centers; cp->conservative_memory_allocation = p.conservative_memory_allocation;
  index idx;
  b200::detail::dl2 d(dataset);
  const cuvsError_t e = cuvsIvfFlatBuild(res.get(), cp, d.ptr(), idx.get());
  cuvsIvfFlatIndexParamsDestroy(cp);
  b200::check(e, "cuvsIvfFlatBuild");
  return idx;
}
template <typename QueryView, typename NeighborView, typename DistanceView>
void search(const b200::resources& res, const search_params& p, const index& idx, const QueryView& queries, const NeighborView& neighbors,
            const DistanceView& distances)
{
  cuvsIvfFlatSearchParams sp{p.n_probes};
  b200::detail::dl2 q(queries), n(neighbors), d(distances);
  b200::check(cuvsIvfFlatSearch(res.get(), &sp, idx.get(), q.ptr(), n.ptr(), d.ptr(), cuvsFilter{0, NO_FILTER}), "cuvsIvfFlatSearch");
}
}  // namespace ivf_flat

namespace ivf_pq {
/** ivf_pq.hpp:47-158 index_params, :160-240 search_params: same field names and defaults as the reference. */
struct index_params {
  cuvsDistanceType metric         = L2Expanded;
  float metric_arg                = 2.0f;
  bool add_data_on_build          = true;
  uint32_t n_lists                = 1024;
  uint32_t kmeans_n_iters         = 20;
  double kmeans_trainset_fraction = 0.5;
  uint32_t pq_bits                = 8;
  uint32_t pq_dim                 = 0;
  cuvsIvfPqCodebookGen codebook_kind = CUVS_IVF_PQ_CODEBOOK_GEN_PER_SUBSPACE;
  bool force_random_rotation      = false;
  bool conservative_memory_allocation = false;
  uint32_t max_train_points_per_pq_code = 256;
};
struct search_params {
  uint32_t n_probes                      = 20;
  cudaDataType_t lut_dtype               = CUDA_R_32F;
  cudaDataType_t internal_distance_dtype = CUDA_R_32F;
  cudaDataType_t coarse_search_dtype     = CUDA_R_32F;
  uint32_t max_internal_batch_size       = 4096;
  double preferred_shmem_carveout        = 1.0;
};
class index {
 public:
  index() { b200::check(cuvsIvfPqIndexCreate(&h_), "cuvsIvfPqIndexCreate"); }
  ~index() { if (h_) cuvsIvfPqIndexDestroy(h_); }
  index(index&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
  index(const index&) = delete;
  cuvsIvfPqIndex_t get() const { return h_; }
  int64_t size() const { int64_t v = 0; b200::check(cuvsIvfPqIndexGetSize(h_, &v), "cuvsIvfPqIndexGetSize"); return v; }

 private:
  cuvsIvfPqIndex_t h_ = nullptr;
};
template <typename DatasetView>
index build(const b200::resources& res, const index_params& p, const DatasetView& dataset)
{
  cuvsIvfPqIndexParams_t cp;
  b200::check(cuvsIvfPqIndexParamsCreate(&cp), "cuvsIvfPqIndexParamsCreate");
  cp->metric = p.metric; cp->metric_arg = p.metric_arg; cp->add_data_on_build = p.add_data_on_build; cp->n_lists = p.n_lists;
  cp->kmeans_n_iters = p.kmeans_n_iters; cp->kmeans_trainset_fraction = p.kmeans_trainset_fraction; cp->pq_bits = p.pq_bits;
  cp->pq_dim = p.pq_dim; cp->codebook_kind = p.codebook_kind; cp->force_random_rotation = p.force_random_rotation;
  cp->conservative_memory_allocation = p.conservative_memory_allocation; cp->max_train_points_per_pq_code = p.max_train_points_per_pq_code;
  index idx;
  b200::detail::dl2 d(dataset);
  const cuvsError_t e = cuvsIvfPqBuild(res.get(), cp, d.ptr(), idx.get());
  cuvsIvfPqIndexParamsDestroy(cp);
  b200::check(e, "cuvsIvfPqBuild");
  return idx;
}
/** ivf_pq.hpp:1821-1828 */
template <typename QueryView, typename NeighborView, typename DistanceView>
void search(const b200::resources& res, const search_params& p, const index& idx, const QueryView& queries, const NeighborView& neighbors,
            const DistanceView& distances)
{
  cuvsIvfPqSearchParams sp{p.n_probes, p.lut_dtype, p.internal_distance_dtype, p.coarse_search_dtype, p.max_internal_batch_size,
                           p.preferred_shmem_carveout};
  b200::detail::dl2 q(queries), n(neighbors), d(distances);
  b200::check(cuvsIvfPqSearch(res.get(), &sp, idx.get(), q.ptr(), n.ptr(), d.ptr()), "cuvsIvfPqSearch");
}
}  // namespace ivf_pq

namespace cagra {
/** cagra.hpp search_params (the fields the search hot path reads; same names and defaults as the reference). */
struct search_params {
  size_t max_queries    = 0;
  size_t itopk_size     = 64;
  size_t max_iterations = 0;
  cuvsCagraSearchAlgo algo = AUTO;
  size_t team_size      = 0;
  size_t search_width   = 1;
  size_t min_iterations = 0;
  size_t thread_block_size = 0;
  cuvsCagraHashMode hashmap_mode = AUTO_HASH;
  size_t hashmap_min_bitlen   = 0;
  float hashmap_max_fill_rate = 0.5f;
  uint32_t num_random_samplings = 1;
  uint64_t rand_xor_mask = 0x128394;
};
class index {
 public:
  index() { b200::check(cuvsCagraIndexCreate(&h_), "cuvsCagraIndexCreate"); }
  ~index() { if (h_) cuvsCagraIndexDestroy(h_); }
  index(index&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
  index(const index&) = delete;
  cuvsCagraIndex_t get() const { return h_; }

 private:
  cuvsCagraIndex_t h_ = nullptr;
};
/** cagra::index(res, metric, dataset, knn_graph) (cagra.hpp:560-640): attach an existing graph. */
template <typename DatasetView, typename GraphView>
index from_graph(const b200::resources& res, cuvsDistanceType metric, const DatasetView& dataset, const GraphView& graph)
{
  index idx;
  b200::detail::dl2 d(dataset), g(graph);
  b200::check(cuvsCagraIndexFromArgs(res.get(), metric, g.ptr(), d.ptr(), idx.get()), "cuvsCagraIndexFromArgs");
  return idx;
}
/** cagra.hpp:1552-1559 (neighbors uint32 or int64) */
template <typename QueryView, typename NeighborView, typename DistanceView>
void search(const b200::resources& res, const search_params& p, const index& idx, const QueryView& queries, const NeighborView& neighbors,
            const DistanceView& distances)
{
  cuvsCagraSearchParams_t sp;
  b200::check(cuvsCagraSearchParamsCreate(&sp), "cuvsCagraSearchParamsCreate");
  sp->max_queries = p.max_queries; sp->itopk_size = p.itopk_size; sp->max_iterations = p.max_iterations; sp->algo = p.algo;
  sp->team_size = p.team_size; sp->search_width = p.search_width; sp->min_iterations = p.min_iterations;
  sp->thread_block_size = p.thread_block_size; sp->hashmap_mode = p.hashmap_mode; sp->hashmap_min_bitlen = p.hashmap_min_bitlen;
  sp->hashmap_max_fill_rate = p.hashmap_max_fill_rate; sp->num_random_samplings = p.num_random_samplings; sp->rand_xor_mask = p.rand_xor_mask;
  b200::detail::dl2 q(queries), n(neighbors), d(distances);
  const cuvsError_t e = cuvsCagraSearch(res.get(), sp, idx.get(), q.ptr(), n.ptr(), d.ptr(), cuvsFilter{0, NO_FILTER});
  cuvsCagraSearchParamsDestroy(sp);
  b200::check(e, "cuvsCagraSearch");
}
}  // namespace cagra

}  // namespace neighbors
}  // namespace cuvs
