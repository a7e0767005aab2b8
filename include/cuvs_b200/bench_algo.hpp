/*
 * The reference benchmark harness's algorithm interface over this library: `cuvs::bench::algo<T>` with the virtuals of
 *   cpp/bench/ann/src/common/ann_types.hpp:83-166   (algo_base, algo_gpu, algo<T>: build / set_search_param / search / save /
 *                                                    load / get_preference / set_search_dataset / copy)
 * and wrappers with the class names, parameter structs and call order of the reference's
 *   cpp/bench/ann/src/cuvs/cuvs_ivf_pq_wrapper.h:30-120, cuvs_ivf_flat_wrapper.h, cuvs_wrapper.h (brute force),
 * implemented over the C ABI (the headers under include/cuvs) through include/cuvs_b200/cuvs.hpp.  The reference's harness
 * (benchmark.hpp:300-341) drives an algo<T> as  build(dataset, nrow)  ->  set_search_param(param, filter)  ->
 * [set_search_dataset]  ->  search(queries, batch, k, neighbors, distances)  and times it on get_sync_stream().
 *
 * Left out on purpose: the JSON parameter parsing (nlohmann-json is a third-party header that is not in this image; the
 * parameter structs below are what the reference's parse_build_param / parse_search_param fill) and Google Benchmark.
 * A maintainer instantiates these classes inside the reference's `create_algo<T>` switch — nothing else changes there.
 */
#pragma once
#include <cuvs/neighbors/refine.h>
#include <cuvs_b200/cuvs.hpp>

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>

namespace cuvs::bench {

enum class Metric { kInnerProduct, kEuclidean };            // ann_types.hpp:44-47
enum class MemoryType { kHost, kHostMmap, kHostPinned, kDevice, kManaged };
struct algo_property {
  MemoryType dataset_memory_type;
  MemoryType query_memory_type;  // neighbors / distances share the queries' memory type
};

inline cuvsDistanceType parse_metric_type(Metric m) { return m == Metric::kInnerProduct ? InnerProduct : L2Expanded; }

class algo_base {
 public:
  using index_type = int64_t;
  algo_base(Metric metric, int dim) : metric_(metric), dim_(dim) {}
  virtual ~algo_base() noexcept = default;

 protected:
  Metric metric_;
  int dim_;
};

class algo_gpu {
 public:
  [[nodiscard]] virtual auto get_sync_stream() const noexcept -> cudaStream_t = 0;
  [[nodiscard]] virtual auto uses_stream() const noexcept -> bool { return true; }
  virtual ~algo_gpu() noexcept = default;
};

template <typename T>
class algo : public algo_base {
 public:
  struct search_param {
    virtual ~search_param() = default;
    [[nodiscard]] virtual auto needs_dataset() const -> bool { return false; }
  };
  algo(Metric metric, int dim) : algo_base(metric, dim) {}
  virtual void build(const T* dataset, size_t nrow) = 0;
  virtual void set_search_param(const search_param& param, const void* filter_bitset) = 0;
  virtual void search(const T* queries, int batch_size, int k, algo_base::index_type* neighbors, float* distances) const = 0;
  virtual void save(const std::string& file) const = 0;
  virtual void load(const std::string& file)       = 0;
  [[nodiscard]] virtual auto get_preference() const -> algo_property = 0;
  virtual void set_search_dataset(const T* /*dataset*/, size_t /*nrow*/) {}
  virtual auto copy() -> std::unique_ptr<algo<T>> = 0;
};

namespace detail {
/** Resources + stream shared by the copies of a wrapper (the role of configured_raft_resources). */
struct shared_handle {
  cuvs::b200::resources res;
  cudaStream_t stream = nullptr;
  shared_handle()
  {
    if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) throw std::runtime_error("cudaStreamCreate failed");
    res.set_stream(stream);
  }
  ~shared_handle() { if (stream) cudaStreamDestroy(stream); }
};
/** Device copy of a host (or device) matrix: the wrappers ask for host datasets like the reference's (get_preference). */
template <typename T>
struct device_rows {
  T* ptr = nullptr;
  size_t rows = 0;
  int dim = 0;
  void assign(const T* src, size_t nrow, int d, cudaStream_t s)
  {
    release();
    rows = nrow;
    dim  = d;
    if (cudaMalloc(reinterpret_cast<void**>(&ptr), sizeof(T) * nrow * d) != cudaSuccess) throw std::runtime_error("cudaMalloc failed");
    if (cudaMemcpyAsync(ptr, src, sizeof(T) * nrow * d, cudaMemcpyDefault, s) != cudaSuccess) throw std::runtime_error("cudaMemcpy failed");
  }
  void release() { if (ptr) { cudaFree(ptr); ptr = nullptr; } }
  ~device_rows() { release(); }
};
}  // namespace detail

/** cuvs_ivf_pq_wrapper.h:30-120.  T = float (f16 / int8 / uint8 datasets go through the C ABI's dtype switch the same way). */
template <typename T = float, typename IdxT = int64_t>
class cuvs_ivf_pq : public algo<T>, public algo_gpu {
 public:
  using search_param_base = typename algo<T>::search_param;
  struct search_param : public search_param_base {
    cuvs::neighbors::ivf_pq::search_params pq_param;
    float refine_ratio = 1.0f;
    [[nodiscard]] auto needs_dataset() const -> bool override { return refine_ratio > 1.0f; }
  };
  using build_param = cuvs::neighbors::ivf_pq::index_params;

  cuvs_ivf_pq(Metric metric, int dim, const build_param& param)
    : algo<T>(metric, dim), handle_(std::make_shared<detail::shared_handle>()), index_params_(param)
  {
    index_params_.metric = parse_metric_type(metric);
  }

  void build(const T* dataset, size_t nrow) final
  {
    auto rows = std::make_shared<detail::device_rows<T>>();
    rows->assign(dataset, nrow, this->dim_, handle_->stream);
    cuvs::b200::matrix_view<const T> view{rows->ptr, static_cast<int64_t>(nrow), this->dim_};
    index_   = std::make_shared<cuvs::neighbors::ivf_pq::index>(cuvs::neighbors::ivf_pq::build(handle_->res, index_params_, view));
    dataset_ = rows;  // kept for the refine step (set_search_dataset may replace it)
    handle_->res.sync();
  }
  void set_search_param(const search_param_base& param, const void* /*filter_bitset*/) override
  {
    const auto& sp = dynamic_cast<const search_param&>(param);
    search_params_ = sp.pq_param;
    refine_ratio_  = sp.refine_ratio;
  }
  void set_search_dataset(const T* dataset, size_t nrow) override
  {
    auto rows = std::make_shared<detail::device_rows<T>>();
    rows->assign(dataset, nrow, this->dim_, handle_->stream);
    dataset_ = rows;
  }
  /** queries / neighbors / distances in DEVICE memory (get_preference), asynchronous on get_sync_stream(). */
  void search(const T* queries, int batch_size, int k, algo_base::index_type* neighbors, float* distances) const override
  {
    cuvs::b200::matrix_view<const T> q{queries, batch_size, this->dim_};
    cuvs::b200::matrix_view<int64_t> n{neighbors, batch_size, k};
    cuvs::b200::matrix_view<float> d{distances, batch_size, k};
    if (refine_ratio_ > 1.0f && dataset_) {
      const int k0 = static_cast<int>(k * refine_ratio_);
      int64_t* cand = nullptr;
      float* cand_d = nullptr;
      if (cudaMallocAsync(reinterpret_cast<void**>(&cand), sizeof(int64_t) * batch_size * k0, handle_->stream) != cudaSuccess ||
          cudaMallocAsync(reinterpret_cast<void**>(&cand_d), sizeof(float) * batch_size * k0, handle_->stream) != cudaSuccess)
        throw std::runtime_error("cudaMallocAsync failed");
      cuvs::b200::matrix_view<int64_t> cn{cand, batch_size, k0};
      cuvs::b200::matrix_view<float> cd{cand_d, batch_size, k0};
      cuvs::neighbors::ivf_pq::search(handle_->res, search_params_, *index_, q, cn, cd);
      cuvs::b200::matrix_view<const T> ds{dataset_->ptr, static_cast<int64_t>(dataset_->rows), this->dim_};
      cuvs::b200::detail::dl2 t_ds(ds), t_q(q), t_c(cn), t_n(n), t_d(d);
      cuvs::b200::check(cuvsRefine(handle_->res.get(), t_ds.ptr(), t_q.ptr(), t_c.ptr(), index_params_.metric, t_n.ptr(), t_d.ptr()), "cuvsRefine");
      cudaFreeAsync(cand, handle_->stream);
      cudaFreeAsync(cand_d, handle_->stream);
    } else {
      cuvs::neighbors::ivf_pq::search(handle_->res, search_params_, *index_, q, n, d);
    }
  }
  [[nodiscard]] auto get_sync_stream() const noexcept -> cudaStream_t override { return handle_->stream; }
  [[nodiscard]] auto get_preference() const -> algo_property override { return algo_property{MemoryType::kHost, MemoryType::kDevice}; }
  void save(const std::string& file) const override
  {
    cuvs::b200::check(cuvsIvfPqSerialize(handle_->res.get(), file.c_str(), index_->get()), "cuvsIvfPqSerialize");
  }
  void load(const std::string& file) override
  {
    index_ = std::make_shared<cuvs::neighbors::ivf_pq::index>();
    cuvs::b200::check(cuvsIvfPqDeserialize(handle_->res.get(), file.c_str(), index_->get()), "cuvsIvfPqDeserialize");
  }
  auto copy() -> std::unique_ptr<algo<T>> override { return std::make_unique<cuvs_ivf_pq<T, IdxT>>(*this); }

 private:
  std::shared_ptr<detail::shared_handle> handle_;  // copies share resources, index and dataset (copy() is a shallow copy)
  build_param index_params_;
  cuvs::neighbors::ivf_pq::search_params search_params_{};
  std::shared_ptr<cuvs::neighbors::ivf_pq::index> index_;
  std::shared_ptr<detail::device_rows<T>> dataset_;
  float refine_ratio_ = 1.0f;
};

/** cuvs_ivf_flat_wrapper.h */
template <typename T = float, typename IdxT = int64_t>
class cuvs_ivf_flat : public algo<T>, public algo_gpu {
 public:
  using search_param_base = typename algo<T>::search_param;
  struct search_param : public search_param_base {
    cuvs::neighbors::ivf_flat::search_params ivf_flat_params;
  };
  using build_param = cuvs::neighbors::ivf_flat::index_params;

  cuvs_ivf_flat(Metric metric, int dim, const build_param& param)
    : algo<T>(metric, dim), handle_(std::make_shared<detail::shared_handle>()), index_params_(param)
  {
    index_params_.metric = parse_metric_type(metric);
  }
  void build(const T* dataset, size_t nrow) final
  {
    detail::device_rows<T> rows;
    rows.assign(dataset, nrow, this->dim_, handle_->stream);
    cuvs::b200::matrix_view<const T> view{rows.ptr, static_cast<int64_t>(nrow), this->dim_};
    index_ = std::make_shared<cuvs::neighbors::ivf_flat::index>(cuvs::neighbors::ivf_flat::build(handle_->res, index_params_, view));
    handle_->res.sync();  // the index owns its copy of the rows: the staging buffer can go
  }
  void set_search_param(const search_param_base& param, const void* /*filter_bitset*/) override
  {
    search_params_ = dynamic_cast<const search_param&>(param).ivf_flat_params;
  }
  void search(const T* queries, int batch_size, int k, algo_base::index_type* neighbors, float* distances) const override
  {
    cuvs::b200::matrix_view<const T> q{queries, batch_size, this->dim_};
    cuvs::b200::matrix_view<int64_t> n{neighbors, batch_size, k};
    cuvs::b200::matrix_view<float> d{distances, batch_size, k};
    cuvs::neighbors::ivf_flat::search(handle_->res, search_params_, *index_, q, n, d);
  }
  [[nodiscard]] auto get_sync_stream() const noexcept -> cudaStream_t override { return handle_->stream; }
  [[nodiscard]] auto get_preference() const -> algo_property override { return algo_property{MemoryType::kHost, MemoryType::kDevice}; }
  void save(const std::string& file) const override
  {
    cuvs::b200::check(cuvsIvfFlatSerialize(handle_->res.get(), file.c_str(), index_->get()), "cuvsIvfFlatSerialize");
  }
  void load(const std::string& file) override
  {
    index_ = std::make_shared<cuvs::neighbors::ivf_flat::index>();
    cuvs::b200::check(cuvsIvfFlatDeserialize(handle_->res.get(), file.c_str(), index_->get()), "cuvsIvfFlatDeserialize");
  }
  auto copy() -> std::unique_ptr<algo<T>> override { return std::make_unique<cuvs_ivf_flat<T, IdxT>>(*this); }

 private:
  std::shared_ptr<detail::shared_handle> handle_;
  build_param index_params_;
  cuvs::neighbors::ivf_flat::search_params search_params_{};
  std::shared_ptr<cuvs::neighbors::ivf_flat::index> index_;
};

/** cuvs_wrapper.h: cuvs_gpu (brute force) */
template <typename T = float>
class cuvs_brute_force : public algo<T>, public algo_gpu {
 public:
  using search_param_base = typename algo<T>::search_param;
  cuvs_brute_force(Metric metric, int dim) : algo<T>(metric, dim), handle_(std::make_shared<detail::shared_handle>()) {}
  void build(const T* dataset, size_t nrow) final
  {
    auto rows = std::make_shared<detail::device_rows<T>>();
    rows->assign(dataset, nrow, this->dim_, handle_->stream);
    cuvs::b200::matrix_view<const T> view{rows->ptr, static_cast<int64_t>(nrow), this->dim_};
    index_   = std::make_shared<cuvs::neighbors::brute_force::index>(
      cuvs::neighbors::brute_force::build(handle_->res, view, parse_metric_type(this->metric_)));
    dataset_ = rows;  // the brute-force index is a view of the rows it was built on
    handle_->res.sync();
  }
  void set_search_param(const search_param_base&, const void*) override {}
  void search(const T* queries, int batch_size, int k, algo_base::index_type* neighbors, float* distances) const override
  {
    cuvs::b200::matrix_view<const T> q{queries, batch_size, this->dim_};
    cuvs::b200::matrix_view<int64_t> n{neighbors, batch_size, k};
    cuvs::b200::matrix_view<float> d{distances, batch_size, k};
    cuvs::neighbors::brute_force::search(handle_->res, *index_, q, n, d);
  }
  [[nodiscard]] auto get_sync_stream() const noexcept -> cudaStream_t override { return handle_->stream; }
  [[nodiscard]] auto get_preference() const -> algo_property override { return algo_property{MemoryType::kHost, MemoryType::kDevice}; }
  void save(const std::string& file) const override
  {
    cuvs::b200::check(cuvsBruteForceSerialize(handle_->res.get(), file.c_str(), index_->get()), "cuvsBruteForceSerialize");
  }
  void load(const std::string& file) override
  {
    index_ = std::make_shared<cuvs::neighbors::brute_force::index>();
    cuvs::b200::check(cuvsBruteForceDeserialize(handle_->res.get(), file.c_str(), index_->get()), "cuvsBruteForceDeserialize");
  }
  auto copy() -> std::unique_ptr<algo<T>> override { return std::make_unique<cuvs_brute_force<T>>(*this); }

 private:
  std::shared_ptr<detail::shared_handle> handle_;
  std::shared_ptr<cuvs::neighbors::brute_force::index> index_;
  std::shared_ptr<detail::device_rows<T>> dataset_;
};

}  // namespace cuvs::bench
