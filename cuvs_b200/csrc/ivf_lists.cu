// List layout helpers (see ivf_lists.cuh).
#include "ivf_lists.cuh"
#include "timing.hpp"

namespace b200 {
namespace {
__global__ void count_labels_kernel(const uint32_t* __restrict__ labels, int64_t n, unsigned long long* __restrict__ counts)
{
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) atomicAdd(&counts[labels[i]], 1ull);
}
__global__ void place_rows_kernel(const uint32_t* __restrict__ labels, int64_t n, const int64_t* __restrict__ offsets,
                                  const int64_t* __restrict__ base_fill, unsigned long long* __restrict__ cursor,
                                  int64_t* __restrict__ dst)
{
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint32_t l = labels[i];
  dst[i]     = offsets[l] + base_fill[l] + static_cast<int64_t>(atomicAdd(&cursor[l], 1ull));
}
}  // namespace

void list_layout::set_sizes(cudaStream_t s, const std::vector<int64_t>& sizes)
{
  n_lists = static_cast<int64_t>(sizes.size());
  h_sizes = sizes;
  h_offsets.assign(n_lists + 1, 0);
  size = 0;
  for (int64_t l = 0; l < n_lists; ++l) {
    h_offsets[l + 1] = h_offsets[l] + (sizes[l] + 127) / 128 * 128;
    size += sizes[l];
  }
  rows_total = h_offsets[n_lists];
  d_offsets.alloc(static_cast<size_t>(n_lists + 1));
  d_sizes.alloc(static_cast<size_t>(std::max<int64_t>(n_lists, 1)));
  std::vector<uint32_t> s32(sizes.begin(), sizes.end());
  B2_CUDA(cudaMemcpyAsync(d_offsets.data(), h_offsets.data(), sizeof(int64_t) * (n_lists + 1), cudaMemcpyHostToDevice, s));
  if (n_lists) B2_CUDA(cudaMemcpyAsync(d_sizes.data(), s32.data(), sizeof(uint32_t) * n_lists, cudaMemcpyHostToDevice, s));
  B2_CUDA(cudaStreamSynchronize(s));  // host vectors go out of scope
}

std::vector<int64_t> count_labels(cudaStream_t s, const uint32_t* labels, int64_t n, int64_t n_lists)
{
  dbuf<unsigned long long> c(static_cast<size_t>(n_lists), s);
  B2_CUDA(cudaMemsetAsync(c.data(), 0, sizeof(unsigned long long) * n_lists, s));
  if (n) {
    count_launch();
    count_labels_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(labels, n, c.data());
    B2_CUDA(cudaGetLastError());
  }
  std::vector<unsigned long long> h(static_cast<size_t>(n_lists));
  B2_CUDA(cudaMemcpyAsync(h.data(), c.data(), sizeof(unsigned long long) * n_lists, cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaStreamSynchronize(s));
  return std::vector<int64_t>(h.begin(), h.end());
}

void place_rows(cudaStream_t s, const uint32_t* labels, int64_t n, const list_layout& layout,
                const std::vector<int64_t>& base_fill, int64_t* dst_rows)
{
  if (n == 0) return;
  dbuf<unsigned long long> cursor(static_cast<size_t>(layout.n_lists), s);
  dbuf<int64_t> fill(static_cast<size_t>(layout.n_lists), s);
  B2_CUDA(cudaMemsetAsync(cursor.data(), 0, sizeof(unsigned long long) * layout.n_lists, s));
  B2_CUDA(cudaMemcpyAsync(fill.data(), base_fill.data(), sizeof(int64_t) * layout.n_lists, cudaMemcpyHostToDevice, s));
  count_launch();
  place_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(labels, n, layout.d_offsets.data(), fill.data(),
                                                                             cursor.data(), dst_rows);
  B2_CUDA(cudaGetLastError());
  B2_CUDA(cudaStreamSynchronize(s));  // base_fill is a host vector owned by the caller
}

}  // namespace b200
