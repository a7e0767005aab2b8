// C++ client of the header-only adaptor include/cuvs_b200/cuvs.hpp: the reference's C++ call shape
//   auto index = cuvs::neighbors::ivf_pq::build(res, index_params, dataset_view);
//   cuvs::neighbors::ivf_pq::search(res, search_params, index, queries_view, neighbors_view, distances_view);
// (cpp/include/cuvs/neighbors/ivf_pq.hpp:1821-1828) over libcuvs_c.so.  `--no-gpu` only exercises what needs no device.
#include <cuvs_b200/cuvs.hpp>

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace pq = cuvs::neighbors::ivf_pq;
namespace bf = cuvs::neighbors::brute_force;

int main(int argc, char** argv)
{
  pq::index_params ip;
  ip.n_lists = 64;
  ip.pq_dim  = 32;
  pq::search_params sp;
  sp.n_probes = 16;
  uint16_t major = 0, minor = 0, patch = 0;
  cuvs::b200::check(cuvsVersionGet(&major, &minor, &patch), "cuvsVersionGet");
  std::printf("libcuvs_c %u.%02u.%u, ivf_pq defaults: n_lists %u pq_bits %u n_probes %u\n", major, minor, patch, pq::index_params{}.n_lists,
              pq::index_params{}.pq_bits, pq::search_params{}.n_probes);
  if (argc > 1 && std::strcmp(argv[1], "--no-gpu") == 0) return 0;

  const int64_t n = 20000, d = 64, nq = 200, k = 5;
  std::vector<float> h(n * d);
  uint64_t s = 88172645463325252ull;
  for (auto& v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = static_cast<float>(s % 10000) / 10000.0f; }
  float *dx = nullptr, *dd = nullptr;
  int64_t* di = nullptr;
  cudaMalloc(&dx, sizeof(float) * n * d);
  cudaMalloc(&dd, sizeof(float) * nq * k);
  cudaMalloc(&di, sizeof(int64_t) * nq * k);
  cudaMemcpy(dx, h.data(), sizeof(float) * n * d, cudaMemcpyHostToDevice);
  try {
    cuvs::b200::resources res;
    cuvs::b200::matrix_view<const float> dataset{dx, n, d}, queries{dx, nq, d};  // the first nq rows query themselves
    cuvs::b200::matrix_view<int64_t> neighbors{di, nq, k};
    cuvs::b200::matrix_view<float> distances{dd, nq, k};
    auto index = pq::build(res, ip, dataset);
    pq::search(res, sp, index, queries, neighbors, distances);
    res.sync();
    std::vector<int64_t> hi(nq * k);
    cudaMemcpy(hi.data(), di, sizeof(int64_t) * nq * k, cudaMemcpyDeviceToHost);
    int self = 0;
    for (int64_t q = 0; q < nq; ++q)
      for (int j = 0; j < k; ++j) self += hi[q * k + j] == q;
    std::printf("ivf_pq: %lld rows indexed, %d of %lld queries found themselves\n", (long long)index.size(), self, (long long)nq);
    auto exact = bf::build(res, dataset);
    bf::search(res, exact, queries, neighbors, distances);
    res.sync();
    cudaMemcpy(hi.data(), di, sizeof(int64_t) * nq * k, cudaMemcpyDeviceToHost);
    int self_bf = 0;
    for (int64_t q = 0; q < nq; ++q) self_bf += hi[q * k] == q;
    std::printf("brute_force: %d of %lld queries are their own nearest neighbour\n", self_bf, (long long)nq);
    if (self < nq * 9 / 10 || self_bf != nq) { std::printf("CPP_ADAPTOR_FAILED\n"); return 1; }
  } catch (const cuvs::b200::error& e) {
    std::printf("cuvs error: %s\n", e.what());
    return 2;
  }
  std::printf("CPP_ADAPTOR_OK\n");
  return 0;
}
