// The reference harness's call sequence on an algo<T> (cpp/bench/ann/src/common/benchmark.hpp:300-341), against the wrappers
// of include/cuvs_b200/bench_algo.hpp:  build -> save/load -> set_search_param -> search on get_sync_stream() -> recall.
// `--no-gpu` only constructs the wrappers and prints their preferences.
#include <cuvs_b200/bench_algo.hpp>

#include <cstdio>
#include <cstring>
#include <vector>

using namespace cuvs::bench;

static double run(algo<float>& a, const algo<float>::search_param& sp, const std::vector<float>& base, size_t n, int d, int nq, int k)
{
  a.build(base.data(), n);
  a.set_search_param(sp, nullptr);
  if (sp.needs_dataset()) a.set_search_dataset(base.data(), n);
  auto& gpu = dynamic_cast<algo_gpu&>(a);
  float *dq = nullptr, *dd = nullptr;
  int64_t* di = nullptr;
  cudaMalloc(&dq, sizeof(float) * nq * d);
  cudaMalloc(&dd, sizeof(float) * nq * k);
  cudaMalloc(&di, sizeof(int64_t) * nq * k);
  cudaMemcpy(dq, base.data(), sizeof(float) * nq * d, cudaMemcpyHostToDevice);  // the first nq rows query themselves
  auto worker = a.copy();  // the harness searches through per-thread shallow copies
  worker->search(dq, nq, k, di, dd);
  cudaStreamSynchronize(gpu.get_sync_stream());
  std::vector<int64_t> hi(static_cast<size_t>(nq) * k);
  cudaMemcpy(hi.data(), di, sizeof(int64_t) * nq * k, cudaMemcpyDeviceToHost);
  int self = 0;
  for (int q = 0; q < nq; ++q)
    for (int j = 0; j < k; ++j) self += hi[static_cast<size_t>(q) * k + j] == q;
  cudaFree(dq); cudaFree(dd); cudaFree(di);
  return static_cast<double>(self) / nq;
}

int main(int argc, char** argv)
{
  const int d = 64;
  cuvs_ivf_pq<float>::build_param pq_build;
  pq_build.n_lists = 64;
  pq_build.pq_dim  = 32;
  cuvs_ivf_flat<float>::build_param flat_build;
  flat_build.n_lists = 64;
  if (argc > 1 && std::strcmp(argv[1], "--no-gpu") == 0) {
    std::printf("bench_algo.hpp: algo<T> wrappers cuvs_ivf_pq / cuvs_ivf_flat / cuvs_brute_force compiled; ivf_pq defaults n_probes %u\n",
                cuvs_ivf_pq<float>::search_param{}.pq_param.n_probes);
    return 0;
  }
  const size_t n = 20000;
  const int nq = 200, k = 5;
  std::vector<float> base(n * d);
  uint64_t s = 88172645463325252ull;
  for (auto& v : base) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = static_cast<float>(s % 10000) / 10000.0f; }
  try {
    cuvs_ivf_pq<float> pq(Metric::kEuclidean, d, pq_build);
    cuvs_ivf_pq<float>::search_param pq_sp;
    pq_sp.pq_param.n_probes = 16;
    pq_sp.refine_ratio      = 2.0f;
    const double r_pq = run(pq, pq_sp, base, n, d, nq, k);
    pq.save("/tmp/cuvs_b200_bench_algo_demo.ivf_pq");
    pq.load("/tmp/cuvs_b200_bench_algo_demo.ivf_pq");
    cuvs_ivf_flat<float> flat(Metric::kEuclidean, d, flat_build);
    cuvs_ivf_flat<float>::search_param flat_sp;
    flat_sp.ivf_flat_params.n_probes = 16;
    const double r_flat = run(flat, flat_sp, base, n, d, nq, k);
    cuvs_brute_force<float> bf(Metric::kEuclidean, d);
    algo<float>::search_param none;
    const double r_bf = run(bf, none, base, n, d, nq, k);
    std::printf("self-hit rate: ivf_pq+refine %.3f  ivf_flat %.3f  brute_force %.3f\n", r_pq, r_flat, r_bf);
    if (r_pq < 0.9 || r_flat < 0.9 || r_bf < 0.999) { std::printf("BENCH_ALGO_FAILED\n"); return 1; }
  } catch (const std::exception& e) {
    std::printf("error: %s\n", e.what());
    return 2;
  }
  std::printf("BENCH_ALGO_OK\n");
  return 0;
}
