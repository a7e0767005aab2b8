/*
 * Harness for the reference's own C test drivers (TEST INFRASTRUCTURE, like the rest of oracle/).
 *
 * The drivers  c/tests/neighbors/run_brute_force_c.c, run_ivf_flat_c.c, run_ivf_pq_c.c  are compiled UNCHANGED from
 * /root/reference against this repo's include/ and linked with cuvs_b200/lib/libcuvs_c.so (oracle/ref_c_tests/Makefile;
 * objects and the binary land in oracle/_ref/, which is git-ignored: reference sources are never copied).  Upstream wraps
 * them in gtest/RAFT fixtures (brute_force_c.cu:395-433, ann_ivf_flat_c.cu:86-131, ann_ivf_pq_c.cu:86-131); RAFT is not
 * available here, so this file plays the fixture's part in plain C: same shapes (8096 x 32, 128 queries, k = 8,
 * n_lists 1024, n_probes 20), same input distribution (uniform(0.1, 2.0)), same acceptance rule — eval_neighbours
 * (id match OR distance within eps of a ground-truth distance, cpp/tests/neighbors/ann_utils.cuh:257-289) against an
 * exact host kNN, min_recall 0.95 for brute force and n_probes / n_lists for the IVF indexes.
 */
#include <cuda_runtime.h>
#include <cuvs/core/c_api.h>
#include <cuvs/neighbors/brute_force.h>
#include <cuvs/neighbors/ivf_flat.h>
#include <cuvs/neighbors/ivf_pq.h>

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void run_brute_force(int64_t n_rows, int64_t n_queries, int64_t n_dim, uint32_t n_neighbors, float* index_data, float* query_data,
                     uint32_t* prefilter_data, enum cuvsFilterType prefilter_type, float* distances_data, int64_t* neighbors_data,
                     cuvsDistanceType metric);
void run_ivf_flat(int64_t n_rows, int64_t n_queries, int64_t n_dim, uint32_t n_neighbors, float* index_data, float* query_data,
                  float* distances_data, int64_t* neighbors_data, cuvsDistanceType metric, size_t n_probes, size_t n_lists);
void run_ivf_pq(int64_t n_rows, int64_t n_queries, int64_t n_dim, uint32_t n_neighbors, float* index_data, float* query_data,
                float* distances_data, int64_t* neighbors_data, cuvsDistanceType metric, size_t n_probes, size_t n_lists);

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } \
  } while (0)

static uint64_t rng_state = 1234ULL;
static float uniform01(void)
{
  rng_state = rng_state * 6364136223846793005ULL + 1442695040888963407ULL;
  return (float)((rng_state >> 40) & 0xffffff) / 16777216.0f;
}
static void fill_uniform(float* p, size_t n, float lo, float hi)
{
  for (size_t i = 0; i < n; ++i) p[i] = lo + (hi - lo) * uniform01();
}

/* exact kNN on the host (squared L2), optional bitset / bitmap filter */
static void host_knn(const float* x, int64_t n, const float* q, int64_t nq, int64_t d, int k, const uint32_t* filter, int bitmap,
                     float* od, int64_t* oi)
{
  for (int64_t a = 0; a < nq; ++a) {
    for (int j = 0; j < k; ++j) { od[a * k + j] = INFINITY; oi[a * k + j] = -1; }
    for (int64_t r = 0; r < n; ++r) {
      if (filter) {
        int64_t bit = bitmap ? a * n + r : r;
        if (!((filter[bit >> 5] >> (bit & 31)) & 1u)) continue;
      }
      float s = 0.f;
      for (int64_t c = 0; c < d; ++c) { float df = q[a * d + c] - x[r * d + c]; s += df * df; }
      if (s < od[a * k + k - 1]) {
        int j = k - 1;
        while (j > 0 && od[a * k + j - 1] > s) { od[a * k + j] = od[a * k + j - 1]; oi[a * k + j] = oi[a * k + j - 1]; --j; }
        od[a * k + j] = s;
        oi[a * k + j] = r;
      }
    }
  }
}

/* eval_neighbours: a returned slot counts when its id is in the true set or its distance matches a true distance within eps */
static double eval_neighbours(const int64_t* ti, const float* td, const int64_t* fi, const float* fd, int64_t nq, int k, double eps)
{
  int64_t hit = 0;
  for (int64_t a = 0; a < nq; ++a)
    for (int j = 0; j < k; ++j) {
      int ok = 0;
      for (int t = 0; t < k && !ok; ++t)
        ok = (fi[a * k + j] == ti[a * k + t]) || (fabs((double)fd[a * k + j] - (double)td[a * k + t]) <= eps * fmax(1.0, fabs((double)td[a * k + t])));
      hit += ok;
    }
  return (double)hit / (double)(nq * k);
}

int main(void)
{
  const int64_t n_rows = 8096, n_queries = 128, n_dim = 32;
  const uint32_t k = 8;
  const size_t n_probes = 20, n_lists = 1024;
  float* hx  = (float*)malloc(sizeof(float) * n_rows * n_dim);
  float* hq  = (float*)malloc(sizeof(float) * n_queries * n_dim);
  float* td  = (float*)malloc(sizeof(float) * n_queries * k);
  int64_t* ti = (int64_t*)malloc(sizeof(int64_t) * n_queries * k);
  float* fd  = (float*)malloc(sizeof(float) * n_queries * k);
  int64_t* fi = (int64_t*)malloc(sizeof(int64_t) * n_queries * k);
  fill_uniform(hx, (size_t)(n_rows * n_dim), 0.1f, 2.0f);
  fill_uniform(hq, (size_t)(n_queries * n_dim), 0.1f, 2.0f);
  float *dx, *dq, *dd;
  int64_t* di;
  CK(cudaMalloc((void**)&dx, sizeof(float) * n_rows * n_dim));
  CK(cudaMalloc((void**)&dq, sizeof(float) * n_queries * n_dim));
  CK(cudaMalloc((void**)&dd, sizeof(float) * n_queries * k));
  CK(cudaMalloc((void**)&di, sizeof(int64_t) * n_queries * k));
  CK(cudaMemcpy(dx, hx, sizeof(float) * n_rows * n_dim, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dq, hq, sizeof(float) * n_queries * n_dim, cudaMemcpyHostToDevice));
  int failures = 0;

  /* ---- brute force, no filter (brute_force_c.cu:395-433) */
  host_knn(hx, n_rows, hq, n_queries, n_dim, (int)k, NULL, 0, td, ti);
  run_brute_force(n_rows, n_queries, n_dim, k, dx, dq, NULL, NO_FILTER, dd, di, L2Expanded);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(fd, dd, sizeof(float) * n_queries * k, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(fi, di, sizeof(int64_t) * n_queries * k, cudaMemcpyDeviceToHost));
  double r = eval_neighbours(ti, td, fi, fd, n_queries, (int)k, 1e-3);
  printf("run_brute_force           recall %.4f (min 0.95)\n", r);
  failures += r < 0.95;

  /* ---- brute force with a bitset and a bitmap filter, sparsity 0.2 (brute_force_c.cu:435-496) */
  for (int bitmap = 0; bitmap <= 1; ++bitmap) {
    const int64_t n_bits = bitmap ? n_queries * n_rows : n_rows;
    const int64_t n_words = (n_bits + 31) / 32;
    uint32_t* hf = (uint32_t*)calloc((size_t)n_words, sizeof(uint32_t));
    for (int64_t b = 0; b < n_bits; ++b)
      if (uniform01() < 0.2f) hf[b >> 5] |= 1u << (b & 31);
    uint32_t* df;
    CK(cudaMalloc((void**)&df, sizeof(uint32_t) * n_words));
    CK(cudaMemcpy(df, hf, sizeof(uint32_t) * n_words, cudaMemcpyHostToDevice));
    host_knn(hx, n_rows, hq, n_queries, n_dim, (int)k, hf, bitmap, td, ti);
    run_brute_force(n_rows, n_queries, n_dim, k, dx, dq, df, bitmap ? BITMAP : BITSET, dd, di, L2Expanded);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(fd, dd, sizeof(float) * n_queries * k, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(fi, di, sizeof(int64_t) * n_queries * k, cudaMemcpyDeviceToHost));
    r = eval_neighbours(ti, td, fi, fd, n_queries, (int)k, 1e-3);
    printf("run_brute_force (%s)  recall %.4f (min 0.95)\n", bitmap ? "bitmap" : "bitset", r);
    failures += r < 0.95;
    CK(cudaFree(df));
    free(hf);
  }

  /* ---- IVF-Flat and IVF-PQ (ann_ivf_flat_c.cu:86-131, ann_ivf_pq_c.cu:86-131): min_recall = n_probes / n_lists */
  host_knn(hx, n_rows, hq, n_queries, n_dim, (int)k, NULL, 0, td, ti);
  const double min_recall = (double)n_probes / (double)n_lists;
  run_ivf_flat(n_rows, n_queries, n_dim, k, dx, dq, dd, di, L2Expanded, n_probes, n_lists);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(fd, dd, sizeof(float) * n_queries * k, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(fi, di, sizeof(int64_t) * n_queries * k, cudaMemcpyDeviceToHost));
  r = eval_neighbours(ti, td, fi, fd, n_queries, (int)k, 1e-3);
  printf("run_ivf_flat              recall %.4f (min %.4f)\n", r, min_recall);
  failures += r < min_recall;

  run_ivf_pq(n_rows, n_queries, n_dim, k, dx, dq, dd, di, L2Expanded, n_probes, n_lists);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(fd, dd, sizeof(float) * n_queries * k, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(fi, di, sizeof(int64_t) * n_queries * k, cudaMemcpyDeviceToHost));
  r = eval_neighbours(ti, td, fi, fd, n_queries, (int)k, 1e-3);
  printf("run_ivf_pq                recall %.4f (min %.4f)\n", r, min_recall);
  failures += r < min_recall;

  const char* err = cuvsGetLastErrorText();
  if (err != NULL) { printf("last cuvs error text: %s\n", err); failures += 1; }
  printf(failures ? "REFERENCE C DRIVERS: %d FAILED\n" : "REFERENCE C DRIVERS: ALL PASSED\n", failures);
  return failures ? 1 : 0;
}
