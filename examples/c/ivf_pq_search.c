/*
 * Plain-C client of the drop-in library, written against the reference's own C API only (the calls a cgo / JNI / bindgen
 * binding makes): build an IVF-PQ index over device vectors, search a batch, read the result.
 *
 *   gcc examples/c/ivf_pq_search.c -Iinclude -I/usr/local/cuda/include -Lcuvs_b200/lib -lcuvs_c \
 *       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/cuvs_b200/lib -o ivf_pq_search
 *   ./ivf_pq_search            # needs a GPU
 *   ./ivf_pq_search --version  # no GPU: only cuvsVersionGet (used by tests/test_capi_cpu.py as a link test)
 */
#include <cuda_runtime_api.h>
#include <cuvs/core/c_api.h>
#include <cuvs/neighbors/ivf_pq.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(call)                                                                  \
  do {                                                                               \
    if ((call) != CUVS_SUCCESS) {                                                    \
      const char* msg = cuvsGetLastErrorText();                                      \
      fprintf(stderr, "%s failed: %s\n", #call, msg ? msg : "(no message)");         \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

static DLManagedTensor make_tensor(void* data, int64_t* shape, int ndim, uint8_t code, uint8_t bits)
{
  DLManagedTensor t;
  memset(&t, 0, sizeof(t));
  t.dl_tensor.data        = data;
  t.dl_tensor.device.device_type = kDLCUDA;
  t.dl_tensor.device.device_id   = 0;
  t.dl_tensor.ndim        = ndim;
  t.dl_tensor.dtype.code  = code;
  t.dl_tensor.dtype.bits  = bits;
  t.dl_tensor.dtype.lanes = 1;
  t.dl_tensor.shape       = shape;
  t.dl_tensor.strides     = NULL; /* row-major */
  return t;
}

int main(int argc, char** argv)
{
  uint16_t major = 0, minor = 0, patch = 0;
  CHECK(cuvsVersionGet(&major, &minor, &patch));
  printf("libcuvs_c %u.%u.%u\n", major, minor, patch);
  if (argc > 1 && strcmp(argv[1], "--version") == 0) return 0;

  const int64_t n = 100000, dim = 64, nq = 1000, k = 10;
  float* h = (float*)malloc(sizeof(float) * n * dim);
  srand(1234);
  for (int64_t i = 0; i < n * dim; ++i) h[i] = (float)rand() / RAND_MAX;
  float *d_data, *d_queries, *d_dist;
  int64_t* d_idx;
  cudaMalloc((void**)&d_data, sizeof(float) * n * dim);
  cudaMalloc((void**)&d_queries, sizeof(float) * nq * dim);
  cudaMalloc((void**)&d_dist, sizeof(float) * nq * k);
  cudaMalloc((void**)&d_idx, sizeof(int64_t) * nq * k);
  cudaMemcpy(d_data, h, sizeof(float) * n * dim, cudaMemcpyHostToDevice);
  cudaMemcpy(d_queries, h, sizeof(float) * nq * dim, cudaMemcpyHostToDevice); /* the first rows as queries */

  cuvsResources_t res;
  CHECK(cuvsResourcesCreate(&res));
  int64_t ds_shape[2] = {n, dim}, q_shape[2] = {nq, dim}, o_shape[2] = {nq, k};
  DLManagedTensor dataset = make_tensor(d_data, ds_shape, 2, kDLFloat, 32);
  DLManagedTensor queries = make_tensor(d_queries, q_shape, 2, kDLFloat, 32);
  DLManagedTensor neighbors = make_tensor(d_idx, o_shape, 2, kDLInt, 64);
  DLManagedTensor distances = make_tensor(d_dist, o_shape, 2, kDLFloat, 32);

  cuvsIvfPqIndexParams_t ip;
  cuvsIvfPqSearchParams_t sp;
  cuvsIvfPqIndex_t index;
  CHECK(cuvsIvfPqIndexParamsCreate(&ip));
  CHECK(cuvsIvfPqSearchParamsCreate(&sp));
  CHECK(cuvsIvfPqIndexCreate(&index));
  ip->n_lists = 256;
  ip->pq_dim  = 32;
  sp->n_probes = 32;
  CHECK(cuvsIvfPqBuild(res, ip, &dataset, index));
  CHECK(cuvsIvfPqSearch(res, sp, index, &queries, &neighbors, &distances));
  CHECK(cuvsStreamSync(res)); /* search is asynchronous on the handle's stream */

  int64_t first[10];
  cudaMemcpy(first, d_idx, sizeof(first), cudaMemcpyDeviceToHost);
  printf("query 0 (= dataset row 0): nearest ids");
  for (int j = 0; j < 10; ++j) printf(" %lld", (long long)first[j]);
  printf("\n");

  CHECK(cuvsIvfPqIndexDestroy(index));
  CHECK(cuvsIvfPqSearchParamsDestroy(sp));
  CHECK(cuvsIvfPqIndexParamsDestroy(ip));
  CHECK(cuvsResourcesDestroy(res));
  cudaFree(d_data); cudaFree(d_queries); cudaFree(d_dist); cudaFree(d_idx);
  free(h);
  return 0;
}
