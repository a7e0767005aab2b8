/*
 * Single-process multi-GPU IVF-Flat C boundary.
 * Replaces c/include/cuvs/neighbors/mg_ivf_flat.h: index params :26-45 (+:45/:53),
 * search params :62-84 (+:92/:100), handle :113-116 (+:124/:132), Build :152,
 * Search :177, Extend :202, Serialize :223, Deserialize :244, Distribute :266.
 * Queries / outputs are HOST tensors, as in the reference (mg API contract).
 * The one-process-per-GPU NCCL path used by bench.py lives in
 * cuvs_b200/distributed.py and calls the per-shard C entry points plus
 * cuvsKnnMergeParts from <cuvs/selection/select_k.h>.
 */
#pragma once
#include <cuvs/core/c_api.h>
#include <cuvs/core/export.h>
#include <cuvs/neighbors/ivf_flat.h>
#include <cuvs/neighbors/mg_common.h>
#include <dlpack/dlpack.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

struct cuvsMultiGpuIvfFlatIndexParams {
  cuvsIvfFlatIndexParams_t base_params;
  cuvsMultiGpuDistributionMode mode;
};
typedef struct cuvsMultiGpuIvfFlatIndexParams* cuvsMultiGpuIvfFlatIndexParams_t;
CUVS_EXPORT cuvsError_t cuvsMultiGpuIvfFlatIndexParamsCreate(
  cuvsMultiGpuIvfFlatIndexParams_t* index_params);
CUVS_EXPORT cuvsError_t cuvsMultiGpuIvfFlatIndexParamsDestroy(
  cuvsMultiGpuIvfFlatIndexParams_t index_params);

struct cuvsMultiGpuIvfFlatSearchParams {
  cuvsIvfFlatSearchParams_t base_params;
  cuvsMultiGpuReplicatedSearchMode search_mode;
  cuvsMultiGpuShardedMergeMode merge_mode;
  int64_t n_rows_per_batch;
};
typedef struct cuvsMultiGpuIvfFlatSearchParams* cuvsMultiGpuIvfFlatSearchParams_t;
CUVS_EXPORT cuvsError_t cuvsMultiGpuIvfFlatSearchParamsCreate(
  cuvsMultiGpuIvfFlatSearchParams_t* params);
CUVS_EXPORT cuvsError_t cuvsMultiGpuIvfFlatSearchParamsDestroy(
  cuvsMultiGpuIvfFlatSearchParams_t params);

typedef struct {
  uintptr_t addr;
  DLDataType dtype;
} cuvsMultiGpuIvfFlatIndex;
typedef cuvsMultiGpuIvfFlatIndex* cuvsMultiGpuIvfFlatIndex_t;
CUVS_EXPORT cuvsError_t cuvsMultiGpuIvfFlatIndexCreate(cuvsMultiGpuIvfFlatIndex_t* index);
CUVS_EXPORT cuvsError_t cuvsMultiGpuIvfFlatIndexDestroy(cuvsMultiGpuIvfFlatIndex_t index);

CUVS_EXPORT cuvsError_t cuvsMultiGpuIvfFlatBuild(cuvsResources_t res,
                                                 cuvsMultiGpuIvfFlatIndexParams_t params,
                                                 DLManagedTensor* dataset_tensor,
                                                 cuvsMultiGpuIvfFlatIndex_t index);
CUVS_EXPORT cuvsError_t cuvsMultiGpuIvfFlatSearch(cuvsResources_t res,
                                                  cuvsMultiGpuIvfFlatSearchParams_t params,
                                                  cuvsMultiGpuIvfFlatIndex_t index,
                                                  DLManagedTensor* queries_tensor,
                                                  DLManagedTensor* neighbors_tensor,
                                                  DLManagedTensor* distances_tensor);
CUVS_EXPORT cuvsError_t cuvsMultiGpuIvfFlatExtend(cuvsResources_t res,
                                                  cuvsMultiGpuIvfFlatIndex_t index,
                                                  DLManagedTensor* new_vectors_tensor,
                                                  DLManagedTensor* new_indices_tensor);
CUVS_EXPORT cuvsError_t cuvsMultiGpuIvfFlatSerialize(cuvsResources_t res,
                                                     cuvsMultiGpuIvfFlatIndex_t index,
                                                     const char* filename);
CUVS_EXPORT cuvsError_t cuvsMultiGpuIvfFlatDeserialize(cuvsResources_t res,
                                                       const char* filename,
                                                       cuvsMultiGpuIvfFlatIndex_t index);
CUVS_EXPORT cuvsError_t cuvsMultiGpuIvfFlatDistribute(cuvsResources_t res,
                                                      const char* filename,
                                                      cuvsMultiGpuIvfFlatIndex_t index);
#ifdef __cplusplus
}
#endif
