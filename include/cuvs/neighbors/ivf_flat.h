/*
 * IVF-Flat C boundary.
 * Replaces c/include/cuvs/neighbors/ivf_flat.h: index params :29-74 (+Create/
 * Destroy :82/:90), search params :98-110 (:116/:124), index handle :137-140
 * (:150/:157), getters :166-184, Build :236, Search :293, Serialize/Deserialize
 * :329/:342, Extend :362.  Field order and types are ABI and are kept.
 */
#pragma once
#include <cuvs/core/c_api.h>
#include <cuvs/core/export.h>
#include <cuvs/distance/distance.h>
#include <cuvs/neighbors/common.h>
#include <dlpack/dlpack.h>
#include <stdbool.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

struct cuvsIvfFlatIndexParams {
  cuvsDistanceType metric;
  float metric_arg;
  bool add_data_on_build;
  uint32_t n_lists;
  uint32_t kmeans_n_iters;
  double kmeans_trainset_fraction;
  bool adaptive_centers;
  bool conservative_memory_allocation;
};
typedef struct cuvsIvfFlatIndexParams* cuvsIvfFlatIndexParams_t;
CUVS_EXPORT cuvsError_t cuvsIvfFlatIndexParamsCreate(cuvsIvfFlatIndexParams_t* index_params);
CUVS_EXPORT cuvsError_t cuvsIvfFlatIndexParamsDestroy(cuvsIvfFlatIndexParams_t index_params);

struct cuvsIvfFlatSearchParams {
  uint32_t n_probes;
};
typedef struct cuvsIvfFlatSearchParams* cuvsIvfFlatSearchParams_t;
CUVS_EXPORT cuvsError_t cuvsIvfFlatSearchParamsCreate(cuvsIvfFlatSearchParams_t* params);
CUVS_EXPORT cuvsError_t cuvsIvfFlatSearchParamsDestroy(cuvsIvfFlatSearchParams_t params);

typedef struct {
  uintptr_t addr;
  DLDataType dtype;
} cuvsIvfFlatIndex;
typedef cuvsIvfFlatIndex* cuvsIvfFlatIndex_t;
CUVS_EXPORT cuvsError_t cuvsIvfFlatIndexCreate(cuvsIvfFlatIndex_t* index);
CUVS_EXPORT cuvsError_t cuvsIvfFlatIndexDestroy(cuvsIvfFlatIndex_t index);

CUVS_EXPORT cuvsError_t cuvsIvfFlatIndexGetNLists(cuvsIvfFlatIndex_t index, int64_t* n_lists);
CUVS_EXPORT cuvsError_t cuvsIvfFlatIndexGetDim(cuvsIvfFlatIndex_t index, int64_t* dim);
/* centers: filled as a non-owning [n_lists, dim] f32 device view (shape owned by the tensor, freed by its deleter). */
CUVS_EXPORT cuvsError_t cuvsIvfFlatIndexGetCenters(cuvsIvfFlatIndex_t index,
                                                   DLManagedTensor* centers);

CUVS_EXPORT cuvsError_t cuvsIvfFlatBuild(cuvsResources_t res,
                                         cuvsIvfFlatIndexParams_t index_params,
                                         DLManagedTensor* dataset,
                                         cuvsIvfFlatIndex_t index);
CUVS_EXPORT cuvsError_t cuvsIvfFlatSearch(cuvsResources_t res,
                                          cuvsIvfFlatSearchParams_t search_params,
                                          cuvsIvfFlatIndex_t index,
                                          DLManagedTensor* queries,
                                          DLManagedTensor* neighbors,
                                          DLManagedTensor* distances,
                                          cuvsFilter filter);
CUVS_EXPORT cuvsError_t cuvsIvfFlatSerialize(cuvsResources_t res,
                                             const char* filename,
                                             cuvsIvfFlatIndex_t index);
CUVS_EXPORT cuvsError_t cuvsIvfFlatDeserialize(cuvsResources_t res,
                                               const char* filename,
                                               cuvsIvfFlatIndex_t index);
CUVS_EXPORT cuvsError_t cuvsIvfFlatExtend(cuvsResources_t res,
                                          DLManagedTensor* new_vectors,
                                          DLManagedTensor* new_indices,
                                          cuvsIvfFlatIndex_t index);
#ifdef __cplusplus
}
#endif
