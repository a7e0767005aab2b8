/*
 * IVF-PQ C boundary.
 * Replaces c/include/cuvs/neighbors/ivf_pq.h: enums :28-62, index params
 * :64-138 (+:146/:154), search params :167-221 (+:229/:237), index handle
 * :250-253 (+:263/:270), scalar getters :273-289, tensor getters :298-354,
 * list accessors :373/:388, Build :442, BuildPrecomputed :476, Search :536,
 * Serialize/Deserialize :570/:581, Extend :599, Transform :621.
 * Struct field order/types are ABI and are kept exactly.
 */
#pragma once
#include <cuvs/core/c_api.h>
#include <cuvs/core/export.h>
#include <cuvs/distance/distance.h>
#include <dlpack/dlpack.h>
#include <stdbool.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum cuvsIvfPqCodebookGen {
  CUVS_IVF_PQ_CODEBOOK_GEN_PER_SUBSPACE = 0,
  CUVS_IVF_PQ_CODEBOOK_GEN_PER_CLUSTER  = 1,
};
enum cuvsIvfPqListLayout {
  CUVS_IVF_PQ_LIST_LAYOUT_FLAT        = 0,
  CUVS_IVF_PQ_LIST_LAYOUT_INTERLEAVED = 1,
};

struct cuvsIvfPqIndexParams {
  cuvsDistanceType metric;
  float metric_arg;
  bool add_data_on_build;
  uint32_t n_lists;
  uint32_t kmeans_n_iters;
  double kmeans_trainset_fraction;
  uint32_t pq_bits; /* 4..8 */
  uint32_t pq_dim;  /* 0 = choose from dim */
  enum cuvsIvfPqCodebookGen codebook_kind;
  bool force_random_rotation;
  bool conservative_memory_allocation;
  uint32_t max_train_points_per_pq_code;
  enum cuvsIvfPqListLayout codes_layout;
};
typedef struct cuvsIvfPqIndexParams* cuvsIvfPqIndexParams_t;
CUVS_EXPORT cuvsError_t cuvsIvfPqIndexParamsCreate(cuvsIvfPqIndexParams_t* index_params);
CUVS_EXPORT cuvsError_t cuvsIvfPqIndexParamsDestroy(cuvsIvfPqIndexParams_t index_params);

struct cuvsIvfPqSearchParams {
  uint32_t n_probes;
  cudaDataType_t lut_dtype;               /* CUDA_R_32F | CUDA_R_16F | CUDA_R_8U */
  cudaDataType_t internal_distance_dtype; /* CUDA_R_32F | CUDA_R_16F */
  cudaDataType_t coarse_search_dtype;     /* CUDA_R_32F | CUDA_R_16F | CUDA_R_8I */
  uint32_t max_internal_batch_size;
  double preferred_shmem_carveout;
};
typedef struct cuvsIvfPqSearchParams* cuvsIvfPqSearchParams_t;
CUVS_EXPORT cuvsError_t cuvsIvfPqSearchParamsCreate(cuvsIvfPqSearchParams_t* params);
CUVS_EXPORT cuvsError_t cuvsIvfPqSearchParamsDestroy(cuvsIvfPqSearchParams_t params);

typedef struct {
  uintptr_t addr;
  DLDataType dtype;
} cuvsIvfPqIndex;
typedef cuvsIvfPqIndex* cuvsIvfPqIndex_t;
CUVS_EXPORT cuvsError_t cuvsIvfPqIndexCreate(cuvsIvfPqIndex_t* index);
CUVS_EXPORT cuvsError_t cuvsIvfPqIndexDestroy(cuvsIvfPqIndex_t index);

CUVS_EXPORT cuvsError_t cuvsIvfPqIndexGetNLists(cuvsIvfPqIndex_t index, int64_t* n_lists);
CUVS_EXPORT cuvsError_t cuvsIvfPqIndexGetDim(cuvsIvfPqIndex_t index, int64_t* dim);
CUVS_EXPORT cuvsError_t cuvsIvfPqIndexGetSize(cuvsIvfPqIndex_t index, int64_t* size);
CUVS_EXPORT cuvsError_t cuvsIvfPqIndexGetPqDim(cuvsIvfPqIndex_t index, int64_t* pq_dim);
CUVS_EXPORT cuvsError_t cuvsIvfPqIndexGetPqBits(cuvsIvfPqIndex_t index, int64_t* pq_bits);
CUVS_EXPORT cuvsError_t cuvsIvfPqIndexGetPqLen(cuvsIvfPqIndex_t index, int64_t* pq_len);

/* The tensor getters fill a caller-provided DLManagedTensor with a non-owning
 * view (data/shape point into the index; deleter == NULL), as the reference does. */
CUVS_EXPORT cuvsError_t cuvsIvfPqIndexGetCenters(cuvsIvfPqIndex_t index, DLManagedTensor* centers);
CUVS_EXPORT cuvsError_t cuvsIvfPqIndexGetCentersPadded(cuvsIvfPqIndex_t index,
                                                       DLManagedTensor* centers);
CUVS_EXPORT cuvsError_t cuvsIvfPqIndexGetPqCenters(cuvsIvfPqIndex_t index,
                                                   DLManagedTensor* pq_centers);
CUVS_EXPORT cuvsError_t cuvsIvfPqIndexGetCentersRot(cuvsIvfPqIndex_t index,
                                                    DLManagedTensor* centers_rot);
CUVS_EXPORT cuvsError_t cuvsIvfPqIndexGetRotationMatrix(cuvsIvfPqIndex_t index,
                                                        DLManagedTensor* rotation_matrix);
CUVS_EXPORT cuvsError_t cuvsIvfPqIndexGetListSizes(cuvsIvfPqIndex_t index,
                                                   DLManagedTensor* list_sizes);
/* out_codes: [n_take, ceil(pq_dim*pq_bits/8)] uint8 device tensor. */
CUVS_EXPORT cuvsError_t cuvsIvfPqIndexUnpackContiguousListData(cuvsResources_t res,
                                                               cuvsIvfPqIndex_t index,
                                                               DLManagedTensor* out_codes,
                                                               uint32_t label,
                                                               uint32_t offset);
CUVS_EXPORT cuvsError_t cuvsIvfPqIndexGetListIndices(cuvsIvfPqIndex_t index,
                                                     uint32_t label,
                                                     DLManagedTensor* out_labels);

CUVS_EXPORT cuvsError_t cuvsIvfPqBuild(cuvsResources_t res,
                                       cuvsIvfPqIndexParams_t params,
                                       DLManagedTensor* dataset,
                                       cuvsIvfPqIndex_t index);
/* Build an empty index around caller-supplied codebooks (device tensors):
 * pq_centers [pq_dim, pq_len, 2^pq_bits], centers [n_lists, dim] or [n_lists, dim_ext],
 * centers_rot [n_lists, rot_dim], rotation_matrix [rot_dim, dim]. */
CUVS_EXPORT cuvsError_t cuvsIvfPqBuildPrecomputed(cuvsResources_t res,
                                                  cuvsIvfPqIndexParams_t params,
                                                  uint32_t dim,
                                                  DLManagedTensor* pq_centers,
                                                  DLManagedTensor* centers,
                                                  DLManagedTensor* centers_rot,
                                                  DLManagedTensor* rotation_matrix,
                                                  cuvsIvfPqIndex_t index);
CUVS_EXPORT cuvsError_t cuvsIvfPqSearch(cuvsResources_t res,
                                        cuvsIvfPqSearchParams_t search_params,
                                        cuvsIvfPqIndex_t index,
                                        DLManagedTensor* queries,
                                        DLManagedTensor* neighbors,
                                        DLManagedTensor* distances);
CUVS_EXPORT cuvsError_t cuvsIvfPqSerialize(cuvsResources_t res,
                                           const char* filename,
                                           cuvsIvfPqIndex_t index);
CUVS_EXPORT cuvsError_t cuvsIvfPqDeserialize(cuvsResources_t res,
                                             const char* filename,
                                             cuvsIvfPqIndex_t index);
CUVS_EXPORT cuvsError_t cuvsIvfPqExtend(cuvsResources_t res,
                                        DLManagedTensor* new_vectors,
                                        DLManagedTensor* new_indices,
                                        cuvsIvfPqIndex_t index);
CUVS_EXPORT cuvsError_t cuvsIvfPqTransform(cuvsResources_t res,
                                           cuvsIvfPqIndex_t index,
                                           DLManagedTensor* input_dataset,
                                           DLManagedTensor* output_labels,
                                           DLManagedTensor* output_dataset);
#ifdef __cplusplus
}
#endif
