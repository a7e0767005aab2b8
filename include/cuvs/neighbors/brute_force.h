/*
 * Exact kNN ("brute force") C boundary.
 * Replaces c/include/cuvs/neighbors/brute_force.h: index handle :28-33,
 * Create/Destroy :41/:48, Build :94, Search :150, Serialize/Deserialize :186/:213.
 *
 * B200 path behind cuvsBruteForceSearch (cuvs_b200/csrc/brute_force.cu):
 * split-bf16 Q.D^T on tcgen05 with the |x|^2 term folded into the K extension,
 * a per-row register top-k' epilogue out of TMEM, exact fp32 re-scoring of the
 * k' candidates and a certificate check.
 */
#pragma once
#include <cuvs/core/c_api.h>
#include <cuvs/core/export.h>
#include <cuvs/distance/distance.h>
#include <cuvs/neighbors/common.h>
#include <dlpack/dlpack.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  uintptr_t addr;   /* heap object owned by the library */
  DLDataType dtype; /* dtype the index was built from */
} cuvsBruteForceIndex;
typedef cuvsBruteForceIndex* cuvsBruteForceIndex_t;

CUVS_EXPORT cuvsError_t cuvsBruteForceIndexCreate(cuvsBruteForceIndex_t* index);
CUVS_EXPORT cuvsError_t cuvsBruteForceIndexDestroy(cuvsBruteForceIndex_t index);

/* dataset: [n, dim] f32, host or device, C- or F-contiguous. */
CUVS_EXPORT cuvsError_t cuvsBruteForceBuild(cuvsResources_t res,
                                            DLManagedTensor* dataset,
                                            cuvsDistanceType metric,
                                            float metric_arg,
                                            cuvsBruteForceIndex_t index);

/* queries [nq, dim] f32 device; neighbors [nq, k] int64; distances [nq, k] f32. */
CUVS_EXPORT cuvsError_t cuvsBruteForceSearch(cuvsResources_t res,
                                             cuvsBruteForceIndex_t index,
                                             DLManagedTensor* queries,
                                             DLManagedTensor* neighbors,
                                             DLManagedTensor* distances,
                                             cuvsFilter prefilter);

CUVS_EXPORT cuvsError_t cuvsBruteForceSerialize(cuvsResources_t res,
                                                const char* filename,
                                                cuvsBruteForceIndex_t index);
CUVS_EXPORT cuvsError_t cuvsBruteForceDeserialize(cuvsResources_t res,
                                                  const char* filename,
                                                  cuvsBruteForceIndex_t index);
#ifdef __cplusplus
}
#endif
