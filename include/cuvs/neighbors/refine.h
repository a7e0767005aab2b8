/* Exact re-ranking of ANN candidates
 * (reference: c/include/cuvs/neighbors/refine.h:42; device path
 * cpp/src/neighbors/refine/refine_device.cuh:30-130).
 * dataset [n,dim] f32, queries [nq,dim] f32, candidates [nq,n_cand] int64,
 * indices [nq,k] int64, distances [nq,k] f32 — all on the device. */
#pragma once
#include <cuvs/core/c_api.h>
#include <cuvs/core/export.h>
#include <cuvs/distance/distance.h>
#include <dlpack/dlpack.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
CUVS_EXPORT cuvsError_t cuvsRefine(cuvsResources_t res,
                                   DLManagedTensor* dataset,
                                   DLManagedTensor* queries,
                                   DLManagedTensor* candidates,
                                   cuvsDistanceType metric,
                                   DLManagedTensor* indices,
                                   DLManagedTensor* distances);
#ifdef __cplusplus
}
#endif
