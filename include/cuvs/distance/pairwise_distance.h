/* Dense pairwise distance C entry point
 * (reference: c/include/cuvs/distance/pairwise_distance.h:40-56).
 * x:[m,k], y:[n,k], dist:[m,n], all device f32 row-major; L2/IP/cosine only. */
#pragma once
#include <cuvs/core/c_api.h>
#include <cuvs/distance/distance.h>
#include <dlpack/dlpack.h>
#ifdef __cplusplus
extern "C" {
#endif
CUVS_EXPORT cuvsError_t cuvsPairwiseDistance(cuvsResources_t res,
                                             DLManagedTensor* x,
                                             DLManagedTensor* y,
                                             DLManagedTensor* dist,
                                             cuvsDistanceType metric,
                                             float metric_arg);
#ifdef __cplusplus
}
#endif
