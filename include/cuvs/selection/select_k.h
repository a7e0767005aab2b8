/*
 * C entry points for the selection / merge primitives of the scan+top-k path.
 *
 * The reference exposes these only in C++:
 *   cuvs::selection::select_k        cpp/include/cuvs/selection/select_k.hpp:70-198
 *                                    (implementation delegated to RAFT,
 *                                     cpp/src/selection/select_k.cuh:39-51)
 *   cuvs::neighbors::knn_merge_parts cpp/include/cuvs/neighbors/knn_merge_parts.hpp:20-46
 *                                    (cpp/src/neighbors/detail/knn_merge_parts.cuh:24-170)
 * and its bindings reach them only through the index searches.  This library
 * exports them in C as well so that the per-GPU shard merge of the sharded
 * search (cuvs_b200/distributed.py) and the parity tests can call them
 * directly.  Same argument meaning as the C++ functions.
 */
#pragma once
#include <cuvs/core/c_api.h>
#include <cuvs/core/export.h>
#include <dlpack/dlpack.h>
#include <stdbool.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* in_val [batch,len] f32; in_idx [batch,len] (int64|uint32|int32) or NULL (=> column
 * position); out_val [batch,k] f32; out_idx [batch,k] same dtype as in_idx (int64 when
 * in_idx is NULL).  Ties are broken towards the smaller column position; output rows
 * are sorted best-first when `sorted`.  Rows with fewer than k finite candidates are
 * padded with +/-FLT_MAX and the all-ones index. */
CUVS_EXPORT cuvsError_t cuvsSelectK(cuvsResources_t res,
                                    DLManagedTensor* in_val,
                                    DLManagedTensor* in_idx,
                                    DLManagedTensor* out_val,
                                    DLManagedTensor* out_idx,
                                    bool select_min,
                                    bool sorted);

/* in_keys/in_values [n_parts * n_rows, k] (part-major), out [n_rows, k];
 * translations[n_parts] (host int64 array, may be NULL) is added to the ids of each part. */
CUVS_EXPORT cuvsError_t cuvsKnnMergeParts(cuvsResources_t res,
                                          DLManagedTensor* in_keys,
                                          DLManagedTensor* in_values,
                                          DLManagedTensor* out_keys,
                                          DLManagedTensor* out_values,
                                          int64_t n_parts,
                                          const int64_t* translations,
                                          bool select_min);
#ifdef __cplusplus
}
#endif
