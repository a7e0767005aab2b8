/* Multi-GPU mode enums (reference: c/include/cuvs/neighbors/mg_common.h:20-50). */
#pragma once
#include <stdint.h>
#include <cuvs/core/export.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef enum {
  CUVS_NEIGHBORS_MG_REPLICATED = 0,
  CUVS_NEIGHBORS_MG_SHARDED    = 1
} cuvsMultiGpuDistributionMode;
typedef enum {
  CUVS_NEIGHBORS_MG_LOAD_BALANCER = 0,
  CUVS_NEIGHBORS_MG_ROUND_ROBIN   = 1
} cuvsMultiGpuReplicatedSearchMode;
typedef enum {
  CUVS_NEIGHBORS_MG_MERGE_ON_ROOT_RANK = 0,
  CUVS_NEIGHBORS_MG_TREE_MERGE         = 1
} cuvsMultiGpuShardedMergeMode;
#ifdef __cplusplus
}
#endif
