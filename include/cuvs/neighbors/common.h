/* Pre-filter handle and merge-strategy enum shared by the neighbor indexes
 * (reference: c/include/cuvs/neighbors/common.h:25-41, :53-56).
 * v1 of this library searches unfiltered: NO_FILTER is accepted everywhere,
 * BITSET/BITMAP are accepted by brute_force/ivf_flat/cagra search. */
#pragma once
#include <stdint.h>
#include <cuvs/core/export.h>
#ifdef __cplusplus
extern "C" {
#endif
enum cuvsFilterType { NO_FILTER = 0, BITSET = 1, BITMAP = 2 };

typedef struct {
  uintptr_t addr; /* DLManagedTensor* of uint32 words when type != NO_FILTER */
  enum cuvsFilterType type;
} cuvsFilter;

typedef enum { MERGE_STRATEGY_PHYSICAL = 0, MERGE_STRATEGY_LOGICAL = 1 } cuvsMergeStrategy;
#ifdef __cplusplus
}
#endif
