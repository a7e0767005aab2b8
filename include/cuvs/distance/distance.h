/* Distance selector shared by every index type.
 * Numeric values must equal the reference's enum
 * (c/include/cuvs/distance/distance.h:14-60) because bindings pass raw ints.
 * Implemented on the hot path: L2Expanded, L2SqrtExpanded, L2Unexpanded,
 * L2SqrtUnexpanded, InnerProduct, CosineExpanded; the rest return CUVS_ERROR. */
#pragma once
#include <cuvs/core/export.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef enum {
  L2Expanded          = 0,  /* |x|^2 + |y|^2 - 2 x.y  (squared) */
  L2SqrtExpanded      = 1,  /* sqrt of the above */
  CosineExpanded      = 2,  /* 1 - x.y / (|x||y|) */
  L1                  = 3,
  L2Unexpanded        = 4,  /* sum (x-y)^2 */
  L2SqrtUnexpanded    = 5,
  InnerProduct        = 6,  /* x.y, larger is closer */
  Linf                = 7,
  Canberra            = 8,
  LpUnexpanded        = 9,
  CorrelationExpanded = 10,
  JaccardExpanded     = 11,
  HellingerExpanded   = 12,
  Haversine           = 13,
  BrayCurtis          = 14,
  JensenShannon       = 15,
  HammingUnexpanded   = 16,
  KLDivergence        = 17,
  RusselRaoExpanded   = 18,
  DiceExpanded        = 19,
  BitwiseHamming      = 20,
  Precomputed         = 100
} cuvsDistanceType;
#ifdef __cplusplus
}
#endif
