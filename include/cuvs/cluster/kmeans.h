/*
 * k-means C boundary.
 * Replaces c/include/cuvs/cluster/kmeans.h: init enum :28-45, params :47-136,
 * params_v2 :138-205, Create/Destroy :216-246, kmeans type :255, Fit :295,
 * Fit_v2 :328, Predict :358, Predict_v2 :389, ClusterCost :411.
 * The hot-path row is the assignment step (argmin_j |x_i - c_j|^2,
 * cpp/src/cluster/detail/minClusterDistanceCompute.cu:18-165): here it is one
 * fused scan+top-1 kernel (cuvs_b200/csrc/kmeans.cu), no n x k distance matrix.
 */
#pragma once
#include <cuvs/core/c_api.h>
#include <cuvs/core/export.h>
#include <cuvs/distance/distance.h>
#include <dlpack/dlpack.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum { KMeansPlusPlus = 0, Random = 1, Array = 2 } cuvsKMeansInitMethod;

struct cuvsKMeansParams {
  cuvsDistanceType metric;
  int n_clusters;
  cuvsKMeansInitMethod init;
  int max_iter;
  double tol;
  int n_init;
  double oversampling_factor;
  int batch_samples;
  int batch_centroids;
  bool inertia_check;
  bool hierarchical; /* true = balanced hierarchical k-means */
  int hierarchical_n_iters;
  int64_t streaming_batch_size;
  int64_t init_size;
};

struct cuvsKMeansParams_v2 {
  cuvsDistanceType metric;
  int n_clusters;
  cuvsKMeansInitMethod init;
  int max_iter;
  double tol;
  int n_init;
  double oversampling_factor;
  int batch_samples;
  int batch_centroids;
  bool hierarchical;
  int hierarchical_n_iters;
  int64_t streaming_batch_size;
  int64_t init_size;
};

typedef struct cuvsKMeansParams* cuvsKMeansParams_t;
typedef struct cuvsKMeansParams_v2* cuvsKMeansParams_v2_t;

CUVS_EXPORT cuvsError_t cuvsKMeansParamsCreate(cuvsKMeansParams_t* params);
CUVS_EXPORT cuvsError_t cuvsKMeansParamsDestroy(cuvsKMeansParams_t params);
CUVS_EXPORT cuvsError_t cuvsKMeansParamsCreate_v2(cuvsKMeansParams_v2_t* params);
CUVS_EXPORT cuvsError_t cuvsKMeansParamsDestroy_v2(cuvsKMeansParams_v2_t params);

typedef enum { CUVS_KMEANS_TYPE_KMEANS = 0, CUVS_KMEANS_TYPE_KMEANS_BALANCED = 1 } cuvsKMeansType;

/* X [n, d] f32 (device, or host for the streaming path), sample_weight [n] or NULL,
 * centroids [k, d] f32 device (input when init == Array, always output). */
CUVS_EXPORT cuvsError_t cuvsKMeansFit(cuvsResources_t res,
                                      cuvsKMeansParams_t params,
                                      DLManagedTensor* X,
                                      DLManagedTensor* sample_weight,
                                      DLManagedTensor* centroids,
                                      double* inertia,
                                      int* n_iter);
CUVS_EXPORT cuvsError_t cuvsKMeansFit_v2(cuvsResources_t res,
                                         cuvsKMeansParams_v2_t params,
                                         DLManagedTensor* X,
                                         DLManagedTensor* sample_weight,
                                         DLManagedTensor* centroids,
                                         double* inertia,
                                         int* n_iter);
/* labels [n] int32 device. */
CUVS_EXPORT cuvsError_t cuvsKMeansPredict(cuvsResources_t res,
                                          cuvsKMeansParams_t params,
                                          DLManagedTensor* X,
                                          DLManagedTensor* sample_weight,
                                          DLManagedTensor* centroids,
                                          DLManagedTensor* labels,
                                          bool normalize_weight,
                                          double* inertia);
CUVS_EXPORT cuvsError_t cuvsKMeansPredict_v2(cuvsResources_t res,
                                             cuvsKMeansParams_v2_t params,
                                             DLManagedTensor* X,
                                             DLManagedTensor* sample_weight,
                                             DLManagedTensor* centroids,
                                             DLManagedTensor* labels,
                                             bool normalize_weight,
                                             double* inertia);
CUVS_EXPORT cuvsError_t cuvsKMeansClusterCost(cuvsResources_t res,
                                              DLManagedTensor* X,
                                              DLManagedTensor* centroids,
                                              double* cost);
#ifdef __cplusplus
}
#endif
