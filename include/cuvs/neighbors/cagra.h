/*
 * CAGRA C boundary.
 * Replaces c/include/cuvs/neighbors/cagra.h: build enums :34-69, compression
 * params :76-110, ivf-pq graph-build params :113-121, ACE params :127-186,
 * index params :193-229, Create/Destroy family :237-339, search enums/params
 * :347-441 (+:449/:457), index handle :473-476 (+:486/:493), getters :502-562,
 * Build :617, Extend :646, Search :709, Serialize :747, SerializeToHnswlib :776,
 * Deserialize :790, IndexFromArgs :826, Merge :893.
 *
 * The hot path of this library is cuvsCagraSearch (single-CTA and multi-CTA
 * graph walk, cuvs_b200/csrc/cagra.cu) over an index supplied through
 * cuvsCagraIndexFromArgs / cuvsCagraDeserialize; cuvsCagraBuild constructs a
 * kNN graph with the library's own exact/IVF scan and prunes it.
 */
#pragma once
#include <cuvs/core/c_api.h>
#include <cuvs/core/export.h>
#include <cuvs/distance/distance.h>
#include <cuvs/neighbors/common.h>
#include <cuvs/neighbors/ivf_pq.h>
#include <dlpack/dlpack.h>
#include <stdbool.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum cuvsCagraGraphBuildAlgo {
  AUTO_SELECT            = 0,
  IVF_PQ                 = 1,
  NN_DESCENT             = 2,
  ITERATIVE_CAGRA_SEARCH = 3,
  ACE                    = 4
};
enum cuvsCagraHnswHeuristicType {
  CUVS_CAGRA_HEURISTIC_SIMILAR_SEARCH_PERFORMANCE = 0,
  CUVS_CAGRA_HEURISTIC_SAME_GRAPH_FOOTPRINT       = 1
};

struct cuvsCagraCompressionParams {
  uint32_t pq_bits;
  uint32_t pq_dim;
  uint32_t vq_n_centers;
  uint32_t kmeans_n_iters;
  double vq_kmeans_trainset_fraction;
  double pq_kmeans_trainset_fraction;
};
typedef struct cuvsCagraCompressionParams* cuvsCagraCompressionParams_t;

struct cuvsIvfPqParams {
  cuvsIvfPqIndexParams_t ivf_pq_build_params;
  cuvsIvfPqSearchParams_t ivf_pq_search_params;
  float refinement_rate;
};
typedef struct cuvsIvfPqParams* cuvsIvfPqParams_t;

struct cuvsAceParams {
  size_t npartitions;
  size_t ef_construction;
  const char* build_dir;
  bool use_disk;
  double max_host_memory_gb;
  double max_gpu_memory_gb;
};
typedef struct cuvsAceParams* cuvsAceParams_t;

struct cuvsCagraIndexParams {
  cuvsDistanceType metric;
  size_t intermediate_graph_degree;
  size_t graph_degree;
  enum cuvsCagraGraphBuildAlgo build_algo;
  size_t nn_descent_niter;
  cuvsCagraCompressionParams_t compression;
  void* graph_build_params; /* cuvsIvfPqParams_t or cuvsAceParams_t, by build_algo */
};
typedef struct cuvsCagraIndexParams* cuvsCagraIndexParams_t;

CUVS_EXPORT cuvsError_t cuvsCagraIndexParamsCreate(cuvsCagraIndexParams_t* params);
CUVS_EXPORT cuvsError_t cuvsCagraIndexParamsDestroy(cuvsCagraIndexParams_t params);
CUVS_EXPORT cuvsError_t cuvsCagraCompressionParamsCreate(cuvsCagraCompressionParams_t* params);
CUVS_EXPORT cuvsError_t cuvsCagraCompressionParamsDestroy(cuvsCagraCompressionParams_t params);
CUVS_EXPORT cuvsError_t cuvsAceParamsCreate(cuvsAceParams_t* params);
CUVS_EXPORT cuvsError_t cuvsAceParamsDestroy(cuvsAceParams_t params);
CUVS_EXPORT cuvsError_t cuvsCagraIndexParamsFromHnswParams(cuvsCagraIndexParams_t params,
                                                           int64_t n_rows,
                                                           int64_t dim,
                                                           int M,
                                                           int ef_construction,
                                                           enum cuvsCagraHnswHeuristicType heuristic,
                                                           cuvsDistanceType metric);

struct cuvsCagraExtendParams {
  uint32_t max_chunk_size;
};
typedef struct cuvsCagraExtendParams* cuvsCagraExtendParams_t;
CUVS_EXPORT cuvsError_t cuvsCagraExtendParamsCreate(cuvsCagraExtendParams_t* params);
CUVS_EXPORT cuvsError_t cuvsCagraExtendParamsDestroy(cuvsCagraExtendParams_t params);

enum cuvsCagraSearchAlgo { SINGLE_CTA = 0, MULTI_CTA = 1, MULTI_KERNEL = 2, AUTO = 100 };
enum cuvsCagraHashMode { HASH = 0, SMALL = 1, AUTO_HASH = 100 };

struct cuvsCagraSearchParams {
  size_t max_queries;    /* 0 = auto */
  size_t itopk_size;     /* internal candidate list length, multiple of 32 */
  size_t max_iterations; /* 0 = auto */
  enum cuvsCagraSearchAlgo algo;
  size_t team_size;      /* 0 = auto */
  size_t search_width;
  size_t min_iterations;
  size_t thread_block_size; /* 0 = auto */
  enum cuvsCagraHashMode hashmap_mode;
  size_t hashmap_min_bitlen;
  float hashmap_max_fill_rate;
  uint32_t num_random_samplings;
  uint64_t rand_xor_mask;
  bool persistent;
  float persistent_lifetime;
  float persistent_device_usage;
};
typedef struct cuvsCagraSearchParams* cuvsCagraSearchParams_t;
CUVS_EXPORT cuvsError_t cuvsCagraSearchParamsCreate(cuvsCagraSearchParams_t* params);
CUVS_EXPORT cuvsError_t cuvsCagraSearchParamsDestroy(cuvsCagraSearchParams_t params);

typedef struct {
  uintptr_t addr;
  DLDataType dtype;
} cuvsCagraIndex;
typedef cuvsCagraIndex* cuvsCagraIndex_t;
CUVS_EXPORT cuvsError_t cuvsCagraIndexCreate(cuvsCagraIndex_t* index);
CUVS_EXPORT cuvsError_t cuvsCagraIndexDestroy(cuvsCagraIndex_t index);

CUVS_EXPORT cuvsError_t cuvsCagraIndexGetDims(cuvsCagraIndex_t index, int64_t* dim);
CUVS_EXPORT cuvsError_t cuvsCagraIndexGetSize(cuvsCagraIndex_t index, int64_t* size);
CUVS_EXPORT cuvsError_t cuvsCagraIndexGetGraphDegree(cuvsCagraIndex_t index,
                                                     int64_t* graph_degree);
/* Non-owning views into the index (deleter == NULL). */
CUVS_EXPORT cuvsError_t cuvsCagraIndexGetDataset(cuvsCagraIndex_t index, DLManagedTensor* dataset);
CUVS_EXPORT cuvsError_t cuvsCagraIndexGetGraph(cuvsCagraIndex_t index, DLManagedTensor* graph);

CUVS_EXPORT cuvsError_t cuvsCagraBuild(cuvsResources_t res,
                                       cuvsCagraIndexParams_t params,
                                       DLManagedTensor* dataset,
                                       cuvsCagraIndex_t index);
CUVS_EXPORT cuvsError_t cuvsCagraExtend(cuvsResources_t res,
                                        cuvsCagraExtendParams_t params,
                                        DLManagedTensor* additional_dataset,
                                        cuvsCagraIndex_t index);
/* queries [nq,dim] (index dtype); neighbors [nq,k] uint32 or int64; distances [nq,k] f32. */
CUVS_EXPORT cuvsError_t cuvsCagraSearch(cuvsResources_t res,
                                        cuvsCagraSearchParams_t params,
                                        cuvsCagraIndex_t index,
                                        DLManagedTensor* queries,
                                        DLManagedTensor* neighbors,
                                        DLManagedTensor* distances,
                                        cuvsFilter filter);
CUVS_EXPORT cuvsError_t cuvsCagraSerialize(cuvsResources_t res,
                                           const char* filename,
                                           cuvsCagraIndex_t index,
                                           bool include_dataset);
CUVS_EXPORT cuvsError_t cuvsCagraSerializeToHnswlib(cuvsResources_t res,
                                                    const char* filename,
                                                    cuvsCagraIndex_t index);
CUVS_EXPORT cuvsError_t cuvsCagraDeserialize(cuvsResources_t res,
                                             const char* filename,
                                             cuvsCagraIndex_t index);
/* graph [n, degree] uint32 (host or device); dataset [n, dim] f32 (host or device). */
CUVS_EXPORT cuvsError_t cuvsCagraIndexFromArgs(cuvsResources_t res,
                                               cuvsDistanceType metric,
                                               DLManagedTensor* graph,
                                               DLManagedTensor* dataset,
                                               cuvsCagraIndex_t index);
CUVS_EXPORT cuvsError_t cuvsCagraMerge(cuvsResources_t res,
                                       cuvsCagraIndexParams_t params,
                                       cuvsCagraIndex_t* indices,
                                       size_t num_indices,
                                       cuvsFilter filter,
                                       cuvsCagraIndex_t output_index);
#ifdef __cplusplus
}
#endif
