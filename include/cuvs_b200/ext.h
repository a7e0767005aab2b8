/*
 * cuvs_b200 extensions to the cuVS C ABI (not part of the reference's surface):
 * per-kernel CUDA-event timing for bench.py's roofline block, and introspection hooks used by the
 * parity tests.  None of these is needed by a binding that only uses the reference's API.
 */
#pragma once
#include <cuvs/core/c_api.h>
#include <cuvs/neighbors/brute_force.h>
#include <cuvs/neighbors/cagra.h>
#include <cuvs/neighbors/ivf_flat.h>
#include <cuvs/neighbors/ivf_pq.h>
#include <stdbool.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Enable/disable recording of CUDA events around the library's dominant kernels (on the stream they
 * are launched on).  Disabled by default (zero overhead). */
CUVS_EXPORT void cuvsB200TimingEnable(int on);
/* Discards everything recorded so far. */
CUVS_EXPORT void cuvsB200TimingReset(void);
/* Synchronises the recorded events of section `name` ("tc_scan", "pq_scan", "cagra_search", ...)
 * and returns their summed duration in milliseconds; *count receives the number of launches. */
CUVS_EXPORT double cuvsB200TimingTotalMs(const char* name, int* count);

/* Number of queries of the calling thread's last cuvsBruteForceSearch that failed the candidate
 * certificate and were recomputed on the exact path, and the number of kernels that search launched. */
CUVS_EXPORT int cuvsB200LastFlagged(void);
CUVS_EXPORT long long cuvsB200KernelLaunches(void);

/* Raw candidate output of the tensor-core scan of brute force (after the per-split merge):
 * cand_pos [nq, kc] uint32, cand_score [nq, kc] f32 (approximate selection-form distance). kc in {16, 32}. */
CUVS_EXPORT cuvsError_t cuvsB200BruteForceCandidates(cuvsResources_t res,
                                                     cuvsBruteForceIndex_t index,
                                                     DLManagedTensor* queries,
                                                     DLManagedTensor* cand_pos,
                                                     DLManagedTensor* cand_score);
/* IVF-Flat list introspection (the reference exposes the equivalent only for IVF-PQ:
 * cuvsIvfPqIndexGetListSizes / cuvsIvfPqIndexGetListIndices).  Non-owning views into the index. */
CUVS_EXPORT cuvsError_t cuvsB200IvfFlatGetListSizes(cuvsIvfFlatIndex_t index, DLManagedTensor* list_sizes /*[n_lists] u32*/);
CUVS_EXPORT cuvsError_t cuvsB200IvfFlatGetListIndices(cuvsIvfFlatIndex_t index, uint32_t label, DLManagedTensor* ids /*[size] i64*/);
CUVS_EXPORT cuvsError_t cuvsB200IvfFlatGetSize(cuvsIvfFlatIndex_t index, int64_t* size);
/* Replace the coarse centres of an (empty) index — used by the list-sharded multi-GPU build so that every rank
 * partitions the data with bit-identical centres (trained on one rank, broadcast over NCCL). centers: [n_lists, dim] f32. */
CUVS_EXPORT cuvsError_t cuvsB200IvfFlatSetCenters(cuvsResources_t res, cuvsIvfFlatIndex_t index, DLManagedTensor* centers);

/* IVF-PQ: which fine-scan kernel cuvsIvfPqSearch will use on this index by default, and the device bytes the index holds.
 * *path: bit 1 (value 2) = the index holds the code stream (code-streaming tcgen05 scan, scan_pq.cu), bit 0 (value 1) = it holds
 * decoded bf16 rows (decoded-row tcgen05 scan, scan_tc.cu; on a streamed index: the small-index cache used by densely probing
 * batches), 0 = LUT kernel only. */
CUVS_EXPORT cuvsError_t cuvsB200IvfPqIndexInfo(cuvsIvfPqIndex_t index, int* path, int64_t* device_bytes);

/* ---- multi-GPU exchange step (one process per GPU) ------------------------------------------------
 * The index is sharded by IVF list: every rank searches the lists it owns for the whole query batch and holds a partial
 * top-k [n_queries, k] with GLOBAL ids.  cuvsB200AllGatherMergeTopK enqueues, on the handle's stream, ONE ncclAllGather of
 * the packed partials (n_queries*k*12 bytes per rank) and the k-way merge; the merged result is identical on all ranks.
 * No host synchronisation.  Bootstrap: rank 0 calls cuvsB200NcclUniqueId, the 128 bytes travel to the other ranks by any
 * means (torch.distributed broadcast, MPI, a file), every rank calls cuvsB200CommCreate.  NCCL is dlopen'ed on first use. */
typedef struct cuvsB200Comm* cuvsB200Comm_t;
CUVS_EXPORT cuvsError_t cuvsB200NcclUniqueId(void* id128);
CUVS_EXPORT cuvsError_t cuvsB200CommCreate(cuvsResources_t res, const void* id128, int rank, int world, cuvsB200Comm_t* comm);
CUVS_EXPORT cuvsError_t cuvsB200CommDestroy(cuvsB200Comm_t comm);
CUVS_EXPORT cuvsError_t cuvsB200AllGatherMergeTopK(cuvsResources_t res,
                                                   cuvsB200Comm_t comm,
                                                   DLManagedTensor* distances,     /* [n_queries, k] f32, this rank's partial */
                                                   DLManagedTensor* neighbors,     /* [n_queries, k] i64, global ids */
                                                   DLManagedTensor* out_distances, /* [n_queries, k] f32 */
                                                   DLManagedTensor* out_neighbors, /* [n_queries, k] i64 */
                                                   bool select_min);

/* CAGRA: the graph walk is bound by random row gathers from HBM.  bits = 16 makes the index keep an fp16 copy of the
 * vectors that the walk reads instead (half the bytes); the best 32 entries of every query's final list are re-ranked
 * with the fp32 rows before the k results are returned (k <= 32).  bits = 32 (default) restores the exact fp32 walk. */
CUVS_EXPORT cuvsError_t cuvsB200CagraSetWalkPrecision(cuvsResources_t res, cuvsCagraIndex_t index, int bits);

#ifdef __cplusplus
}
#endif
