/*
 * cuvs_b200 — core C boundary: errors, logging, resources, memory, matrix copy.
 *
 * Binary-compatible restatement of the reference's c/include/cuvs/core/c_api.h
 * (error enum :27, error text :32/:38, log level :50-67, cuvsResources_t :80,
 * resource/stream calls :88-132, multi-GPU resources :141-169, RMM alloc
 * :186-235, version :244, matrix copy/slice :261/:272).  Every entry point is
 * implemented in cuvs_b200/csrc/c_api_core.cu on plain CUDA streams and
 * stream-ordered allocation; no RAFT/rmm behind it.
 */
#pragma once

#include <cuda_runtime.h>
#include <dlpack/dlpack.h>
#include <stdbool.h>
#include <stdint.h>

#include <cuvs/core/export.h>

#ifdef __cplusplus
extern "C" {
#endif

/* c_api.h:27 — status of every call; text of the last failure is per-thread. */
typedef enum { CUVS_ERROR = 0, CUVS_SUCCESS = 1 } cuvsError_t;

CUVS_EXPORT const char* cuvsGetLastErrorText(); /* NULL when the last call succeeded (:32) */
CUVS_EXPORT void cuvsSetLastErrorText(const char* error); /* NULL clears (:38) */

/* c_api.h:50-58 */
typedef enum {
  CUVS_LOG_LEVEL_TRACE    = 0,
  CUVS_LOG_LEVEL_DEBUG    = 1,
  CUVS_LOG_LEVEL_INFO     = 2,
  CUVS_LOG_LEVEL_WARN     = 3,
  CUVS_LOG_LEVEL_ERROR    = 4,
  CUVS_LOG_LEVEL_CRITICAL = 5,
  CUVS_LOG_LEVEL_OFF      = 6
} cuvsLogLevel_t;

CUVS_EXPORT cuvsLogLevel_t cuvsGetLogLevel();
CUVS_EXPORT void cuvsSetLogLevel(cuvsLogLevel_t);

/* c_api.h:80 — opaque handle: device id + stream + workspace of one caller thread. */
typedef uintptr_t cuvsResources_t;

CUVS_EXPORT cuvsError_t cuvsResourcesCreate(cuvsResources_t* res);
CUVS_EXPORT cuvsError_t cuvsResourcesDestroy(cuvsResources_t res);
CUVS_EXPORT cuvsError_t cuvsStreamSet(cuvsResources_t res, cudaStream_t stream);
CUVS_EXPORT cuvsError_t cuvsStreamGet(cuvsResources_t res, cudaStream_t* stream);
CUVS_EXPORT cuvsError_t cuvsStreamSync(cuvsResources_t res);
CUVS_EXPORT cuvsError_t cuvsDeviceIdGet(cuvsResources_t res, int* device_id);

/* c_api.h:141-169 — single-process multi-GPU handle (one stream per visible device). */
CUVS_EXPORT cuvsError_t cuvsMultiGpuResourcesCreate(cuvsResources_t* res);
CUVS_EXPORT cuvsError_t cuvsMultiGpuResourcesCreateWithDeviceIds(cuvsResources_t* res,
                                                                 DLManagedTensor* device_ids);
CUVS_EXPORT cuvsError_t cuvsMultiGpuResourcesDestroy(cuvsResources_t res);
CUVS_EXPORT cuvsError_t cuvsMultiGpuResourcesSetMemoryPool(cuvsResources_t res,
                                                           int percent_of_free_memory);

/* c_api.h:186-235 — device / pinned-host allocation on the handle's stream. */
CUVS_EXPORT cuvsError_t cuvsRMMAlloc(cuvsResources_t res, void** ptr, size_t bytes);
CUVS_EXPORT cuvsError_t cuvsRMMFree(cuvsResources_t res, void* ptr, size_t bytes);
CUVS_EXPORT cuvsError_t cuvsRMMPoolMemoryResourceEnable(int initial_pool_size_percent,
                                                        int max_pool_size_percent,
                                                        bool managed);
CUVS_EXPORT cuvsError_t cuvsRMMMemoryResourceReset();
CUVS_EXPORT cuvsError_t cuvsRMMHostAlloc(void** ptr, size_t bytes);
CUVS_EXPORT cuvsError_t cuvsRMMHostFree(void* ptr, size_t bytes);

/* c_api.h:244 */
CUVS_EXPORT cuvsError_t cuvsVersionGet(uint16_t* major, uint16_t* minor, uint16_t* patch);

/* c_api.h:261 / :272 — strided 2-D copy between host/device tensors, and a row slice view. */
CUVS_EXPORT cuvsError_t cuvsMatrixCopy(cuvsResources_t res,
                                       DLManagedTensor* src,
                                       DLManagedTensor* dst);
CUVS_EXPORT cuvsError_t cuvsMatrixSliceRows(
  cuvsResources_t res, DLManagedTensor* src, int64_t start, int64_t end, DLManagedTensor* dst);

#ifdef __cplusplus
}
#endif
