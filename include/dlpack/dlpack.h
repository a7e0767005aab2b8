/*
 * Minimal DLPack (v0.8 ABI) declarations used by the cuVS C boundary.
 *
 * The cuVS C ABI passes every tensor as a `DLManagedTensor*`
 * (reference: c/include/cuvs/core/c_api.h:10 includes <dlpack/dlpack.h>;
 * pinned to dlpack 0.8 by cpp/cmake/thirdparty/get_dlpack.cmake:32).
 * Only the struct layouts matter for binary compatibility; they are the
 * public DLPack standard layouts, restated here so that the library builds
 * without any third-party tree.
 */
#ifndef CUVS_B200_DLPACK_H_
#define CUVS_B200_DLPACK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLPACK_MAJOR_VERSION 0
#define DLPACK_MINOR_VERSION 8

typedef enum {
  kDLCPU         = 1,
  kDLCUDA        = 2,
  kDLCUDAHost    = 3,
  kDLOpenCL      = 4,
  kDLVulkan      = 7,
  kDLMetal       = 8,
  kDLVPI         = 9,
  kDLROCM        = 10,
  kDLROCMHost    = 11,
  kDLExtDev      = 12,
  kDLCUDAManaged = 13,
  kDLOneAPI      = 14,
  kDLWebGPU      = 15,
  kDLHexagon     = 16
} DLDeviceType;

typedef struct {
  DLDeviceType device_type;
  int32_t device_id;
} DLDevice;

typedef enum {
  kDLInt          = 0U,
  kDLUInt         = 1U,
  kDLFloat        = 2U,
  kDLOpaqueHandle = 3U,
  kDLBfloat       = 4U,
  kDLComplex      = 5U,
  kDLBool         = 6U
} DLDataTypeCode;

typedef struct {
  uint8_t code;   /* DLDataTypeCode */
  uint8_t bits;   /* bits per lane, e.g. 32 */
  uint16_t lanes; /* 1 for scalars */
} DLDataType;

typedef struct {
  void* data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t* shape;
  int64_t* strides; /* in elements; NULL means compact row-major */
  uint64_t byte_offset;
} DLTensor;

typedef struct DLManagedTensor {
  DLTensor dl_tensor;
  void* manager_ctx;
  void (*deleter)(struct DLManagedTensor* self);
} DLManagedTensor;

#ifdef __cplusplus
}
#endif
#endif /* CUVS_B200_DLPACK_H_ */
