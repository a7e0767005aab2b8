/* Symbol visibility for the drop-in libcuvs_c boundary.
 * Replaces: c/include/cuvs/core/export.h:8-18 (reference). */
#pragma once
#if defined(__GNUC__) && !defined(__MINGW32__) && !defined(__MINGW64__)
#define CUVS_EXPORT __attribute__((visibility("default")))
#define CUVS_HIDDEN __attribute__((visibility("hidden")))
#else
#define CUVS_EXPORT
#define CUVS_HIDDEN
#endif
