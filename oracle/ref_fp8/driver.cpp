// C entry points over the REFERENCE's own fp_8bit<5, Signed> (cpp/src/neighbors/ivf_pq/ivf_pq_fp_8bit.cuh:31-100), compiled
// from the reference source where it lies (see Makefile) — test infrastructure: validates oracle.c's restatement
// (oracle_fp8_encode / oracle_fp8_decode) against the real code.  Nothing of the reference is copied into this file.
#include <cmath>
#include <cstdint>
#include <cstdlib>

#include <cuda_fp16.h>

#include <neighbors/ivf_pq/ivf_pq_fp_8bit.cuh>  // resolved against $(REF)/cpp/src

using cuvs::neighbors::ivf_pq::detail::fp_8bit;

extern "C" {
uint8_t ref_fp8_encode(float v, int is_signed)
{
  return is_signed ? fp_8bit<5, true>(v).bitstring : fp_8bit<5, false>(v).bitstring;
}
float ref_fp8_decode(uint8_t b, int is_signed)
{
  return is_signed ? static_cast<float>(fp_8bit<5, true>(b)) : static_cast<float>(fp_8bit<5, false>(b));
}
float ref_fp8_decode_half(uint8_t b, int is_signed)
{
  return is_signed ? __half2float(static_cast<half>(fp_8bit<5, true>(b))) : __half2float(static_cast<half>(fp_8bit<5, false>(b)));
}
}
