// tcgen05 scan + fused top-k' kernel for sm_100a.
//
// Replaces, on B200, the reference's SIMT distance+select kernels
//   fusedL2kNN                cpp/src/neighbors/detail/fused_l2_knn.cuh:186-508   (fp32 FFMA tile + WarpSelect)
//   interleaved_scan (dense)  cpp/src/neighbors/ivf_flat/detail/jit_lto_kernels/interleaved_scan_impl.cuh:70-206
//   coarse GEMM + select_k    cpp/src/neighbors/ivf_pq/ivf_pq_search.cuh:60-168
//   unfused distance-NN       cpp/src/distance/unfused_distance_nn.cuh:54-118
// with one warp-specialised persistent kernel:
//
//   warp 0   TMA producer : query tile (A, resident per work item) + dataset k-blocks (B ring)
//   warp 1   MMA issuer   : tcgen05.mma kind::f16 (bf16 in, fp32 accumulate in TMEM), 128x128 tiles,
//                           split-bf16 products hi*hi + lo*hi + hi*lo  (passes = 3) or hi*hi (passes = 1)
//   warps 2-5 epilogue    : tcgen05.ld one accumulator row per thread, s = hn[col] - acc, compare with
//                           the thread's running k'-th best, rare inserts go through a per-thread smem
//                           queue and are merged into a sorted register list warp-convergently.
//
// The accumulator is double buffered in TMEM (2 x 128 columns) so the epilogue of tile t overlaps
// the MMAs of tile t+1.  Nothing but the k' (score, position) pairs per query row ever leaves the SM:
// the n x nq distance matrix of the reference's unfused path (n*nq*4 bytes through HBM) is never
// materialised.
#include "common.hpp"
#include "ptx_sm100.cuh"
#include "scan_tc.cuh"
#include "timing.hpp"

#include <cuda.h>

#include <cmath>
#include <cstdlib>
#include <mutex>

namespace b200 {
namespace {

constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 bf16
constexpr int kChunk     = 16;         // accumulator columns per tcgen05.ld
constexpr int kQueue     = 24;         // per-thread pending-insert queue; flushed (warp-convergently) when > kQueue - kChunk

template <int KB, int NPL, int EPIW>
struct cfg {
  static constexpr int threads     = 64 + 32 * EPIW;
  static constexpr int stages      = NPL == 2 ? 3 : 6;
  static constexpr int a_bytes     = NPL * KB * kTileBytes;
  static constexpr int stage_bytes = NPL * kTileBytes;
  static constexpr int n_bars      = 2 * stages + 2 + 4;
  static constexpr size_t smem     = 1024 /*align slack*/ + a_bytes + stages * stage_bytes + EPIW * 2 * (512 / EPIW) * 4 /*hn, per warp*/ +
                                 kQueue * (32 * EPIW) * 8 /*queues*/ + n_bars * 8 + 16;
};

// KC > 0: fused top-KC epilogue.  KC == 0: "store" epilogue — every score of the tile is written to
// out_score[out_off + row * out_row_stride + (column within the item's range)] (dense distance block).
// EPIW = 4: one epilogue warp per TMEM lane quarter (128 columns each); EPIW = 8: two per quarter
// (64 columns each, two candidate lists per query row and item).
template <int KB, int NPL, int KC, int EPIW>
__global__ void __launch_bounds__(64 + 32 * EPIW, 1)
tc_scan_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const float* __restrict__ hn, const tc_item* __restrict__ items, int n_items_host,
               const int* __restrict__ n_items_dev, float* __restrict__ out_score, uint32_t* __restrict__ out_pos,
               int64_t out_row_stride, int dbg_skip_epilogue, tc_bound bound)
{
  using C = cfg<KB, NPL, EPIW>;
  constexpr int kEpiThreads = 32 * EPIW;
  const int n_items = n_items_dev ? *n_items_dev : n_items_host;
  // 1024-byte alignment is what SWIZZLE_128B operand tiles need; declared on the array (no integer
  // round-trip of the pointer) so that the compiler keeps every access in the shared address space.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if (threadIdx.x == 0 && (ptx::smem_u32(smem_raw) & 1023u) != 0) __trap();
  uint8_t* sA   = smem_raw;
  uint8_t* sB   = sA + C::a_bytes;
  float* sHn    = reinterpret_cast<float*>(sB + C::stages * C::stage_bytes);  // [EPIW warps][2 buffers][512/EPIW columns]
  uint2* qe     = reinterpret_cast<uint2*>(sHn + EPIW * 2 * (512 / EPIW));      // [kQueue][kEpiThreads] {packed score, group base row}
  uint64_t* bars = reinterpret_cast<uint64_t*>(qe + kQueue * kEpiThreads);
  uint64_t* full    = bars;
  uint64_t* empty   = bars + C::stages;
  uint64_t* a_full  = bars + 2 * C::stages;
  uint64_t* a_empty = a_full + 1;
  uint64_t* tfull   = a_empty + 1;  // [2]
  uint64_t* tempty  = tfull + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA_hi);
    ptx::prefetch_tmap(&tmB_hi);
    if (NPL == 2) {
      ptx::prefetch_tmap(&tmA_lo);
      ptx::prefetch_tmap(&tmB_lo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::stages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(a_full, 1);
    ptx::mbar_init(a_empty, 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tfull[s], 1);
      ptx::mbar_init(&tempty[s], kEpiThreads);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) { ptx::tmem_alloc<256>(tmem_slot); }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, a_phase = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const tc_item item = items[it];
        ptx::mbar_wait(a_empty, a_phase ^ 1);
        ptx::mbar_arrive_expect_tx(a_full, C::a_bytes);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          ptx::tma_load_2d(sA + (0 * KB + kb) * kTileBytes, &tmA_hi, a_full, kb * 64, static_cast<int32_t>(item.a_row0));
          if (NPL == 2)
            ptx::tma_load_2d(sA + (1 * KB + kb) * kTileBytes, &tmA_lo, a_full, kb * 64, static_cast<int32_t>(item.a_row0));
        }
        a_phase ^= 1;
        for (uint32_t t = 0; t < item.n_tiles; ++t) {
          const int32_t brow = static_cast<int32_t>(item.b_row0 + t * 128);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            ptx::mbar_arrive_expect_tx(&full[stage], C::stage_bytes);
            uint8_t* dst = sB + stage * C::stage_bytes;
            ptx::tma_load_2d(dst, &tmB_hi, &full[stage], kb * 64, brow);
            if (NPL == 2) ptx::tma_load_2d(dst + kTileBytes, &tmB_lo, &full[stage], kb * 64, brow);
            if (++stage == C::stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 128);
      const uint32_t a_addr = ptx::smem_u32(sA), b_addr = ptx::smem_u32(sB);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, a_phase = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const uint32_t n_tiles = items[it].n_tiles;
        ptx::mbar_wait(a_full, a_phase);
        ptx::tc_fence_after_sync();
        for (uint32_t t = 0; t < n_tiles; ++t) {
          ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
          ptx::tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + acc * 128;
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            ptx::mbar_wait(&full[stage], phase);
            ptx::tc_fence_after_sync();
            const uint32_t bs = b_addr + stage * C::stage_bytes;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t a_hi = ptx::make_smem_desc_sw128(a_addr + (0 * KB + kb) * kTileBytes + k * 32);
              const uint64_t b_hi = ptx::make_smem_desc_sw128(bs + k * 32);
              ptx::mma_bf16_ss(d_tmem, a_hi, b_hi, idesc, (kb | k) != 0 ? 1u : 0u);
              if (NPL == 2) {
                const uint64_t a_lo = ptx::make_smem_desc_sw128(a_addr + (1 * KB + kb) * kTileBytes + k * 32);
                const uint64_t b_lo = ptx::make_smem_desc_sw128(bs + kTileBytes + k * 32);
                ptx::mma_bf16_ss(d_tmem, a_lo, b_hi, idesc, 1u);
                ptx::mma_bf16_ss(d_tmem, a_hi, b_lo, idesc, 1u);
              }
            }
            ptx::mma_commit(&empty[stage]);  // frees the B stage once these MMAs have read it
            if (++stage == C::stages) { stage = 0; phase ^= 1; }
          }
          ptx::mma_commit(&tfull[acc]);  // accumulator complete -> epilogue
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
        ptx::mma_commit(a_empty);  // all MMAs reading this A tile are done
        a_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (EPIW warps, 1 row x (512/EPIW) columns per thread)
    const int quarter = warp & 3;             // TMEM lanes [32*quarter, 32*quarter+32) belong to this warp
    const int row     = quarter * 32 + lane;  // accumulator row == query row within the tile
    const int half    = (warp - 2) >> 2;      // which column range of the tile (0 when EPIW == 4)
    const int et      = half * 128 + row;     // slot in the queue arrays
    constexpr int kCols   = 512 / EPIW;       // columns per thread and tile
    constexpr int kChunks = kCols / kChunk;
    const int col0        = half * kCols;
    uint32_t acc = 0, acc_phase = 0;
    if constexpr (KC == 0) {
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const tc_item item   = items[it];
        float* hn_w          = sHn + (warp - 2) * 2 * kCols;
        const float* hn_item = hn + item.b_row0 + col0;
        float4 hn_reg        = make_float4(0.f, 0.f, 0.f, 0.f);
        if (item.n_tiles && lane < kCols / 4) hn_reg = *reinterpret_cast<const float4*>(hn_item + lane * 4);
        float* orow          = out_score + item.out_off + static_cast<int64_t>(row) * out_row_stride;
        const bool live      = static_cast<uint32_t>(row) < item.valid_rows;
        for (uint32_t t = 0; t < item.n_tiles; ++t) {
          if (lane < kCols / 4) *reinterpret_cast<float4*>(hn_w + acc * kCols + lane * 4) = hn_reg;
          __syncwarp();
          if (t + 1 < item.n_tiles && lane < kCols / 4) hn_reg = *reinterpret_cast<const float4*>(hn_item + (t + 1) * 128 + lane * 4);
          ptx::mbar_wait(&tfull[acc], acc_phase);
          ptx::tc_fence_after_sync();
#pragma unroll 1
          for (int ch = 0; ch < kChunks; ++ch) {
            uint32_t v[kChunk];
            const int c0 = col0 + ch * kChunk;
            ptx::tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * 128 + c0, v);
            ptx::tmem_ld_wait();
            if (live) {
              const float4* h4 = reinterpret_cast<const float4*>(hn_w + acc * kCols + ch * kChunk);
              float4* o4       = reinterpret_cast<float4*>(orow + t * 128 + c0);
#pragma unroll
              for (int c4 = 0; c4 < kChunk / 4; ++c4) {
                const float4 h = h4[c4];
                o4[c4] = make_float4(h.x - __uint_as_float(v[c4 * 4 + 0]), h.y - __uint_as_float(v[c4 * 4 + 1]),
                                     h.z - __uint_as_float(v[c4 * 4 + 2]), h.w - __uint_as_float(v[c4 * 4 + 3]));
              }
            }
          }
          ptx::tc_fence_before_sync();
          ptx::mbar_arrive(&tempty[acc]);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    } else {
      // Fused top-KC.  The low 4 mantissa bits of every score carry the column index inside its group of 16
      // (a <= 2^-19 relative perturbation, covered by the certificate's eps), so a queue entry is one 64-bit
      // store {packed score, group base row} and the min-tree doubles as an arg-min.
      float* hn_w = sHn + (warp - 2) * 2 * kCols;  // this warp's private staging of the tile's half-norms: no CTA barrier
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const tc_item item = items[it];
        float lv[KC > 0 ? KC : 1];
        uint32_t li[KC > 0 ? KC : 1];
#pragma unroll
        for (int j = 0; j < KC; ++j) { lv[j] = INFINITY; li[j] = 0xffffffffu; }
        // cross-item pruning bound of this row's query (see tc_bound)
        float thr0 = INFINITY, b_add = 0.f;
        int* b_key = nullptr;
        if (bound.keys != nullptr && static_cast<uint32_t>(row) < item.valid_rows) {
          const uint32_t arow = item.a_row0 + row;
          b_key               = bound.keys + (bound.idx ? bound.idx[arow] : arow);
          b_add               = bound.add ? bound.add[arow] : 0.f;
          const int kb        = *reinterpret_cast<volatile int*>(b_key);
          const float bv      = __int_as_float(kb >= 0 ? kb : kb ^ 0x7fffffff);
          thr0                = (bv - b_add) / bound.scale;
        }
        float thr = thr0;
        int cnt   = 0;

        auto flush = [&]() {
          for (int e = 0; e < cnt; ++e) {
            const uint2 en   = qe[e * kEpiThreads + et];
            const float s    = __uint_as_float(en.x);
            const uint32_t p = en.y + (en.x & 15u);
            if (s < lv[KC - 1]) {
#pragma unroll
              for (int j = KC - 1; j > 0; --j) {
                if (s < lv[j - 1]) { lv[j] = lv[j - 1]; li[j] = li[j - 1]; }
                else if (s < lv[j]) { lv[j] = s; li[j] = p; }
              }
              if (s < lv[0]) { lv[0] = s; li[0] = p; }
            }
          }
          cnt = 0;
          thr = fminf(thr0, lv[KC - 1]);
        };

        const float* hn_item = hn + item.b_row0 + col0;
        // lanes 0..kCols/4-1 each stage one float4 of the warp's column range
        float4 hn_reg = make_float4(0.f, 0.f, 0.f, 0.f);
        if (item.n_tiles && lane < kCols / 4) hn_reg = *reinterpret_cast<const float4*>(hn_item + lane * 4);
        for (uint32_t t = 0; t < item.n_tiles; ++t) {
          if (lane < kCols / 4) *reinterpret_cast<float4*>(hn_w + acc * kCols + lane * 4) = hn_reg;
          __syncwarp();
          if (t + 1 < item.n_tiles && lane < kCols / 4) hn_reg = *reinterpret_cast<const float4*>(hn_item + (t + 1) * 128 + lane * 4);
          ptx::mbar_wait(&tfull[acc], acc_phase);
          ptx::tc_fence_after_sync();
          const uint32_t pos0 = item.b_row0 + t * 128 + col0;
          if (!dbg_skip_epilogue) {
#pragma unroll 1
            for (int ch = 0; ch < kCols / 32; ++ch) {
              uint32_t v[32];
              ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * 128 + col0 + ch * 32, v);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                if (__any_sync(0xffffffffu, cnt > kQueue - kChunk)) flush();
                const float4* h4 = reinterpret_cast<const float4*>(hn_w + acc * kCols + ch * 32 + g * 16);
                float sc[16];
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                  const float4 h = h4[c4];
                  sc[c4 * 4 + 0] = __uint_as_float((__float_as_uint(h.x - __uint_as_float(v[g * 16 + c4 * 4 + 0])) & 0xfffffff0u) | (c4 * 4 + 0));
                  sc[c4 * 4 + 1] = __uint_as_float((__float_as_uint(h.y - __uint_as_float(v[g * 16 + c4 * 4 + 1])) & 0xfffffff0u) | (c4 * 4 + 1));
                  sc[c4 * 4 + 2] = __uint_as_float((__float_as_uint(h.z - __uint_as_float(v[g * 16 + c4 * 4 + 2])) & 0xfffffff0u) | (c4 * 4 + 2));
                  sc[c4 * 4 + 3] = __uint_as_float((__float_as_uint(h.w - __uint_as_float(v[g * 16 + c4 * 4 + 3])) & 0xfffffff0u) | (c4 * 4 + 3));
                }
                const float q0 = fminf(fminf(sc[0], sc[1]), fminf(sc[2], sc[3]));
                const float q1 = fminf(fminf(sc[4], sc[5]), fminf(sc[6], sc[7]));
                const float q2 = fminf(fminf(sc[8], sc[9]), fminf(sc[10], sc[11]));
                const float q3 = fminf(fminf(sc[12], sc[13]), fminf(sc[14], sc[15]));
                const float mn = fminf(fminf(q0, q1), fminf(q2, q3));
                if (mn < thr) {
                  const uint32_t gbase = pos0 + ch * 32 + g * 16;
                  uint2* qp = qe + cnt * kEpiThreads + et;
#pragma unroll
                  for (int c = 0; c < 16; ++c) {
                    if (sc[c] < thr) { *qp = make_uint2(__float_as_uint(sc[c]), gbase); qp += kEpiThreads; ++cnt; }
                  }
                }
              }
            }
          }
          ptx::tc_fence_before_sync();
          ptx::mbar_arrive(&tempty[acc]);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
        flush();
        if (b_key != nullptr && lv[KC - 1] < INFINITY) {
          const float pub = b_add + bound.scale * lv[KC - 1];
          const int kp    = __float_as_int(pub);
          atomicMin(b_key, kp >= 0 ? kp : kp ^ 0x7fffffff);
        }
        if (static_cast<uint32_t>(row) < item.valid_rows) {
          float* os    = out_score + item.out_off + static_cast<int64_t>(row) * out_row_stride + half * KC;
          uint32_t* op = out_pos + item.out_off + static_cast<int64_t>(row) * out_row_stride + half * KC;
#pragma unroll
          for (int j = 0; j < KC; ++j) { os[j] = lv[j]; op[j] = li[j]; }
        }
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) { ptx::tmem_dealloc<256>(tmem_base); }
}

// ------------------------------------------------------------------------------------ host side
using encode_fn_t = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_fn_t get_encode_fn()
{
  static encode_fn_t fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_fn_t>(p);
  });
  B2_EXPECTS(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  return fn;
}

CUtensorMap make_plane_map(const __nv_bfloat16* ptr, int64_t rows, int Kp)
{
  CUtensorMap m;
  cuuint64_t gdim[2]    = {static_cast<cuuint64_t>(Kp), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(Kp) * sizeof(__nv_bfloat16)};
  cuuint32_t box[2]     = {64, 128};
  cuuint32_t estr[2]    = {1, 1};
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(ptr), gdim, gstride, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2_EXPECTS(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d (rows=%lld Kp=%d)", int(r), (long long)rows, Kp);
  return m;
}

int env_int(const char* name, int dflt)
{
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

template <int KB, int NPL, int KC, int EPIW>
void launch(cudaStream_t stream, int sm_count, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
            const CUtensorMap& b_lo, const float* hn, const tc_item* items, int n_items, const int* n_items_dev,
            float* out_score, uint32_t* out_pos, int64_t out_row_stride, const tc_bound& bound)
{
  auto kern = tc_scan_kernel<KB, NPL, KC, EPIW>;
  using C   = cfg<KB, NPL, EPIW>;
  static const int skip_epi = env_int("CUVS_B200_TC_SKIP_EPI", 0);  // profiling knob: MMA/TMA pipeline only
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(C::smem)));
    attr_set = true;
  }
  int grid = n_items < sm_count ? n_items : sm_count;
  timed_section ts("tc_scan", stream);
  count_launch();
  kern<<<grid, C::threads, C::smem, stream>>>(a_hi, a_lo, b_hi, b_lo, hn, items, n_items, n_items_dev, out_score, out_pos,
                                               out_row_stride, skip_epi, bound);
  B2_CUDA(cudaGetLastError());
}

__global__ void split_planes_kernel(const float* __restrict__ x, int64_t n, int64_t ld, int d, int Kp,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t rows_pad,
                                    const float* __restrict__ row_scale)
{
  const int64_t total = rows_pad * (Kp / 2);
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = t / (Kp / 2);
    const int c     = static_cast<int>(t % (Kp / 2)) * 2;
    float v0 = 0.f, v1 = 0.f;
    if (r < n) {
      const float sc = row_scale ? row_scale[r] : 1.0f;
      if (c < d) v0 = x[r * ld + c] * sc;
      if (c + 1 < d) v1 = x[r * ld + c + 1] * sc;
    }
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
    reinterpret_cast<__nv_bfloat162*>(hi)[t] = __nv_bfloat162(h0, h1);
    if (lo) {
      const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0));
      const __nv_bfloat16 l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
      reinterpret_cast<__nv_bfloat162*>(lo)[t] = __nv_bfloat162(l0, l1);
    }
  }
}

__global__ void half_norms_kernel(const float* __restrict__ xn, int64_t n, int64_t rows_pad, float* __restrict__ hn)
{
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= rows_pad) return;
  hn[i] = i < n ? (xn ? 0.5f * xn[i] : 0.0f) : INFINITY;
}

}  // namespace

int tc_lists_per_item()
{
  static const int epiw = env_int("CUVS_B200_TC_EPIW", 8);
  return epiw == 4 ? 1 : 2;
}

bool tc_supported(int device, int d)
{
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) return false;
  return major == 10 && d >= 1 && tc_pad_k(d) <= 128;
}

void tc_split_planes(cudaStream_t stream, const float* x, int64_t n, int64_t ld, int d, int Kp, __nv_bfloat16* hi,
                     __nv_bfloat16* lo, int64_t rows_pad, const float* row_scale)
{
  if (rows_pad == 0) return;
  int64_t total = rows_pad * (Kp / 2);
  int blocks    = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 16));
  count_launch();
  split_planes_kernel<<<blocks, 256, 0, stream>>>(x, n, ld, d, Kp, hi, lo, rows_pad, row_scale);
  B2_CUDA(cudaGetLastError());
}

void tc_half_norms(cudaStream_t stream, const float* xn, int64_t n, int64_t rows_pad, float* hn)
{
  if (rows_pad == 0) return;
  count_launch();
  half_norms_kernel<<<static_cast<unsigned>((rows_pad + 255) / 256), 256, 0, stream>>>(xn, n, rows_pad, hn);
  B2_CUDA(cudaGetLastError());
}

void tc_scan_topk(cudaStream_t stream, int device, const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo,
                  int64_t a_rows_pad, const __nv_bfloat16* b_hi, const __nv_bfloat16* b_lo, int64_t b_rows_pad, int Kp,
                  const float* hn, const tc_item* items_dev, int n_items, const int* n_items_dev, int KC, int passes,
                  float* out_score, uint32_t* out_pos, int64_t out_row_stride, const tc_bound* bound)
{
  if (n_items == 0) return;  // n_items is the host-side upper bound (grid sizing); *n_items_dev, when given, is the exact count
  B2_EXPECTS(Kp == 64 || Kp == 128, "tc_scan_topk: padded K must be 64 or 128 (got %d)", Kp);
  B2_EXPECTS(KC == 0 || KC == 16 || KC == 32, "tc_scan_topk: KC must be 0 (store), 16 or 32");
  B2_EXPECTS(passes == 1 || passes == 3, "tc_scan_topk: passes must be 1 or 3");
  B2_EXPECTS(passes == 1 || (a_lo && b_lo), "tc_scan_topk: lo planes required for 3-pass mode");
  const int sms  = sm_count_of(device);
  const int epiw = tc_lists_per_item() * 4;
  static const int no_bound = env_int("CUVS_B200_TC_NO_BOUND", 0);  // profiling knob
  const tc_bound bnd = (bound && !no_bound) ? *bound : tc_bound{};
  CUtensorMap mA  = make_plane_map(a_hi, a_rows_pad, Kp);
  CUtensorMap mB  = make_plane_map(b_hi, b_rows_pad, Kp);
  CUtensorMap mAl = passes == 3 ? make_plane_map(a_lo, a_rows_pad, Kp) : mA;
  CUtensorMap mBl = passes == 3 ? make_plane_map(b_lo, b_rows_pad, Kp) : mB;
#define B2_TC_CASE(KB_, NPL_, KC_)                                                                                     \
  if (Kp == 64 * KB_ && (passes == 3 ? 2 : 1) == NPL_ && KC == KC_)                                                    \
  {                                                                                                                    \
    if (epiw == 8)                                                                                                     \
      return launch<KB_, NPL_, KC_, 8>(stream, sms, mA, mAl, mB, mBl, hn, items_dev, n_items, n_items_dev, out_score, out_pos, out_row_stride, bnd); \
    return launch<KB_, NPL_, KC_, 4>(stream, sms, mA, mAl, mB, mBl, hn, items_dev, n_items, n_items_dev, out_score, out_pos, out_row_stride, bnd);   \
  }
  B2_TC_CASE(1, 1, 16) B2_TC_CASE(1, 1, 32) B2_TC_CASE(1, 2, 16) B2_TC_CASE(1, 2, 32) B2_TC_CASE(1, 1, 0) B2_TC_CASE(1, 2, 0)
  B2_TC_CASE(2, 1, 16) B2_TC_CASE(2, 1, 32) B2_TC_CASE(2, 2, 16) B2_TC_CASE(2, 2, 32) B2_TC_CASE(2, 1, 0) B2_TC_CASE(2, 2, 0)
#undef B2_TC_CASE
  B2_FAIL("tc_scan_topk: no kernel for Kp=%d passes=%d KC=%d", Kp, passes, KC);
}

}  // namespace b200
