// TMA helpers shared by the TMA-fed tcgen05 kernels: PTX wrappers (device) and tensor-map
// construction through the driver entry point (host; no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include "tc_common.cuh"

namespace rcfd {
namespace tma {

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

inline CUtensorMapSwizzle swizzle_for(int bkc) {
  return bkc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bkc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

inline bool make_act_map(CUtensorMap* m, const void* ptr, int n, int h, int w, int c, int bkc, int tw, int th, int stride) {
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  cuuint32_t box[4] = {(cuuint32_t)bkc, (cuuint32_t)(tw * stride), (cuuint32_t)(th * stride), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  return get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bkc), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline bool make_w_map(CUtensorMap* m, const void* ptr, int cout, int K, int bkc, int bn) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)cout};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)bkc, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  return get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bkc), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Row-streaming kernels: `cols` independent column strips of `h` rows are cut into chunks that `ctas` persistent
// CTAs take round-robin.  Pick the chunk height that minimises  waves x (rows per chunk + per-item overhead rows):
// a partly filled last wave costs a full chunk, and every chunk pays its halo rows and pipeline fill.
inline void plan_row_chunks(int h, int cols, int ctas, int overhead_rows, int min_rows, int* rows_per_chunk,
                            int* chunks_per_col) {
  int best_rpc = h;
  long best_cost = -1;
  for (int cpc = 1; cpc <= h; ++cpc) {
    const int rpc = (h + cpc - 1) / cpc;
    if (cpc > 1 && rpc < min_rows) break;
    const int chunks = (h + rpc - 1) / rpc;
    const long items = (long)cols * chunks;
    const long waves = (items + ctas - 1) / ctas;
    const long cost = waves * (rpc + overhead_rows);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_rpc = rpc; }
  }
  *rows_per_chunk = best_rpc;
  *chunks_per_col = (h + best_rpc - 1) / best_rpc;
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}


}  // namespace tma
}  // namespace rcfd
