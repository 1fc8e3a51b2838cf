// Every weight packing (OIHW float -> the kernels' packed layouts) and every weight-gradient unpacking of a
// training step in ONE launch each: a device table of items, block -> item by binary search over the items'
// first block.  Same index maps and roundings as the single-tensor kernels of elementwise.cu
// (rcfd_pack_conv_weight & co), which stay for inference / one-off calls.
//
// Both directions are transpositions of short rows ((channel, tap) <-> (tap, channel) inside one output channel's
// K-row), so a block stages a row (or a [channel tile] x [cout chunk] tile for the dgrad layout) in shared memory:
// global reads and writes are both contiguous runs.  Items whose rows do not fit take the element-wise path.
#include "common.cuh"

namespace rcfd {
namespace {

constexpr int PB_THREADS = 256;
constexpr int ROW_FLOATS = 4864;          // 19 KB staging row: cin * taps of one output channel (512 x 9), or a padded dgrad tile
constexpr int DG_CI = 4;                  // dgrad tile: 4 input channels x 128 output channels x taps
constexpr int DG_CO = 128;

__host__ __device__ inline bool row_path(const rcfd_pack_item& it) {
  switch (it.kind) {
    case RCFD_PACK_FWD: return it.cin_cnt * it.taps <= ROW_FLOATS && it.cpad * it.taps <= 4 * ROW_FLOATS;
    case RCFD_PACK_DGRAD: return (DG_CI * it.taps + 1) * DG_CO <= ROW_FLOATS && it.cin_off + it.cin_cnt <= it.cin;   // padded rows: element-wise path
    case RCFD_PACK_UP2X: return it.cin * 9 <= ROW_FLOATS;
    case RCFD_UNPACK_CONV: return it.cpad * it.taps <= ROW_FLOATS;
    default: return false;
  }
}

__host__ __device__ inline int item_blocks(const rcfd_pack_item& it) {
  if (row_path(it)) {
    switch (it.kind) {
      case RCFD_PACK_DGRAD: return ((it.cin_cnt + DG_CI - 1) / DG_CI) * ((it.cout + DG_CO - 1) / DG_CO);
      default: return it.cout;            // one output channel (K-row) per block
    }
  }
  return (int)((it.total + RCFD_PACK_BLOCK_ELEMS - 1) / RCFD_PACK_BLOCK_ELEMS);
}

template <typename T>
__device__ __forceinline__ void pack_one(const rcfd_pack_item& it, uint32_t i) {
  T* out = reinterpret_cast<T*>(it.dst);
  const float* __restrict__ w = it.src;
  const uint32_t taps = it.taps;
  switch (it.kind) {
    case RCFD_PACK_FWD: {                      // out[co][tap][ci < cpad]
      const uint32_t ci = i % it.cpad, r = i / it.cpad;
      const uint32_t tap = r % taps, co = r / taps;
      out[i] = from_f<T>((int)ci < it.cin_cnt ? __ldg(w + ((size_t)co * it.cin + it.cin_off + ci) * taps + tap) : 0.f);
      break;
    }
    case RCFD_PACK_DGRAD: {                    // out[ci][flipped tap][col_off + co], rows dst_cols wide
      const uint32_t co = i % it.cout, r = i / it.cout;
      const uint32_t tap = r % taps, ci = r / taps;
      out[((size_t)ci * taps + tap) * it.dst_cols + it.col_off + co] =
          from_f<T>((int)(it.cin_off + ci) < it.cin ? __ldg(w + ((size_t)co * it.cin + it.cin_off + ci) * taps + (taps - 1 - tap)) : 0.f);
      break;
    }
    case RCFD_PACK_UP2X: {                     // out[phase][co][2x2 tap][ci]: sums of the 3x3 taps hitting one low-res pixel
      const uint32_t cin = it.cin, cout = it.cout;
      const uint32_t ci = i % cin;
      uint32_t r = i / cin;
      const uint32_t tap = r % 4; r /= 4;
      const uint32_t co = r % cout, ph = r / cout;
      const int a = ph >> 1, b = ph & 1, t = tap >> 1, u = tap & 1;
      const int r0 = a == 0 ? (t == 0 ? 0 : 1) : (t == 0 ? 0 : 2), r1 = a == 0 ? (t == 0 ? 0 : 2) : (t == 0 ? 1 : 2);
      const int s0 = b == 0 ? (u == 0 ? 0 : 1) : (u == 0 ? 0 : 2), s1 = b == 0 ? (u == 0 ? 0 : 2) : (u == 0 ? 1 : 2);
      const float* wp = w + ((size_t)co * cin + ci) * 9;
      float acc = 0.f;
      for (int rr = r0; rr <= r1; ++rr)
        for (int ss = s0; ss <= s1; ++ss) acc += __ldg(wp + rr * 3 + ss);
      out[i] = from_f<T>(acc);
      break;
    }
    case RCFD_PACK_UPCONV_DGRAD: {             // out[ci][4x4 tap][co < cpad]
      const uint32_t co = i % it.cpad;
      const uint32_t r = i / it.cpad;
      const uint32_t tap = r % 16, ci = r / 16;
      out[i] = from_f<T>(upconv_dgrad_weight(w, it.cout, it.cin, it.cin_off + ci, co, tap));
      break;
    }
    case RCFD_PACK_DGRAD_S2: {                 // out[phase][ci][2x2 tap][co < cpad]
      const uint32_t co = i % it.cpad;
      uint32_t r = i / it.cpad;
      const uint32_t tap = r % 4; r /= 4;
      const uint32_t ci = r % it.cin_cnt, ph = r / it.cin_cnt;
      out[i] = from_f<T>(dgrad_s2_weight(w, it.cout, it.cin, it.cin_off + ci, co, ph, tap));
      break;
    }
    default: {                                 // RCFD_PACK_STEM_S2D: out[co][4x4 tap][cpad], 7x7 window on the s2d tensor
      const uint32_t C = it.cin, CP = it.cpad;
      const uint32_t ch = i % CP, r = i / CP;
      const uint32_t tap = r % 16, co = r / 16;
      float v = 0.f;
      if (ch < 4 * C) {
        const int ph = ch / C, c = ch - ph * C;
        const int rr = 2 * (tap >> 2) + (ph >> 1) - 1, ss = 2 * (tap & 3) + (ph & 1) - 1;
        if (rr >= 0 && rr < 7 && ss >= 0 && ss < 7) v = __ldg(w + (((size_t)co * C + c) * 7 + rr) * 7 + ss);
      }
      out[i] = from_f<T>(v);
    }
  }
}

__device__ __forceinline__ void unpack_one(const rcfd_pack_item& it, uint32_t i) {
  float* g = reinterpret_cast<float*>(it.dst);
  const float* __restrict__ packed = it.src;
  if (it.kind == RCFD_UNPACK_CONV) {           // i runs over [co][ci < cin_cnt][tap]: the OIHW slice
    const uint32_t taps = it.taps;
    const uint32_t tap = i % taps, r = i / taps;
    const uint32_t ci = r % it.cin_cnt, co = r / it.cin_cnt;
    g[((size_t)co * it.cin + it.cin_off + ci) * taps + tap] = packed[((size_t)co * taps + tap) * it.cpad + ci];
  } else if (it.kind == RCFD_UNPACK_STEM_S2D) {
    const uint32_t C = it.cin, CP = it.cpad;
    const uint32_t ss = i % 7;
    uint32_t r = i / 7;
    const uint32_t rr = r % 7; r /= 7;
    const uint32_t c = r % C, co = r / C;
    const uint32_t ty = (rr + 1) >> 1, dy = (rr + 1) & 1, tx = (ss + 1) >> 1, dx = (ss + 1) & 1;
    g[i] = packed[((size_t)co * 16 + ty * 4 + tx) * CP + (dy * 2 + dx) * C + c];
  } else {                                     // RCFD_COPY_F32
    g[i] = packed[i];
  }
}

// ---- staged (shared-memory) paths: one K-row / tile per block.  The loops are arranged so that no per-element
// integer division by a run-time value remains (the first version was instruction bound on them).
__device__ __forceinline__ int div_taps(int j, int taps) {      // taps is 1, 9 or 16 in every shipped model
  switch (taps) {
    case 1: return j;
    case 9: return j / 9;
    case 16: return j >> 4;
    case 4: return j >> 2;
    default: return j / taps;
  }
}

// Contiguous global run -> shared memory with several loads in flight per thread (one 4-byte load per thread and
// iteration left the kernel latency bound: ~1.5 TB/s of the 6.4 available).
__device__ __forceinline__ void load_run(const float* __restrict__ src, float* dst, int n, int tid, int nthreads) {
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (n & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    const int n4 = n >> 2;
    int j = tid;
    for (; j + 3 * nthreads < n4; j += 4 * nthreads) {
      const float4 a = __ldg(s4 + j), b = __ldg(s4 + j + nthreads), c = __ldg(s4 + j + 2 * nthreads),
                   d = __ldg(s4 + j + 3 * nthreads);
      d4[j] = a; d4[j + nthreads] = b; d4[j + 2 * nthreads] = c; d4[j + 3 * nthreads] = d;
    }
    for (; j < n4; j += nthreads) d4[j] = __ldg(s4 + j);
    return;
  }
  int j = tid;
  for (; j + 3 * nthreads < n; j += 4 * nthreads) {
    const float a = __ldg(src + j), b = __ldg(src + j + nthreads), c = __ldg(src + j + 2 * nthreads),
                d = __ldg(src + j + 3 * nthreads);
    dst[j] = a; dst[j + nthreads] = b; dst[j + 2 * nthreads] = c; dst[j + 3 * nthreads] = d;
  }
  for (; j < n; j += nthreads) dst[j] = __ldg(src + j);
}

template <typename T>
__device__ __forceinline__ void pack_fwd_row(const rcfd_pack_item& it, int co, float* row) {
  const int taps = it.taps, n = it.cin_cnt * taps;
  const float* __restrict__ src = it.src + ((size_t)co * it.cin + it.cin_off) * taps;     // [ci][tap], contiguous
  load_run(src, row, n, threadIdx.x, PB_THREADS);
  __syncthreads();
  T* out = reinterpret_cast<T*>(it.dst) + (size_t)co * taps * it.cpad;                    // [tap][ci < cpad]
  for (int tap = 0; tap < taps; ++tap)
    for (int ci = threadIdx.x; ci < it.cpad; ci += PB_THREADS)
      out[tap * it.cpad + ci] = from_f<T>(ci < it.cin_cnt ? row[ci * taps + tap] : 0.f);
}

template <typename T>
__device__ __forceinline__ void pack_dgrad_tile(const rcfd_pack_item& it, int blk, float* tile) {
  const int taps = it.taps;
  const int co_chunks = (it.cout + DG_CO - 1) / DG_CO;
  const int ci0 = (blk / co_chunks) * DG_CI, co0 = (blk % co_chunks) * DG_CO;
  const int nci = min(DG_CI, it.cin_cnt - ci0), nco = min(DG_CO, it.cout - co0);
  const int run = nci * taps;                 // contiguous floats per output channel: w[co][cin_off + ci0 ..][tap]
  const int ld = run | 1;                     // odd row pitch: the transposed reads below hit distinct banks
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a warp takes 4 output channels at a time: up to 8 independent loads per lane in flight
  const size_t co_pitch = (size_t)it.cin * taps;
  const float* __restrict__ src0 = it.src + ((size_t)co0 * it.cin + it.cin_off + ci0) * taps;
  for (int co = warp * 4; co < nco; co += (PB_THREADS / 32) * 4) {
    float v[4][2];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int e = lane + 32 * h;
        v[u][h] = (co + u < nco && e < run) ? __ldg(src0 + (co + u) * co_pitch + e) : 0.f;
      }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int e = lane + 32 * h;
        if (co + u < nco && e < run) tile[(co + u) * ld + e] = v[u][h];
      }
    for (int e = lane + 64; e < run; e += 32)            // taps > 16 only
      for (int u = 0; u < 4 && co + u < nco; ++u) tile[(co + u) * ld + e] = __ldg(src0 + (co + u) * co_pitch + e);
  }
  __syncthreads();
  T* out = reinterpret_cast<T*>(it.dst);
  for (int ci = 0; ci < nci; ++ci)
    for (int tap = 0; tap < taps; ++tap) {
      T* orow = out + ((size_t)(ci0 + ci) * taps + (taps - 1 - tap)) * it.dst_cols + it.col_off + co0;
      const int e = ci * taps + tap;
      for (int co = threadIdx.x; co < nco; co += PB_THREADS) orow[co] = from_f<T>(tile[co * ld + e]);
    }
}

template <typename T>
__device__ __forceinline__ void pack_up2x_row(const rcfd_pack_item& it, int co, float* row) {
  const int cin = it.cin, n = cin * 9;
  const float* __restrict__ src = it.src + (size_t)co * n;
  load_run(src, row, n, threadIdx.x, PB_THREADS);
  __syncthreads();
  T* out = reinterpret_cast<T*>(it.dst);
  for (int q = 0; q < 16; ++q) {
    const int tap = q & 3, ph = q >> 2;
    const int a = ph >> 1, b = ph & 1, t = tap >> 1, u = tap & 1;
    const int r0 = a == 0 ? (t == 0 ? 0 : 1) : (t == 0 ? 0 : 2), r1 = a == 0 ? (t == 0 ? 0 : 2) : (t == 0 ? 1 : 2);
    const int s0 = b == 0 ? (u == 0 ? 0 : 1) : (u == 0 ? 0 : 2), s1 = b == 0 ? (u == 0 ? 0 : 2) : (u == 0 ? 1 : 2);
    T* orow = out + (((size_t)ph * it.cout + co) * 4 + tap) * cin;
    for (int ci = threadIdx.x; ci < cin; ci += PB_THREADS) {
      float acc = 0.f;                        // same summation order as the element-wise kernel
      for (int rr = r0; rr <= r1; ++rr)
        for (int ss = s0; ss <= s1; ++ss) acc += row[ci * 9 + rr * 3 + ss];
      orow[ci] = from_f<T>(acc);
    }
  }
}

__device__ __forceinline__ void unpack_conv_row(const rcfd_pack_item& it, int co, float* row) {
  const int taps = it.taps, n = taps * it.cpad;
  const float* __restrict__ src = it.src + (size_t)co * n;                                // [tap][cpad]
  load_run(src, row, n, threadIdx.x, PB_THREADS);
  __syncthreads();
  float* g = reinterpret_cast<float*>(it.dst) + ((size_t)co * it.cin + it.cin_off) * taps;  // [ci][tap]
  const int m = it.cin_cnt * taps;
  for (int j = threadIdx.x; j < m; j += PB_THREADS) {
    const int ci = div_taps(j, taps), tap = j - ci * taps;
    g[j] = row[tap * it.cpad + ci];
  }
}

__global__ void __launch_bounds__(PB_THREADS) pack_batch_kernel(const rcfd_pack_item* __restrict__ items,
                                                                const int32_t* __restrict__ block_item, int n) {
  __shared__ rcfd_pack_item s_it;
  __shared__ __align__(16) float s_row[ROW_FLOATS];
  __shared__ int s_idx;
  static_assert(sizeof(rcfd_pack_item) % 4 == 0, "item copied as words");
  if (block_item != nullptr) {
    // block -> item map: one load, then the item's words by as many threads (no dependent-load chain per block)
    const int idx = __ldg(block_item + blockIdx.x);
    if (threadIdx.x < sizeof(rcfd_pack_item) / 4)
      reinterpret_cast<uint32_t*>(&s_it)[threadIdx.x] = __ldg(reinterpret_cast<const uint32_t*>(items + idx) + threadIdx.x);
  } else {
    if (threadIdx.x == 0) {
      int lo = 0, hi = n - 1;                    // last item with block0 <= blockIdx.x
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (items[mid].block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
      }
      s_idx = lo;
    }
    __syncthreads();
    if (threadIdx.x < sizeof(rcfd_pack_item) / 4)
      reinterpret_cast<uint32_t*>(&s_it)[threadIdx.x] = __ldg(reinterpret_cast<const uint32_t*>(items + s_idx) + threadIdx.x);
  }
  __syncthreads();
  const rcfd_pack_item& it = s_it;
  const int blk = (int)blockIdx.x - it.block0;
  const bool bf = it.dtype == RCFD_BF16;
  if (row_path(it)) {
    switch (it.kind) {
      case RCFD_PACK_FWD:
        if (bf) pack_fwd_row<bf16>(it, blk, s_row); else pack_fwd_row<float>(it, blk, s_row);
        break;
      case RCFD_PACK_DGRAD:
        if (bf) pack_dgrad_tile<bf16>(it, blk, s_row); else pack_dgrad_tile<float>(it, blk, s_row);
        break;
      case RCFD_PACK_UP2X:
        if (bf) pack_up2x_row<bf16>(it, blk, s_row); else pack_up2x_row<float>(it, blk, s_row);
        break;
      default:
        unpack_conv_row(it, blk, s_row);
    }
    return;
  }
  const uint32_t beg = (uint32_t)blk * RCFD_PACK_BLOCK_ELEMS;
  uint32_t end = beg + RCFD_PACK_BLOCK_ELEMS;
  if (end > (uint32_t)it.total) end = (uint32_t)it.total;
  if (it.kind >= RCFD_UNPACK_CONV && it.kind <= RCFD_COPY_F32) {
    for (uint32_t i = beg + threadIdx.x; i < end; i += PB_THREADS) unpack_one(it, i);
  } else if (bf) {
    for (uint32_t i = beg + threadIdx.x; i < end; i += PB_THREADS) pack_one<bf16>(it, i);
  } else {
    for (uint32_t i = beg + threadIdx.x; i < end; i += PB_THREADS) pack_one<float>(it, i);
  }
}

}  // namespace
}  // namespace rcfd

using namespace rcfd;

extern "C" {

int32_t rcfd_pack_item_blocks(const rcfd_pack_item* item) {
  if (item == nullptr || item->total <= 0 || item->total >= ((int64_t)1 << 31)) return 0;
  return item_blocks(*item);
}

int rcfd_pack_batch(const rcfd_pack_item* items, const int32_t* block_item, int32_t n, int32_t total_blocks, void* stream) {
  RCFD_CHECK_ARG(items != nullptr && n > 0 && total_blocks > 0, "pack_batch: bad args");
  pack_batch_kernel<<<total_blocks, PB_THREADS, 0, (cudaStream_t)stream>>>(items, block_item, n);
  RCFD_CHECK_LAUNCH("pack_batch");
  return RCFD_OK;
}

}  // extern "C"
