// Every weight packing (OIHW float -> the kernels' packed layouts) and every weight-gradient unpacking of a
// training step in ONE launch each: a device table of items, RCFD_PACK_BLOCK_ELEMS destination elements per
// block, block -> item by binary search over the items' first block.  Same index maps as the single-tensor
// kernels of elementwise.cu (rcfd_pack_conv_weight & co), which stay for inference / one-off calls.
#include "common.cuh"

namespace rcfd {
namespace {

constexpr int PB_THREADS = 256;

template <typename T>
__device__ __forceinline__ void pack_one(const rcfd_pack_item& it, int64_t i) {
  T* out = reinterpret_cast<T*>(it.dst);
  const float* __restrict__ w = it.src;
  const int taps = it.taps;
  switch (it.kind) {
    case RCFD_PACK_FWD: {                      // out[co][tap][ci < cpad]
      const int ci = (int)(i % it.cpad);
      const int64_t r = i / it.cpad;
      const int tap = (int)(r % taps), co = (int)(r / taps);
      out[i] = from_f<T>(ci < it.cin_cnt ? __ldg(w + ((size_t)co * it.cin + it.cin_off + ci) * taps + tap) : 0.f);
      break;
    }
    case RCFD_PACK_DGRAD: {                    // out[ci][flipped tap][col_off + co], rows dst_cols wide
      const int co = (int)(i % it.cout);
      const int64_t r = i / it.cout;
      const int tap = (int)(r % taps), ci = (int)(r / taps);
      out[((size_t)ci * taps + tap) * it.dst_cols + it.col_off + co] =
          from_f<T>(__ldg(w + ((size_t)co * it.cin + it.cin_off + ci) * taps + (taps - 1 - tap)));
      break;
    }
    case RCFD_PACK_UP2X: {                     // out[phase][co][2x2 tap][ci]: sums of the 3x3 taps hitting one low-res pixel
      const int cin = it.cin, cout = it.cout;
      const int ci = (int)(i % cin);
      int64_t r = i / cin;
      const int tap = (int)(r % 4); r /= 4;
      const int co = (int)(r % cout);
      const int ph = (int)(r / cout);
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
    default: {                                 // RCFD_PACK_STEM_S2D: out[co][4x4 tap][cpad], 7x7 window on the s2d tensor
      const int C = it.cin, CP = it.cpad;
      const int ch = (int)(i % CP);
      const int64_t r = i / CP;
      const int tap = (int)(r % 16), co = (int)(r / 16);
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

__device__ __forceinline__ void unpack_one(const rcfd_pack_item& it, int64_t i) {
  float* g = reinterpret_cast<float*>(it.dst);
  const float* __restrict__ packed = it.src;
  if (it.kind == RCFD_UNPACK_CONV) {           // i runs over [co][ci < cin_cnt][tap]: the OIHW slice, writes coalesced
    const int taps = it.taps;
    const int tap = (int)(i % taps);
    const int64_t r = i / taps;
    const int ci = (int)(r % it.cin_cnt), co = (int)(r / it.cin_cnt);
    g[((size_t)co * it.cin + it.cin_off + ci) * taps + tap] = packed[((size_t)co * taps + tap) * it.cpad + ci];
  } else {                                     // RCFD_UNPACK_STEM_S2D
    const int C = it.cin, CP = it.cpad;
    const int ss = (int)(i % 7);
    int64_t r = i / 7;
    const int rr = (int)(r % 7); r /= 7;
    const int c = (int)(r % C), co = (int)(r / C);
    const int ty = (rr + 1) >> 1, dy = (rr + 1) & 1, tx = (ss + 1) >> 1, dx = (ss + 1) & 1;
    g[i] = packed[((size_t)co * 16 + ty * 4 + tx) * CP + (dy * 2 + dx) * C + c];
  }
}

__global__ void __launch_bounds__(PB_THREADS) pack_batch_kernel(const rcfd_pack_item* __restrict__ items, int n) {
  __shared__ rcfd_pack_item s_it;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n - 1;                    // last item with block0 <= blockIdx.x
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (items[mid].block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    s_it = items[lo];
  }
  __syncthreads();
  const rcfd_pack_item& it = s_it;
  const int64_t beg = (int64_t)((int)blockIdx.x - it.block0) * RCFD_PACK_BLOCK_ELEMS;
  int64_t end = beg + RCFD_PACK_BLOCK_ELEMS;
  if (end > it.total) end = it.total;
  if (it.kind >= RCFD_UNPACK_CONV) {
    for (int64_t i = beg + threadIdx.x; i < end; i += PB_THREADS) unpack_one(it, i);
  } else if (it.dtype == RCFD_BF16) {
    for (int64_t i = beg + threadIdx.x; i < end; i += PB_THREADS) pack_one<bf16>(it, i);
  } else {
    for (int64_t i = beg + threadIdx.x; i < end; i += PB_THREADS) pack_one<float>(it, i);
  }
}

}  // namespace
}  // namespace rcfd

using namespace rcfd;

extern "C" int rcfd_pack_batch(const rcfd_pack_item* items, int32_t n, int32_t total_blocks, void* stream) {
  RCFD_CHECK_ARG(items != nullptr && n > 0 && total_blocks > 0, "pack_batch: bad args");
  pack_batch_kernel<<<total_blocks, PB_THREADS, 0, (cudaStream_t)stream>>>(items, n);
  RCFD_CHECK_LAUNCH("pack_batch");
  return RCFD_OK;
}
