// Support passes of the tensor-core PARITY modes (fp32 storage; every convolution = several tcgen05
// passes over bf16 operand splits, fp32 accumulation in TMEM and in the fp32 output):
//
//   x = x0 + x1 + x2 (+ O(2^-25 |x|)),   x0 = bf16(x), x1 = bf16(x - x0), x2 = bf16(x - x0 - x1)
//   "bf16x3":  x.w ~= x0.w0 + x1.w0 + x0.w1                                  (~2^-16 relative per product)
//   "bf16x6":  x.w ~= x0.w0 + x1.w0 + x0.w1 + x1.w1 + x2.w0 + x0.w2          (~2^-23: fp32-class)
//
// so the SAME tcgen05 / TMA kernels that run the bf16 fast path reproduce the reference's fp32
// convolution (src/net_utils.py:63-69, 85).  This file holds the three HBM-bound passes
// around those launches: the operand split, per-channel BatchNorm statistics of an fp32 tensor, and
// the fp32 epilogue (folded BN / activation / depth head / residual) that the fast path fuses into
// the conv kernels.
#include "common.cuh"

namespace rcfd {
namespace {

constexpr int NT = 256;

// x -> p0 + p1 (+ p2): p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1); the subtractions are exact in fp32
__device__ __forceinline__ void split3(float x, bf16& a, bf16& b, bf16& c) {
  a = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(a);
  b = __float2bfloat16_rn(r1);
  c = __float2bfloat16_rn(r1 - __bfloat162float(b));
}

__global__ void split_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ p0, bf16* __restrict__ p1,
                                  bf16* __restrict__ p2, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(x + i * 4);
    const float f[4] = {v.x, v.y, v.z, v.w};
    bf16 a[4], b[4], c[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split3(f[j], a[j], b[j], c[j]);
    *reinterpret_cast<uint2*>(p0 + i * 4) = *reinterpret_cast<const uint2*>(a);
    *reinterpret_cast<uint2*>(p1 + i * 4) = *reinterpret_cast<const uint2*>(b);
    if (p2) *reinterpret_cast<uint2*>(p2 + i * 4) = *reinterpret_cast<const uint2*>(c);
  }
}

__global__ void split_bf16_tail_kernel(const float* __restrict__ x, bf16* __restrict__ p0, bf16* __restrict__ p1,
                                       bf16* __restrict__ p2, int64_t beg, int64_t n) {
  const int64_t i = beg + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    bf16 a, b, c;
    split3(x[i], a, b, c);
    p0[i] = a;
    p1[i] = b;
    if (p2) p2[i] = c;
  }
}

// per-channel sum / sum of squares of a [pixels][C] fp32 tensor: thread t owns channel (t % CP) and walks pixel
// rows t / CP, t / CP + NT / CP, ...; fp32 partial sums per thread (<= a few thousand terms), fp64 from there on.
__global__ void channel_stats_kernel(const float* __restrict__ y, double* __restrict__ ssum, double* __restrict__ ssq,
                                     int64_t pixels, int C, int CP, int64_t rows_per_block) {
  __shared__ double red[2][NT];
  const int c = threadIdx.x % CP, r0 = threadIdx.x / CP, rstep = NT / CP;
  const int64_t pbeg = blockIdx.x * rows_per_block;
  int64_t pend = pbeg + rows_per_block;
  if (pend > pixels) pend = pixels;
  for (int cb = 0; cb < C; cb += CP) {
    double s = 0.0, q = 0.0;
    if (cb + c < C) {
      float fs = 0.f, fq = 0.f;
      int cnt = 0;
      for (int64_t p = pbeg + r0; p < pend; p += rstep) {
        const float v = y[p * C + cb + c];
        fs += v;
        fq = fmaf(v, v, fq);
        if (++cnt == 256) { s += (double)fs; q += (double)fq; fs = 0.f; fq = 0.f; cnt = 0; }
      }
      s += (double)fs;
      q += (double)fq;
    }
    red[0][threadIdx.x] = s;
    red[1][threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.x < CP && cb + threadIdx.x < C) {
      double a = 0.0, b = 0.0;
      for (int r = 0; r < rstep; ++r) {
        a += red[0][r * CP + threadIdx.x];
        b += red[1][r * CP + threadIdx.x];
      }
      atomicAdd(ssum + cb + threadIdx.x, a);
      atomicAdd(ssq + cb + threadIdx.x, b);
    }
    __syncthreads();
  }
}

__global__ void epilogue_f32_kernel(const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ res, float* __restrict__ out, int64_t n, int C, int act, float p0,
                                    float p1) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float v = y[i];
    if (scale) v = fmaf(v, __ldg(scale + c), __ldg(shift + c));
    v = apply_act(v, act, p0, p1);
    if (res) v = leaky(v + res[i]);
    out[i] = v;
  }
}

inline int blocks_for(int64_t work, int cap = 148 * 16) {
  int64_t g = (work + NT - 1) / NT;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace
}  // namespace rcfd

using namespace rcfd;

extern "C" {

int rcfd_split_bf16(const float* x, void* p0, void* p1, void* p2, int64_t count, void* stream) {
  RCFD_CHECK_ARG(x && p0 && p1 && count > 0, "split_bf16: bad args");
  const int64_t n4 = count / 4;
  if (n4 > 0) {
    RCFD_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) & 15) | (reinterpret_cast<uintptr_t>(p0) & 7) |
                    (reinterpret_cast<uintptr_t>(p1) & 7) | (reinterpret_cast<uintptr_t>(p2) & 7)) == 0, "split_bf16: unaligned");
    split_bf16_kernel<<<blocks_for(n4), NT, 0, (cudaStream_t)stream>>>(x, (bf16*)p0, (bf16*)p1, (bf16*)p2, n4);
    RCFD_CHECK_LAUNCH("split_bf16");
  }
  if (n4 * 4 < count) {
    split_bf16_tail_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(x, (bf16*)p0, (bf16*)p1, (bf16*)p2, n4 * 4, count);
    RCFD_CHECK_LAUNCH("split_bf16 tail");
  }
  return RCFD_OK;
}

int rcfd_channel_stats(const float* y, double* stats_sum, double* stats_sqsum, int64_t pixels, int32_t channels,
                       void* stream) {
  RCFD_CHECK_ARG(y && stats_sum && stats_sqsum && pixels > 0 && channels > 0, "channel_stats: bad args");
  int cp = 1;
  while (cp < channels && cp < NT) cp <<= 1;
  const int rstep = NT / cp;
  int blocks = (int)((pixels + rstep * 16 - 1) / (rstep * 16));
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  const int64_t rpb = (pixels + blocks - 1) / blocks;
  blocks = (int)((pixels + rpb - 1) / rpb);
  channel_stats_kernel<<<blocks, NT, 0, (cudaStream_t)stream>>>(y, stats_sum, stats_sqsum, pixels, channels, cp, rpb);
  RCFD_CHECK_LAUNCH("channel_stats");
  return RCFD_OK;
}

int rcfd_epilogue_f32(const float* y, const float* scale, const float* shift, const float* residual, float* out,
                      int64_t pixels, int32_t channels, int32_t act, float act_p0, float act_p1, void* stream) {
  RCFD_CHECK_ARG(y && out && pixels > 0 && channels > 0, "epilogue_f32: bad args");
  RCFD_CHECK_ARG((scale == nullptr) == (shift == nullptr), "epilogue_f32: scale/shift");
  const int64_t n = pixels * channels;
  epilogue_f32_kernel<<<blocks_for(n), NT, 0, (cudaStream_t)stream>>>(y, scale, shift, residual, out, n, channels, act,
                                                                      act_p0, act_p1);
  RCFD_CHECK_LAUNCH("epilogue_f32");
  return RCFD_OK;
}

}  // extern "C"
