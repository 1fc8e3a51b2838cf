// Glue of the multi-resolution decoder (reference src/networks.py:1595-1642, n_resolution > 1): the 1-channel logits of a
// coarser output are up-sampled 2x bilinearly (align_corners=True) and concatenated behind the next block's skip.
// Float 1-channel maps [N, H, W] for the logits; the concatenated tensor is NHWC in the compute dtype with the logit
// channel followed by zero channels up to a multiple of 16 (TMA boxes / UMMA K need 32-byte rows).
#include "common.cuh"

namespace rcfd {
namespace {

constexpr int MT = 256;

// ATen upsample_bilinear2d, align_corners=True: src = dst * (in - 1) / (out - 1)
__device__ __forceinline__ void bil_coord(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  const float s = scale * (float)dst;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 < in_size - 1 ? i0 + 1 : i0;
  l1 = s - (float)i0;
}

__global__ void bilinear2x_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int HO, int WO,
                                      float sh, float sw) {
  const int64_t total = (int64_t)N * HO * WO;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % WO), oy = (int)((i / WO) % HO), n = (int)(i / ((int64_t)HO * WO));
    int y0, y1, x0, x1; float ly, lx;
    bil_coord(oy, sh, H, y0, y1, ly);
    bil_coord(ox, sw, W, x0, x1, lx);
    const float* p = x + (size_t)n * H * W;
    // same association as ATen: h0lambda * (w0lambda * a + w1lambda * b) + h1lambda * (w0lambda * c + w1lambda * d)
    y[i] = (1.f - ly) * ((1.f - lx) * p[(size_t)y0 * W + x0] + lx * p[(size_t)y0 * W + x1]) +
           ly * ((1.f - lx) * p[(size_t)y1 * W + x0] + lx * p[(size_t)y1 * W + x1]);
  }
}

// gradient: every source pixel gathers from the (at most 3 x 3) destination pixels whose 2 x 2 footprint contains it
__global__ void bilinear2x_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int N, int H, int W, int HO, int WO,
                                      float sh, float sw) {
  const int64_t total = (int64_t)N * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), n = (int)(i / ((int64_t)H * W));
    const float* g = dy + (size_t)n * HO * WO;
    // destination rows whose source coordinate lies in (y - 1, y + 1): a superset is [2y - 2, 2y + 2]
    float acc = 0.f;
    for (int oy = max(0, 2 * y - 2); oy <= min(HO - 1, 2 * y + 2); ++oy) {
      int y0, y1; float ly;
      bil_coord(oy, sh, H, y0, y1, ly);
      float wy = 0.f;
      if (y0 == y) wy += 1.f - ly;
      if (y1 == y) wy += ly;
      if (wy == 0.f) continue;
      for (int ox = max(0, 2 * x - 2); ox <= min(WO - 1, 2 * x + 2); ++ox) {
        int x0, x1; float lx;
        bil_coord(ox, sw, W, x0, x1, lx);
        float wx = 0.f;
        if (x0 == x) wx += 1.f - lx;
        if (x1 == x) wx += lx;
        if (wx != 0.f) acc += wy * wx * g[(size_t)oy * WO + ox];
      }
    }
    dx[i] = acc;
  }
}

template <typename T>
__global__ void concat_logit_kernel(const T* __restrict__ skip, const float* __restrict__ logit, T* __restrict__ out,
                                    int64_t pixels, int C, int CO) {
  const int64_t total = pixels * CO;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % CO);
    const int64_t p = i / CO;
    out[i] = c < C ? skip[p * C + c] : (c == C ? from_f<T>(logit[p]) : from_f<T>(0.f));
  }
}

template <typename T>
__global__ void split_logit_kernel(const T* __restrict__ dcat, T* __restrict__ dskip, float* __restrict__ dlogit, int64_t pixels,
                                   int C, int CO) {
  const int64_t total = pixels * (C + 1);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % (C + 1));
    const int64_t p = i / (C + 1);
    const T v = dcat[p * CO + c];
    if (c < C) dskip[p * C + c] = v; else dlogit[p] = to_f<T>(v);
  }
}

inline int grid_of(int64_t work) {
  int64_t b = (work + MT - 1) / MT;
  return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

}  // namespace
}  // namespace rcfd

using namespace rcfd;

extern "C" {

int rcfd_bilinear2x_fwd(const float* x, float* y, int32_t n, int32_t h, int32_t w, void* stream) {
  RCFD_CHECK_ARG(x && y && n > 0 && h > 0 && w > 0, "bilinear2x_fwd: bad args");
  const int ho = 2 * h, wo = 2 * w;
  const float sh = ho > 1 ? (float)(h - 1) / (float)(ho - 1) : 0.f, sw = wo > 1 ? (float)(w - 1) / (float)(wo - 1) : 0.f;
  bilinear2x_fwd_kernel<<<grid_of((int64_t)n * ho * wo), MT, 0, (cudaStream_t)stream>>>(x, y, n, h, w, ho, wo, sh, sw);
  RCFD_CHECK_LAUNCH("bilinear2x_fwd");
  return RCFD_OK;
}

int rcfd_bilinear2x_bwd(const float* dy, float* dx, int32_t n, int32_t h, int32_t w, void* stream) {
  RCFD_CHECK_ARG(dy && dx && n > 0 && h > 0 && w > 0, "bilinear2x_bwd: bad args");
  const int ho = 2 * h, wo = 2 * w;
  const float sh = ho > 1 ? (float)(h - 1) / (float)(ho - 1) : 0.f, sw = wo > 1 ? (float)(w - 1) / (float)(wo - 1) : 0.f;
  bilinear2x_bwd_kernel<<<grid_of((int64_t)n * h * w), MT, 0, (cudaStream_t)stream>>>(dy, dx, n, h, w, ho, wo, sh, sw);
  RCFD_CHECK_LAUNCH("bilinear2x_bwd");
  return RCFD_OK;
}

int rcfd_concat_logit(const void* skip, const float* logit, void* out, int64_t pixels, int32_t c, int32_t c_out, int32_t dtype,
                      void* stream) {
  RCFD_CHECK_ARG(logit && out && pixels > 0 && c >= 0 && c_out > c && (c == 0 || skip), "concat_logit: bad args");
  if (dtype == RCFD_F32) {
    concat_logit_kernel<float><<<grid_of(pixels * c_out), MT, 0, (cudaStream_t)stream>>>((const float*)skip, logit, (float*)out, pixels, c, c_out);
  } else if (dtype == RCFD_BF16) {
    concat_logit_kernel<bf16><<<grid_of(pixels * c_out), MT, 0, (cudaStream_t)stream>>>((const bf16*)skip, logit, (bf16*)out, pixels, c, c_out);
  } else { set_error("concat_logit: bad dtype"); return RCFD_EINVAL; }
  RCFD_CHECK_LAUNCH("concat_logit");
  return RCFD_OK;
}

int rcfd_split_logit(const void* dcat, void* dskip, float* dlogit, int64_t pixels, int32_t c, int32_t c_out, int32_t dtype,
                     void* stream) {
  RCFD_CHECK_ARG(dcat && dlogit && pixels > 0 && c >= 0 && c_out > c && (c == 0 || dskip), "split_logit: bad args");
  if (dtype == RCFD_F32) {
    split_logit_kernel<float><<<grid_of(pixels * (c + 1)), MT, 0, (cudaStream_t)stream>>>((const float*)dcat, (float*)dskip, dlogit, pixels, c, c_out);
  } else if (dtype == RCFD_BF16) {
    split_logit_kernel<bf16><<<grid_of(pixels * (c + 1)), MT, 0, (cudaStream_t)stream>>>((const bf16*)dcat, (bf16*)dskip, dlogit, pixels, c, c_out);
  } else { set_error("split_logit: bad dtype"); return RCFD_EINVAL; }
  RCFD_CHECK_LAUNCH("split_logit");
  return RCFD_OK;
}

}  // extern "C"
