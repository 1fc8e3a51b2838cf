// Data-path kernels around the networks: batched on-device augmentation, 16-bit PNG
// depth / response decoding + random crop, bilinear up-sampling and the smoothness losses.  All HBM-bound passes
// over NCHW float tensors (the reference's layout at these call sites); one thread per output element, coalesced
// along x.
#include "common.cuh"

namespace rcfd {
namespace {

constexpr int NT = 256;

inline int blocks_for(int64_t work, int cap = 148 * 16) {
  int64_t g = (work + NT - 1) / NT;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------------------------------------ augmentation
// Per-sample parameters written by the host shim from the reference's torch.rand draws (src/fusionnet_transforms.py:76-165):
// [do_b, f_b, 1-f_b, do_c, f_c, 1-f_c, do_s, f_s, 1-f_s, hflip, vflip]
constexpr int XP = 11;

// max over a float tensor of non-negative-or-not values, into *out as an order-preserving int (no host sync)
__device__ __forceinline__ int float_orderable(float v) {
  const int i = __float_as_int(v);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__global__ void xform_max_kernel(const float* __restrict__ x, int* __restrict__ out, int64_t n) {
  float m = -INFINITY;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, x[i]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, float_orderable(m));
}

// torchvision `_blend` on one value (src/fusionnet_transforms.py:91-123 -> torchvision adjust_*): ratio * img + (1 - ratio) * other
// with the two products rounded separately, clamped to the dtype's bound; int32 images truncate toward zero after every op.
__device__ __forceinline__ float blend(float f, float omf, float v, float other, bool is_int) {
  float r = __fadd_rn(__fmul_rn(f, v), __fmul_rn(omf, other));
  r = fminf(fmaxf(r, 0.f), is_int ? 2147483647.f : 1.f);
  return is_int ? truncf(r) : r;
}
__device__ __forceinline__ float gray_of(float r, float g, float b, bool is_int) {
  const float y = __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
  return is_int ? truncf(y) : y;           // (...).to(img.dtype)
}

// per-sample sum of the grey image AFTER the brightness step (the contrast blend partner is its mean)
__global__ void xform_gray_sum_kernel(const float* __restrict__ img, const float* __restrict__ prm, const int* __restrict__ maxbits,
                                      double* __restrict__ sums, int N, int HW) {
  const bool is_int = *maxbits > __float_as_int(1.0f);
  const int n = blockIdx.y;
  const float* p = prm + n * XP;
  const bool do_b = p[0] != 0.f;
  double acc = 0.0;
  const float* base = img + (size_t)n * 3 * HW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    float c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float v = base[(size_t)k * HW + i];
      if (is_int) v = truncf(v);
      if (do_b) v = blend(p[1], p[2], v, 0.f, is_int);
      c[k] = v;
    }
    acc += (double)gray_of(c[0], c[1], c[2], is_int);
  }
  __shared__ double red[NT / 32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < NT / 32; ++w) t += red[w];
    atomicAdd(sums + n, t);
  }
  (void)N;
}

struct XformMaps {
  const float* src[4];
  float* dst[4];
  int ch[4];
  int count;
};

__global__ void xform_apply_kernel(const float* __restrict__ img, float* __restrict__ out, const float* __restrict__ prm,
                                   const int* __restrict__ maxbits, const double* __restrict__ sums, XformMaps maps, int N, int H,
                                   int W, int norm_mode) {
  const bool is_int = img != nullptr && *maxbits > __float_as_int(1.0f);
  const int HW = H * W;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < (int64_t)N * HW; idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / HW);
    const int pix = (int)(idx - (int64_t)n * HW);
    const int y = pix / W, x = pix - y * W;
    const float* p = prm + n * XP;
    const int sy = p[10] != 0.f ? H - 1 - y : y;
    const int sx = p[9] != 0.f ? W - 1 - x : x;
    const int sp = sy * W + sx;
    if (img != nullptr) {
      const float* base = img + (size_t)n * 3 * HW;
      float c[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float v = base[(size_t)k * HW + sp];
        if (is_int) v = truncf(v);
        if (p[0] != 0.f) v = blend(p[1], p[2], v, 0.f, is_int);
        c[k] = v;
      }
      if (p[3] != 0.f) {
        const float mean = (float)(sums[n] / (double)HW);
#pragma unroll
        for (int k = 0; k < 3; ++k) c[k] = blend(p[4], p[5], c[k], mean, is_int);
      }
      if (p[6] != 0.f) {
        const float g = gray_of(c[0], c[1], c[2], is_int);
#pragma unroll
        for (int k = 0; k < 3; ++k) c[k] = blend(p[7], p[8], c[k], g, is_int);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float v = c[k];
        if (norm_mode == 1) v = __fdiv_rn(v, 255.f);
        else if (norm_mode == 2) v = __fsub_rn(__fmul_rn(2.f, __fdiv_rn(v, 255.f)), 1.f);
        out[((size_t)n * 3 + k) * HW + pix] = v;
      }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      if (m < maps.count) {
        for (int k = 0; k < maps.ch[m]; ++k)
          maps.dst[m][((size_t)n * maps.ch[m] + k) * HW + pix] = maps.src[m][((size_t)n * maps.ch[m] + k) * HW + sp];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ 16-bit PNG value codec + crop
// What the reference's data path does between the decoded PNG raster and the network input (src/data_utils.py:238-269,
// 288-318: np.array(Image.open(path), float32) / multiplier, depth <= 0 -> 0; src/datasets.py:19-109: crop every tensor of
// the sample at the same (y0, x0); load_image: HWC uint8 -> CHW float): here for the whole batch, from the on-disk sample
// types (uint16 rasters, uint8 HWC images), so the host ships 2 / 1 bytes per value instead of 4.
template <typename S>
__global__ void decode_crop_kernel(const S* __restrict__ src, float* __restrict__ dst, const int* __restrict__ crop_yx, int N,
                                   int SH, int SW, int C, int OH, int OW, float multiplier, int64_t dst_batch_stride) {
  const int64_t total = (int64_t)N * C * OH * OW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % OW);
    int64_t r = i / OW;
    const int y = (int)(r % OH); r /= OH;
    const int c = (int)(r % C);
    const int n = (int)(r / C);
    const int y0 = crop_yx ? crop_yx[2 * n] : 0, x0 = crop_yx ? crop_yx[2 * n + 1] : 0;
    const S v = src[(((int64_t)n * SH + (y + y0)) * SW + (x + x0)) * C + c];          // HWC raster (C = 1 for the maps)
    float f = __fdiv_rn((float)v, multiplier);
    if (f <= 0.f) f = 0.f;
    dst[(int64_t)n * dst_batch_stride + ((int64_t)c * OH + y) * OW + x] = f;
  }
}

// save_depth / save_response (src/data_utils.py:271-286, 320-335): np.uint32(v * multiplier) stored as a 16-bit PNG sample
__global__ void encode_u16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, float multiplier, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = __fmul_rn(src[i], multiplier);
    const unsigned int u = v > 0.f ? (unsigned int)fminf(v, 4294967040.f) : 0u;       // C cast: truncation toward zero
    dst[i] = (uint16_t)(u & 0xffffu);
  }
}

}  // namespace
}  // namespace rcfd

using namespace rcfd;

extern "C" {

int rcfd_decode_crop(const void* src, int32_t src_bits, float* dst, const int32_t* crop_yx, int32_t n, int32_t src_h,
                     int32_t src_w, int32_t channels, int32_t out_h, int32_t out_w, float multiplier,
                     int64_t dst_batch_stride, void* stream) {
  RCFD_CHECK_ARG(src && dst && n > 0 && channels > 0 && out_h > 0 && out_w > 0 && src_h >= out_h && src_w >= out_w,
                 "decode_crop: bad args");
  RCFD_CHECK_ARG((src_bits == 8 || src_bits == 16) && multiplier > 0.f, "decode_crop: src_bits 8 | 16, multiplier > 0");
  RCFD_CHECK_ARG(dst_batch_stride >= (int64_t)channels * out_h * out_w, "decode_crop: dst_batch_stride");
  const int64_t total = (int64_t)n * channels * out_h * out_w;
  if (src_bits == 8)
    decode_crop_kernel<uint8_t><<<blocks_for(total), NT, 0, (cudaStream_t)stream>>>((const uint8_t*)src, dst, crop_yx, n, src_h, src_w,
                                                                                   channels, out_h, out_w, multiplier, dst_batch_stride);
  else
    decode_crop_kernel<uint16_t><<<blocks_for(total), NT, 0, (cudaStream_t)stream>>>((const uint16_t*)src, dst, crop_yx, n, src_h, src_w,
                                                                                    channels, out_h, out_w, multiplier, dst_batch_stride);
  RCFD_CHECK_LAUNCH("decode_crop");
  return RCFD_OK;
}

int rcfd_encode_u16(const float* src, uint16_t* dst, float multiplier, int64_t count, void* stream) {
  RCFD_CHECK_ARG(src && dst && count > 0 && multiplier > 0.f, "encode_u16: bad args");
  encode_u16_kernel<<<blocks_for(count), NT, 0, (cudaStream_t)stream>>>(src, dst, multiplier, count);
  RCFD_CHECK_LAUNCH("encode_u16");
  return RCFD_OK;
}

int rcfd_transform_batch(const float* image, float* image_out, const float* const* maps, float* const* maps_out,
                         const int32_t* map_channels, int32_t n_maps, const float* params, int32_t* scratch_max,
                         double* scratch_sums, int32_t n, int32_t h, int32_t w, int32_t norm_mode, void* stream) {
  RCFD_CHECK_ARG(params && n > 0 && h > 0 && w > 0 && n_maps >= 0 && n_maps <= 4, "transform_batch: bad args (<= 4 range maps per call)");
  RCFD_CHECK_ARG((image == nullptr) == (image_out == nullptr), "transform_batch: image / image_out");
  RCFD_CHECK_ARG(image == nullptr || (scratch_max && scratch_sums), "transform_batch: scratch");
  RCFD_CHECK_ARG(norm_mode >= 0 && norm_mode <= 2, "transform_batch: norm_mode 0 ([0,255]) | 1 ([0,1]) | 2 ([-1,1])");
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = h * w;
  if (image) {
    cudaError_t e = cudaMemsetAsync(scratch_max, 0x80, sizeof(int32_t), st);          // 0x80808080: below every orderable float
    if (e == cudaSuccess) e = cudaMemsetAsync(scratch_sums, 0, sizeof(double) * n, st);
    if (e != cudaSuccess) { set_error("transform_batch memset: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
    xform_max_kernel<<<blocks_for((int64_t)n * 3 * HW, 148 * 8), NT, 0, st>>>(image, scratch_max, (int64_t)n * 3 * HW);
    RCFD_CHECK_LAUNCH("transform max");
    dim3 grid(blocks_for(HW, 64), n);
    xform_gray_sum_kernel<<<grid, NT, 0, st>>>(image, params, scratch_max, scratch_sums, n, HW);
    RCFD_CHECK_LAUNCH("transform gray sum");
  }
  XformMaps m;
  m.count = n_maps;
  for (int i = 0; i < 4; ++i) {
    m.src[i] = i < n_maps ? maps[i] : nullptr;
    m.dst[i] = i < n_maps ? maps_out[i] : nullptr;
    m.ch[i] = i < n_maps ? map_channels[i] : 0;
    RCFD_CHECK_ARG(i >= n_maps || (m.src[i] && m.dst[i] && m.ch[i] > 0), "transform_batch: null range map");
  }
  xform_apply_kernel<<<blocks_for((int64_t)n * HW), NT, 0, st>>>(image, image_out, params, scratch_max, scratch_sums, m, n, h, w,
                                                                  norm_mode);
  RCFD_CHECK_LAUNCH("transform apply");
  return RCFD_OK;
}

}  // extern "C"
