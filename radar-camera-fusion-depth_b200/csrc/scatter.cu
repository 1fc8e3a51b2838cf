// Radar point -> pixel scatters, the RadarNet ROI pooling and its point MLP.
// Integer index work is bit exact with the reference (see include/rcfd.h for citations).
#include "common.cuh"

namespace rcfd {
namespace {

// ---- S1: np.round (half-to-even) -> ordered img[y, x] = z ; one thread per point.
// Sequential "last writer wins" == point i writes iff no later point lands on its pixel.
__global__ void s1_plot_kernel(const double* __restrict__ pts, const double* __restrict__ depth, int npts,
                               double* __restrict__ img, int H, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts) return;
  const long long x = (long long)rint(pts[i]);            // rint: round-half-to-even like np.round
  const long long y = (long long)rint(pts[npts + i]);
  if (x < 0 || x >= W || y < 0 || y >= H) return;          // reference guarantees in-range (mask :195-200)
  for (int j = i + 1; j < npts; ++j) {
    if ((long long)rint(pts[j]) == x && (long long)rint(pts[npts + j]) == y) return;
  }
  img[(size_t)y * W + x] = depth[i];
}

// ---- S1 merge (z-buffer): sequentially "overwrite iff empty or closer" == the pixel ends with
// min(existing if > 0, all new depths); the point achieving that min (lowest index on ties) writes.
__global__ void s1_merge_kernel(const double* __restrict__ pts, const double* __restrict__ depth, int npts,
                                double* __restrict__ img, int H, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts) return;
  const long long x = (long long)rint(pts[i]);
  const long long y = (long long)rint(pts[npts + i]);
  if (x < 0 || x >= W || y < 0 || y >= H) return;
  const double z = depth[i];
  for (int j = 0; j < npts; ++j) {
    if (j == i) continue;
    if ((long long)rint(pts[j]) == x && (long long)rint(pts[npts + j]) == y) {
      const double zj = depth[j];
      if (zj < z || (zj == z && j < i)) return;
    }
  }
  const double cur = img[(size_t)y * W + x];
  if (!(cur > 0.0) || z < cur) img[(size_t)y * W + x] = z;
}

// ---- S2: one thread per output pixel; running max / first arg-max over the K pasted crops.
__global__ void s2_kernel(const float* __restrict__ crops, const float* __restrict__ points, int K, int ph, int pw,
                          int H, int W, int compat, long long* __restrict__ depth_i64,
                          float* __restrict__ depth_f32, float* __restrict__ response) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * W) return;
  const int y = idx / W, x = idx - y * W;
  const int pad = pw / 2;
  const int cy = y - (H - ph);                 // row inside the crop (crop is pasted at the bottom)
  const int xp = x + pad;                      // column in the padded canvas
  float best = 0.f;
  int arg = 0;
  for (int k = 0; k < K; ++k) {
    float v = 0.f;
    if (cy >= 0) {
      const int col = xp - ((int)points[k * 3] - pad);     // int() truncation, radarnet_main.py:568
      if (col >= 0 && col < pw) {
        v = crops[((size_t)k * ph + cy) * pw + col];
        if (v < 0.5f) v = 0.f;
      }
    }
    if (k == 0 || v > best) {
      if (k == 0 || v > best) { best = v; arg = k; }
    }
  }
  response[idx] = best;
  if (compat) {
    long long v = arg;
    for (int k = 0; k < K; ++k)
      if (v == k) v = (long long)points[k * 3 + 2];          // int64 fill truncates, later k re-match
    if (depth_i64) depth_i64[idx] = best == 0.f ? 0 : v;
  } else {
    if (depth_f32) depth_f32[idx] = best == 0.f ? 0.f : points[arg * 3 + 2];
  }
}

// ---- roi_pool (torchvision semantics), NHWC.  Thread per (box, bin, 16-byte channel vector): the bin geometry
// is computed once per vector and every load / store is a coalesced 16-byte access.
template <typename T>
__global__ void roi_pool_kernel(const T* __restrict__ feat, const float* __restrict__ boxes, T* __restrict__ out,
                                int N, int H, int W, int C, int nbox, int PH, int PW, float scale) {
  constexpr int NV = V16<T>::N;
  const int CV = C / NV;
  const int64_t total = (int64_t)nbox * PH * PW * CV;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV) * NV;
    int64_t r = i / CV;
    const int pw = (int)(r % PW); r /= PW;
    const int ph = (int)(r % PH);
    const int b = (int)(r / PH);
    const float* box = boxes + (size_t)b * 5;
    const int n = (int)box[0];
    const int sw = (int)roundf(box[1] * scale), sh = (int)roundf(box[2] * scale);
    const int ew = (int)roundf(box[3] * scale), eh = (int)roundf(box[4] * scale);
    const int rw = max(ew - sw + 1, 1), rh = max(eh - sh + 1, 1);
    const float bh = (float)rh / (float)PH, bw = (float)rw / (float)PW;
    int hs = (int)floorf((float)ph * bh), he = (int)ceilf((float)(ph + 1) * bh);
    int ws = (int)floorf((float)pw * bw), we = (int)ceilf((float)(pw + 1) * bw);
    hs = min(max(hs + sh, 0), H); he = min(max(he + sh, 0), H);
    ws = min(max(ws + sw, 0), W); we = min(max(we + sw, 0), W);
    const bool empty = (he <= hs) || (we <= ws);
    float m[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) m[k] = empty ? 0.f : -3.402823466e+38f;
    if (n >= 0 && n < N) {
      for (int yy = hs; yy < he; ++yy)
        for (int xx = ws; xx < we; ++xx) {
          float v[NV];
          V16<T>::ld(feat + ((size_t)(n * H + yy) * W + xx) * C + c, v);
#pragma unroll
          for (int k = 0; k < NV; ++k) m[k] = fmaxf(m[k], v[k]);
        }
    }
    V16<T>::st(out + i * NV, m);
  }
}
// any channel count: thread per (box, bin, channel)
template <typename T>
__global__ void roi_pool_scalar_kernel(const T* __restrict__ feat, const float* __restrict__ boxes, T* __restrict__ out,
                                       int N, int H, int W, int C, int nbox, int PH, int PW, float scale) {
  const int64_t total = (int64_t)nbox * PH * PW * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int pw = (int)(r % PW); r /= PW;
    const int ph = (int)(r % PH);
    const int b = (int)(r / PH);
    const float* box = boxes + (size_t)b * 5;
    const int n = (int)box[0];
    const int sw = (int)roundf(box[1] * scale), sh = (int)roundf(box[2] * scale);
    const int ew = (int)roundf(box[3] * scale), eh = (int)roundf(box[4] * scale);
    const int rw = max(ew - sw + 1, 1), rh = max(eh - sh + 1, 1);
    const float bh = (float)rh / (float)PH, bw = (float)rw / (float)PW;
    int hs = (int)floorf((float)ph * bh), he = (int)ceilf((float)(ph + 1) * bh);
    int ws = (int)floorf((float)pw * bw), we = (int)ceilf((float)(pw + 1) * bw);
    hs = min(max(hs + sh, 0), H); he = min(max(he + sh, 0), H);
    ws = min(max(ws + sw, 0), W); we = min(max(we + sw, 0), W);
    const bool empty = (he <= hs) || (we <= ws);
    float m = empty ? 0.f : -3.402823466e+38f;
    if (n >= 0 && n < N) {
      for (int yy = hs; yy < he; ++yy)
        for (int xx = ws; xx < we; ++xx) m = fmaxf(m, to_f<T>(feat[((size_t)(n * H + yy) * W + xx) * C + c]));
    }
    out[i] = from_f<T>(m);
  }
}

// ---- stage-1 -> stage-2 bridge: what the 16-bit PNG round trip of the reference does to the quasi-dense depth and the
// response map (save: uint32(v * multiplier) -> 16-bit PNG; load: / multiplier, depth <= 0 -> 0), written straight into
// FusionNet's N x 2 x H x W input (channel 0 depth, channel 1 response).  float64 products like numpy's.
__global__ void bridge_kernel(const long long* __restrict__ depth_i64, const float* __restrict__ depth_f32,
                              const float* __restrict__ response, float* __restrict__ out, int64_t hw, int quantize,
                              float mult_depth, float mult_response) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
    float d = depth_i64 ? (float)depth_i64[i] : depth_f32[i];
    float r = response[i];
    if (quantize) {
      // numpy: np.uint32(float32 * python float) computes in float32 (NEP 50 weak scalar), truncates toward zero
      uint32_t qd = (uint32_t)(d * mult_depth) & 0xffffu, qr = (uint32_t)(r * mult_response) & 0xffffu;
      d = (float)qd / mult_depth;
      r = (float)qr / mult_response;
    }
    out[i] = d <= 0.f ? 0.f : d;
    out[hw + i] = r;
  }
}

// out[r][j] = leaky(b[j] + sum_k x[r][k] * w[j][k]): shared-memory tiles (coalesced loads of both operands),
// 32 output features x 64 rows per block, 8 rows per thread.
constexpr int LIN_JT = 32, LIN_RT = 64, LIN_KT = 64;
__global__ void __launch_bounds__(256) linear_leaky_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ b, float* __restrict__ out, int rows,
                                                          int fin, int fout) {
  __shared__ float ws[LIN_JT][LIN_KT + 1];
  __shared__ float xs[LIN_RT][LIN_KT + 1];
  const int j0 = blockIdx.x * LIN_JT, r0 = blockIdx.y * LIN_RT;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[LIN_RT / 8];
#pragma unroll
  for (int i = 0; i < LIN_RT / 8; ++i) acc[i] = 0.f;
  for (int k0 = 0; k0 < fin; k0 += LIN_KT) {
    for (int i = threadIdx.x; i < LIN_JT * LIN_KT; i += 256) {
      const int jj = i / LIN_KT, kk = i - jj * LIN_KT;
      ws[jj][kk] = (j0 + jj < fout && k0 + kk < fin) ? w[(size_t)(j0 + jj) * fin + k0 + kk] : 0.f;
    }
    for (int i = threadIdx.x; i < LIN_RT * LIN_KT; i += 256) {
      const int rr = i / LIN_KT, kk = i - rr * LIN_KT;
      xs[rr][kk] = (r0 + rr < rows && k0 + kk < fin) ? x[(size_t)(r0 + rr) * fin + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < LIN_KT; ++kk) {
      const float wv = ws[tx][kk];
#pragma unroll
      for (int i = 0; i < LIN_RT / 8; ++i) acc[i] = fmaf(xs[ty + 8 * i][kk], wv, acc[i]);
    }
    __syncthreads();
  }
  const int j = j0 + tx;
  if (j < fout) {
    const float bias = b[j];
#pragma unroll
    for (int i = 0; i < LIN_RT / 8; ++i) {
      const int r = r0 + ty + 8 * i;
      if (r < rows) out[(size_t)r * fout + j] = leaky(acc[i] + bias);
    }
  }
}

}  // namespace
}  // namespace rcfd

using namespace rcfd;

extern "C" {

int rcfd_scatter_points_to_depth_map(const double* points_xy, const double* depth, int32_t npts, double* img,
                                     int32_t h, int32_t w, int32_t merge, void* stream) {
  RCFD_CHECK_ARG(img && h > 0 && w > 0 && npts >= 0, "scatter S1: bad args");
  RCFD_CHECK_ARG(npts == 0 || (points_xy && depth), "scatter S1: null points");
  cudaStream_t st = (cudaStream_t)stream;
  if (!merge) {
    cudaError_t e = cudaMemsetAsync(img, 0, sizeof(double) * (size_t)h * w, st);
    if (e != cudaSuccess) { set_error("S1 memset: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
  }
  if (npts == 0) return RCFD_OK;
  if (merge) s1_merge_kernel<<<ceil_div(npts, 128), 128, 0, st>>>(points_xy, depth, npts, img, h, w);
  else s1_plot_kernel<<<ceil_div(npts, 128), 128, 0, st>>>(points_xy, depth, npts, img, h, w);
  RCFD_CHECK_LAUNCH("scatter_s1");
  return RCFD_OK;
}

int rcfd_scatter_tiles_argmax(const float* crops, const float* points, int32_t k, int32_t ph, int32_t pw, int32_t h,
                              int32_t w, int32_t compat, int64_t* depth_i64, float* depth_f32, float* response,
                              void* stream) {
  RCFD_CHECK_ARG(crops && points && response && k > 0 && ph > 0 && pw > 0 && h >= ph && w > 0, "scatter S2: bad args");
  RCFD_CHECK_ARG(compat ? depth_i64 != nullptr : depth_f32 != nullptr, "scatter S2: missing depth output for mode");
  s2_kernel<<<ceil_div((int64_t)h * w, 256), 256, 0, (cudaStream_t)stream>>>(
      crops, points, k, ph, pw, h, w, compat, reinterpret_cast<long long*>(depth_i64), depth_f32, response);
  RCFD_CHECK_LAUNCH("scatter_s2");
  return RCFD_OK;
}

int rcfd_stage1_to_stage2(const int64_t* depth_i64, const float* depth_f32, const float* response, float* input_depth,
                          int32_t h, int32_t w, int32_t quantize_png16, void* stream) {
  RCFD_CHECK_ARG((depth_i64 != nullptr) != (depth_f32 != nullptr), "stage1_to_stage2: exactly one depth input");
  RCFD_CHECK_ARG(response && input_depth && h > 0 && w > 0, "stage1_to_stage2: bad args");
  const int64_t hw = (int64_t)h * w;
  int64_t g = (hw + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  bridge_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(depth_i64), depth_f32, response,
                                                         input_depth, hw, quantize_png16, 256.0f, 16384.0f);
  RCFD_CHECK_LAUNCH("stage1_to_stage2");
  return RCFD_OK;
}

int rcfd_roi_pool_fwd(const void* feat, const float* boxes, void* out, int32_t n, int32_t h, int32_t w, int32_t c,
                      int32_t nbox, int32_t ph, int32_t pw, float spatial_scale, int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(feat && boxes && out && n > 0 && h > 0 && w > 0 && c > 0 && nbox > 0 && ph > 0 && pw > 0,
                 "roi_pool: bad args");
  if (dtype != RCFD_F32 && dtype != RCFD_BF16) { set_error("roi_pool: bad dtype"); return RCFD_EINVAL; }
  const int vw = dtype == RCFD_BF16 ? 8 : 4;
  const bool wide = c % vw == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const int64_t total = (int64_t)nbox * ph * pw * (wide ? c / vw : c);
  int64_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (dtype == RCFD_F32) {
    if (wide) roi_pool_kernel<float><<<(int)g, 256, 0, (cudaStream_t)stream>>>((const float*)feat, boxes, (float*)out, n, h, w, c, nbox, ph, pw, spatial_scale);
    else roi_pool_scalar_kernel<float><<<(int)g, 256, 0, (cudaStream_t)stream>>>((const float*)feat, boxes, (float*)out, n, h, w, c, nbox, ph, pw, spatial_scale);
  } else {
    if (wide) roi_pool_kernel<bf16><<<(int)g, 256, 0, (cudaStream_t)stream>>>((const bf16*)feat, boxes, (bf16*)out, n, h, w, c, nbox, ph, pw, spatial_scale);
    else roi_pool_scalar_kernel<bf16><<<(int)g, 256, 0, (cudaStream_t)stream>>>((const bf16*)feat, boxes, (bf16*)out, n, h, w, c, nbox, ph, pw, spatial_scale);
  }
  RCFD_CHECK_LAUNCH("roi_pool");
  return RCFD_OK;
}

int rcfd_linear_leaky_fwd(const float* x, const float* w, const float* b, float* out, int32_t rows, int32_t in_features,
                          int32_t out_features, void* stream) {
  RCFD_CHECK_ARG(x && w && b && out && rows > 0 && in_features > 0 && out_features > 0, "linear: bad args");
  dim3 grid(ceil_div(out_features, LIN_JT), ceil_div(rows, LIN_RT));
  linear_leaky_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, b, out, rows, in_features, out_features);
  RCFD_CHECK_LAUNCH("linear_leaky");
  return RCFD_OK;
}

}  // extern "C"
