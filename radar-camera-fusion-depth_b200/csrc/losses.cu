// Edge-aware smoothness losses of the reference (src/fusionnet_losses.py:49-125) as fused kernels: value and gradient
// w.r.t. the prediction in one pass over the pixels (the normalisers are fixed by the shape, so nothing waits for a
// reduction), instead of ~20 ATen kernels with full-size temporaries and an autograd graph.  float NCHW at the API,
// like rcfd_masked_l1_loss.
#include "common.cuh"

namespace rcfd {
namespace {

constexpr int LT = 256;

__device__ __forceinline__ float sgnf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }   // d|v|/dv as torch.abs defines it

__device__ __forceinline__ void block_accum2(float a, float b, double* accum) {
  __shared__ double red[2][LT / 32];
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = (double)a; red[1][threadIdx.x >> 5] = (double)b; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int w = 0; w < LT / 32; ++w) t += red[threadIdx.x][w];
    atomicAdd(accum + threadIdx.x, t);
  }
}

// exp(-mean_c |I[c][a] - I[c][b]|) for two pixel offsets of one image
__device__ __forceinline__ float edge_weight(const float* __restrict__ img, int C, size_t plane, size_t a, size_t b) {
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += fabsf(img[c * plane + a] - img[c * plane + b]);
  return expf(-s / (float)C);
}

// smoothness_loss_func: mean(wx |p[y][x] - p[y][x+1]|) + mean(wy |p[y][x] - p[y+1][x]|), forward differences
__global__ void smooth_kernel(const float* __restrict__ pred, const float* __restrict__ image, int N, int C, int H, int W,
                              float inv_cx, float inv_cy, double* __restrict__ accum, float* __restrict__ dpred) {
  const size_t plane = (size_t)H * W;
  const int64_t total = (int64_t)N * H * W;
  float sx = 0.f, sy = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int n = (int)(i / plane);
    const float* p = pred + (size_t)n * plane;
    const float* im = image + (size_t)n * C * plane;
    const size_t o = (size_t)y * W + x;
    const float pc = p[o];
    float g = 0.f;
    if (x < W - 1) {
      const float d = pc - p[o + 1], w = edge_weight(im, C, plane, o, o + 1);
      sx += w * fabsf(d);
      g += w * sgnf(d) * inv_cx;
    }
    if (x > 0) g -= edge_weight(im, C, plane, o - 1, o) * sgnf(p[o - 1] - pc) * inv_cx;
    if (y < H - 1) {
      const float d = pc - p[o + W], w = edge_weight(im, C, plane, o, o + W);
      sy += w * fabsf(d);
      g += w * sgnf(d) * inv_cy;
    }
    if (y > 0) g -= edge_weight(im, C, plane, o - W, o) * sgnf(p[o - W] - pc) * inv_cy;
    if (dpred) dpred[i] = g;
  }
  block_accum2(sx, sy, accum);
}

__global__ void smooth_finish_kernel(const double* __restrict__ accum, double sx_scale, double sy_scale, float* __restrict__ loss) {
  loss[0] = (float)(accum[0] * sx_scale + accum[1] * sy_scale);
}

// generalised Sobel pair of the reference's sobel_filter (src/fusionnet_losses.py:147-161)
__device__ __forceinline__ float sobel_gx(int r, int s, int kh, int kw) {
  const int c = kw / 2;
  const float base = s < c ? 1.f : (s == c ? 0.f : -1.f);
  return (r == kh / 2 && (s == c - 1 || s == c + 1)) ? 2.f * base : base;
}
__device__ __forceinline__ float sobel_gy(int r, int s, int kh, int kw) {
  const int c = kh / 2;
  const float base = r < c ? 1.f : (r == c ? 0.f : -1.f);
  return (s == kw / 2 && (r == c - 1 || r == c + 1)) ? 2.f * base : base;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// forward of sobel_smoothness_loss_func + the per-pixel factors of its gradient (u = weights * w * sgn(p_d) * norm)
__global__ void sobel_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ image, const float* __restrict__ weights,
                                 int N, int H, int W, int kh, int kw, float norm, double* __restrict__ accum,
                                 float* __restrict__ ux, float* __restrict__ uy) {
  const size_t plane = (size_t)H * W;
  const int64_t total = (int64_t)N * H * W;
  float sx = 0.f, sy = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int n = (int)(i / plane);
    const float* p = pred + (size_t)n * plane;
    const float* im = image + (size_t)n * 3 * plane;
    float pdx = 0.f, pdy = 0.f;
    for (int r = 0; r < kh; ++r) {
      const size_t row = (size_t)clampi(y + r - kh / 2, 0, H - 1) * W;
      for (int s = 0; s < kw; ++s) {
        const float v = p[row + clampi(x + s - kw / 2, 0, W - 1)];          // replicate padding
        pdx += sobel_gx(r, s, kh, kw) * v;
        pdy += sobel_gy(r, s, kh, kw) * v;
      }
    }
    float idx = 0.f, idy = 0.f;
    for (int r = 0; r < 3; ++r) {
      const size_t row = (size_t)clampi(y + r - 1, 0, H - 1) * W;
      for (int s = 0; s < 3; ++s) {
        const size_t o = row + clampi(x + s - 1, 0, W - 1);
        const float gray = im[o] * 0.30f + im[plane + o] * 0.59f + im[2 * plane + o] * 0.11f;
        idx += sobel_gx(r, s, 3, 3) * gray;
        idy += sobel_gy(r, s, 3, 3) * gray;
      }
    }
    const float wgt = weights[i];
    const float wx = wgt * expf(-fabsf(idx)), wy = wgt * expf(-fabsf(idy));
    sx += wx * fabsf(pdx);
    sy += wy * fabsf(pdy);
    if (ux) { ux[i] = wx * sgnf(pdx) * norm; uy[i] = wy * sgnf(pdy) * norm; }
  }
  block_accum2(sx, sy, accum);
}

// gradient: transpose of the k x k correlation through the replicate padding (a border pixel also collects what the
// padded positions that replicate it would have received)
__global__ void sobel_bwd_kernel(const float* __restrict__ ux, const float* __restrict__ uy, int N, int H, int W, int kh, int kw,
                                 float* __restrict__ dpred) {
  const size_t plane = (size_t)H * W;
  const int64_t total = (int64_t)N * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int n = (int)(i / plane);
    const float* Ux = ux + (size_t)n * plane;
    const float* Uy = uy + (size_t)n * plane;
    const int py0 = y == 0 ? -(kh / 2) : y, py1 = y == H - 1 ? H - 1 + kh / 2 : y;     // padded rows that read pixel row y
    const int px0 = x == 0 ? -(kw / 2) : x, px1 = x == W - 1 ? W - 1 + kw / 2 : x;
    float g = 0.f;
    for (int py = py0; py <= py1; ++py)
      for (int r = 0; r < kh; ++r) {
        const int oy = py - (r - kh / 2);            // output position whose tap r lands on padded row py
        if (oy < 0 || oy >= H) continue;
        for (int px = px0; px <= px1; ++px)
          for (int s = 0; s < kw; ++s) {
            const int ox = px - (s - kw / 2);
            if (ox < 0 || ox >= W) continue;
            const size_t o = (size_t)oy * W + ox;
            g += sobel_gx(r, s, kh, kw) * Ux[o] + sobel_gy(r, s, kh, kw) * Uy[o];
          }
      }
    dpred[i] = g;
  }
}

inline int grid_of(int64_t work) {
  int64_t b = (work + LT - 1) / LT;
  return (int)(b > 148 * 8 ? 148 * 8 : (b < 1 ? 1 : b));
}

}  // namespace
}  // namespace rcfd

using namespace rcfd;

extern "C" {

int rcfd_smoothness_loss(const float* predict, const float* image, int32_t n, int32_t c, int32_t h, int32_t w, double* accum,
                         float* loss, float* dpredict, void* stream) {
  RCFD_CHECK_ARG(predict && image && accum && loss && n > 0 && c > 0 && h > 1 && w > 1, "smoothness_loss: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(accum, 0, 2 * sizeof(double), st);
  if (e != cudaSuccess) { set_error("smoothness_loss memset: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
  const double cx = (double)n * h * (w - 1), cy = (double)n * (h - 1) * w;
  smooth_kernel<<<grid_of((int64_t)n * h * w), LT, 0, st>>>(predict, image, n, c, h, w, (float)(1.0 / cx), (float)(1.0 / cy), accum,
                                                          dpredict);
  RCFD_CHECK_LAUNCH("smoothness_loss");
  smooth_finish_kernel<<<1, 1, 0, st>>>(accum, 1.0 / cx, 1.0 / cy, loss);
  RCFD_CHECK_LAUNCH("smoothness_loss finish");
  return RCFD_OK;
}

int rcfd_sobel_smoothness_loss(const float* predict, const float* image, const float* weights, int32_t n, int32_t h, int32_t w,
                               int32_t kh, int32_t kw, double* accum, float* scratch, float* loss, float* dpredict, void* stream) {
  RCFD_CHECK_ARG(predict && image && weights && accum && loss && n > 0 && h > 0 && w > 0 && kh >= 3 && kw >= 3 && (kh & 1) &&
                     (kw & 1) && kh <= 15 && kw <= 15 && (dpredict == nullptr || scratch != nullptr),
                 "sobel_smoothness_loss: bad args (odd filter sizes 3..15; scratch = 2 * n * h * w floats when the gradient is wanted)");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(accum, 0, 2 * sizeof(double), st);
  if (e != cudaSuccess) { set_error("sobel_smoothness_loss memset: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
  const int64_t total = (int64_t)n * h * w;
  const double norm = 1.0 / ((double)total * kh * kw);
  float* ux = dpredict ? scratch : nullptr;
  float* uy = dpredict ? scratch + total : nullptr;
  sobel_fwd_kernel<<<grid_of(total), LT, 0, st>>>(predict, image, weights, n, h, w, kh, kw, (float)norm, accum, ux, uy);
  RCFD_CHECK_LAUNCH("sobel_smoothness_loss");
  smooth_finish_kernel<<<1, 1, 0, st>>>(accum, norm, norm, loss);
  RCFD_CHECK_LAUNCH("sobel_smoothness_loss finish");
  if (dpredict) {
    sobel_bwd_kernel<<<grid_of(total), LT, 0, st>>>(ux, uy, n, h, w, kh, kw, dpredict);
    RCFD_CHECK_LAUNCH("sobel_smoothness_loss grad");
  }
  return RCFD_OK;
}

}  // extern "C"
