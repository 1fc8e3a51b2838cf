// Backward / loss kernels of the RadarNet stage-1 training step (reference src/radarnet_main.py:320-403,
// src/radarnet_model.py:126-167): gradient of torchvision.ops.roi_pool, gradient of the point MLP
// (Linear + LeakyReLU stack), and the validity-weighted binary cross entropy with logits.
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

// ------------------------------------------------------------------------------------------------ roi_pool backward
// torchvision roi_pool_backward: the gradient of every output bin goes to its arg-max input element (first maximum in
// row-major scan order, strict '>'); bins of different boxes (and neighbouring bins of one box: bin width ~1.007 px)
// overlap, hence atomics into a float accumulator.
template <typename T>
__global__ void roi_pool_bwd_kernel(const T* __restrict__ feat, const float* __restrict__ boxes, const T* __restrict__ dout,
                                    float* __restrict__ dfeat, int N, int H, int W, int C, int nbox, int PH, int PW, float scale) {
  const int64_t total = (int64_t)nbox * PH * PW * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int pw = (int)(r % PW); r /= PW;
    const int ph = (int)(r % PH);
    const int b = (int)(r / PH);
    const float* box = boxes + (size_t)b * 5;
    const int n = (int)box[0];
    if (n < 0 || n >= N) continue;
    const int sw = (int)roundf(box[1] * scale), sh = (int)roundf(box[2] * scale);
    const int ew = (int)roundf(box[3] * scale), eh = (int)roundf(box[4] * scale);
    const int rw = max(ew - sw + 1, 1), rh = max(eh - sh + 1, 1);
    const float bh = (float)rh / (float)PH, bw = (float)rw / (float)PW;
    int hs = (int)floorf((float)ph * bh), he = (int)ceilf((float)(ph + 1) * bh);
    int ws = (int)floorf((float)pw * bw), we = (int)ceilf((float)(pw + 1) * bw);
    hs = min(max(hs + sh, 0), H); he = min(max(he + sh, 0), H);
    ws = min(max(ws + sw, 0), W); we = min(max(we + sw, 0), W);
    float best = -3.402823466e+38f;
    int64_t arg = -1;
    for (int yy = hs; yy < he; ++yy)
      for (int xx = ws; xx < we; ++xx) {
        const int64_t idx = ((int64_t)(n * H + yy) * W + xx) * C + c;
        const float v = to_f<T>(feat[idx]);
        if (v > best) { best = v; arg = idx; }
      }
    if (arg >= 0) atomicAdd(dfeat + arg, to_f<T>(dout[i]));
  }
}

template <typename T>
__global__ void cast_f32_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = from_f<T>(src[i]);
}

// ------------------------------------------------------------------------------------------------ Linear + LeakyReLU backward
// dpre = dy * leaky'(y) (y is the post-activation output: y > 0 <=> pre > 0), db = column sums of dpre
__global__ void linear_dpre_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dpre,
                                   float* __restrict__ db, int rows, int fout) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= fout) return;
  float s = 0.f;
  for (int k = 0; k < rows; ++k) {
    const size_t o = (size_t)k * fout + j;
    const float d = y[o] > 0.f ? dy[o] : kLeakySlope * dy[o];
    dpre[o] = d;
    s += d;
  }
  db[j] = s;
}

// C[m][n] = sum_k A(m, k) * B(k, n) on 32 x 32 tiles.  TA: A is stored [k][m] (transposed), else [m][k]; B stored [k][n].
template <bool TA>
__global__ void __launch_bounds__(256) small_gemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ Cm,
                                                         int M, int Nn, int K, int lda, int ldb, int ldc) {
  __shared__ float as[32][33], bs[32][33];
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // ty 0..7: rows ty, ty+8, ty+16, ty+24
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i;
      // as[m][k]
      if (TA) {
        const int k = k0 + r, m = m0 + tx;
        as[tx][r] = (k < K && m < M) ? A[(size_t)k * lda + m] : 0.f;
      } else {
        const int m = m0 + r, k = k0 + tx;
        as[r][tx] = (k < K && m < M) ? A[(size_t)m * lda + k] : 0.f;
      }
      const int kb = k0 + r, n = n0 + tx;
      bs[r][tx] = (kb < K && n < Nn) ? B[(size_t)kb * ldb + n] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const float bv = bs[kk][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(as[ty + 8 * i][kk], bv, acc[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + 8 * i, n = n0 + tx;
    if (m < M && n < Nn) Cm[(size_t)m * ldc + n] = acc[i];
  }
}

// ------------------------------------------------------------------------------------------------ weighted BCE with logits
// torch.nn.functional.binary_cross_entropy_with_logits(x, t, pos_weight = pw, reduction = 'none'):
//   l = (1 - t) x + (1 + (pw - 1) t) (log1p(exp(-|x|)) + max(-x, 0));     L = sum(v l) / sum(v)
__global__ void bce_accum_kernel(const float* __restrict__ x, const float* __restrict__ t, const float* __restrict__ v, float pw,
                                 double* __restrict__ accum, int64_t n) {
  float sl = 0.f, sv = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xi = x[i], ti = t[i], vi = v[i];
    const float lw = 1.f + (pw - 1.f) * ti;
    const float l = (1.f - ti) * xi + lw * (log1pf(expf(-fabsf(xi))) + fmaxf(-xi, 0.f));
    sl += vi * l;
    sv += vi;
  }
  __shared__ double red[2][NT / 32];
  float a[2] = {sl, sv};
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    float q = a[j];
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if ((threadIdx.x & 31) == 0) red[j][threadIdx.x >> 5] = (double)q;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0.0;
    for (int w = 0; w < NT / 32; ++w) s += red[threadIdx.x][w];
    atomicAdd(accum + threadIdx.x, s);
  }
}
__global__ void bce_finish_kernel(const float* __restrict__ x, const float* __restrict__ t, const float* __restrict__ v, float pw,
                                  const double* __restrict__ accum, float* __restrict__ loss, float* __restrict__ dx, int64_t n) {
  if (blockIdx.x == 0 && threadIdx.x == 0) loss[0] = (float)(accum[0] / accum[1]);
  if (!dx) return;
  const float inv = (float)(1.0 / accum[1]);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xi = x[i], ti = t[i];
    const float lw = 1.f + (pw - 1.f) * ti;
    const float s = sigmoid_precise(xi);
    dx[i] = v[i] * inv * ((1.f - ti) - lw * (1.f - s));
  }
}

}  // namespace
}  // namespace rcfd

using namespace rcfd;

extern "C" {

int rcfd_roi_pool_bwd(const void* feat, const float* boxes, const void* dout, float* dfeat_f32, int32_t n, int32_t h, int32_t w,
                      int32_t c, int32_t nbox, int32_t ph, int32_t pw, float spatial_scale, int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(feat && boxes && dout && dfeat_f32 && n > 0 && h > 0 && w > 0 && c > 0 && nbox > 0 && ph > 0 && pw > 0,
                 "roi_pool_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(dfeat_f32, 0, sizeof(float) * (size_t)n * h * w * c, st);
  if (e != cudaSuccess) { set_error("roi_pool_bwd memset: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
  const int64_t total = (int64_t)nbox * ph * pw * c;
  if (dtype == RCFD_F32)
    roi_pool_bwd_kernel<float><<<blocks_for(total), NT, 0, st>>>((const float*)feat, boxes, (const float*)dout, dfeat_f32, n, h, w, c,
                                                                 nbox, ph, pw, spatial_scale);
  else if (dtype == RCFD_BF16)
    roi_pool_bwd_kernel<bf16><<<blocks_for(total), NT, 0, st>>>((const bf16*)feat, boxes, (const bf16*)dout, dfeat_f32, n, h, w, c,
                                                                nbox, ph, pw, spatial_scale);
  else { set_error("roi_pool_bwd: bad dtype"); return RCFD_EINVAL; }
  RCFD_CHECK_LAUNCH("roi_pool_bwd");
  return RCFD_OK;
}

int rcfd_cast_f32(const float* src, void* dst, int64_t count, int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(src && dst && count > 0, "cast_f32: bad args");
  if (dtype == RCFD_F32) cast_f32_kernel<float><<<blocks_for(count), NT, 0, (cudaStream_t)stream>>>(src, (float*)dst, count);
  else if (dtype == RCFD_BF16) cast_f32_kernel<bf16><<<blocks_for(count), NT, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, count);
  else { set_error("cast_f32: bad dtype"); return RCFD_EINVAL; }
  RCFD_CHECK_LAUNCH("cast_f32");
  return RCFD_OK;
}

int rcfd_linear_leaky_bwd(const float* x, const float* w, const float* y, const float* dy, float* dpre_scratch, float* dx,
                          float* dw, float* db, int32_t rows, int32_t in_features, int32_t out_features, void* stream) {
  RCFD_CHECK_ARG(x && w && y && dy && dpre_scratch && dw && db && rows > 0 && in_features > 0 && out_features > 0,
                 "linear_leaky_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  linear_dpre_kernel<<<ceil_div(out_features, 128), 128, 0, st>>>(y, dy, dpre_scratch, db, rows, out_features);
  RCFD_CHECK_LAUNCH("linear_dpre");
  // dW[j][i] = sum_k dpre[k][j] x[k][i]:  A = dpre stored [k][j] (transposed access), B = x [k][i]
  dim3 gw(ceil_div(in_features, 32), ceil_div(out_features, 32));
  small_gemm_kernel<true><<<gw, 256, 0, st>>>(dpre_scratch, x, dw, out_features, in_features, rows, out_features, in_features,
                                              in_features);
  RCFD_CHECK_LAUNCH("linear_dw");
  if (dx) {
    // dx[k][i] = sum_j dpre[k][j] w[j][i]
    dim3 gx(ceil_div(in_features, 32), ceil_div(rows, 32));
    small_gemm_kernel<false><<<gx, 256, 0, st>>>(dpre_scratch, w, dx, rows, in_features, out_features, out_features, in_features,
                                                 in_features);
    RCFD_CHECK_LAUNCH("linear_dx");
  }
  return RCFD_OK;
}

int rcfd_bce_logits_loss(const float* logits, const float* target, const float* validity, float pos_weight, double* accum,
                         float* loss, float* dlogits, int64_t count, void* stream) {
  RCFD_CHECK_ARG(logits && target && validity && accum && loss && count > 0, "bce_logits_loss: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(accum, 0, 2 * sizeof(double), st);
  if (e != cudaSuccess) { set_error("bce memset: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
  bce_accum_kernel<<<blocks_for(count, 148 * 8), NT, 0, st>>>(logits, target, validity, pos_weight, accum, count);
  RCFD_CHECK_LAUNCH("bce accum");
  bce_finish_kernel<<<dlogits ? blocks_for(count, 148 * 8) : 1, NT, 0, st>>>(logits, target, validity, pos_weight, accum, loss,
                                                                             dlogits, count);
  RCFD_CHECK_LAUNCH("bce finish");
  return RCFD_OK;
}

}  // extern "C"
