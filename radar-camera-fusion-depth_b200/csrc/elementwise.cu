// HBM-bound NHWC passes around the convolutions: BatchNorm finalize / apply / backward,
// gated fusion, max-pool, nearest-upsample backward, layout conversion, depth head,
// loss, outlier removal, Adam, weight packing.  All vectorised 4 channels per thread
// where the channel count allows it (coalesced 16 B fp32 / 8 B bf16 accesses).
#include "common.cuh"

namespace rcfd {
namespace {

constexpr int NT = 256;

}  // namespace (anonymous)
// rcfd_set_option("bn_vectors_per_thread"): 16-byte vectors one thread of the wide BatchNorm kernels handles at least
// (their per-thread parameter prologue is paid once; 1 = one vector per thread until the grid cap, the round-1 sizing)
int g_bn_vectors_per_thread = 8;
int g_bn_reduce_rows_per_thread = 16;  // bn_bwd_reduce: rows one thread sums at least (rcfd_set_option)
int g_bn_fwd_vectors_per_thread = 4;   // bn_act_fwd / bn_train_act_fwd
int g_ew_vectors_per_thread = 4;       // add_inplace / leaky_bwd (no prologue)
namespace {
inline int grid_for(int64_t work, int per_block = NT, int cap = 148 * 16) {
  int64_t g = (work + per_block - 1) / per_block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------ BatchNorm
__global__ void bn_finalize_kernel(const double* sum, const double* sqsum, const float* gamma,
                                   const float* beta, float* rmean, float* rvar, float* scale,
                                   float* shift, float* smean, float* sinv, int C, double count,
                                   float eps, float momentum) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mean = sum[c] / count;
  double var = sqsum[c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  double invstd = 1.0 / sqrt(var + (double)eps);
  float sc = (float)((double)gamma[c] * invstd);
  scale[c] = sc;
  shift[c] = (float)((double)beta[c] - mean * (double)gamma[c] * invstd);
  if (smean) smean[c] = (float)mean;
  if (sinv) sinv[c] = (float)invstd;
  if (rmean) {
    double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    rmean[c] = (float)((1.0 - momentum) * (double)rmean[c] + momentum * mean);
    rvar[c] = (float)((1.0 - momentum) * (double)rvar[c] + momentum * unbiased);
  }
}

__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* rmean,
                               const float* rvar, float* scale, float* shift, int C, float eps) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float inv = 1.0f / sqrtf(rvar[c] + eps);
  float sc = gamma[c] * inv;
  scale[c] = sc;
  shift[c] = beta[c] - rmean[c] * sc;
}

template <typename T>
__global__ void bn_act_fwd_kernel(const T* __restrict__ y, const float* __restrict__ scale,
                                  const float* __restrict__ shift, const T* __restrict__ res,
                                  T* __restrict__ out, int64_t nvec, int C, int act) {
  const int CV = C >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV) * 4;
    float4 v = Vec4<T>::ld(y + i * 4);
    float4 s = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (scale) {
      s = *reinterpret_cast<const float4*>(scale + c);
      b = *reinterpret_cast<const float4*>(shift + c);
    }
    v.x = apply_act(fmaf(v.x, s.x, b.x), act, 0.f, 0.f);
    v.y = apply_act(fmaf(v.y, s.y, b.y), act, 0.f, 0.f);
    v.z = apply_act(fmaf(v.z, s.z, b.z), act, 0.f, 0.f);
    v.w = apply_act(fmaf(v.w, s.w, b.w), act, 0.f, 0.f);
    if (res) {
      float4 r = Vec4<T>::ld(res + i * 4);
      v.x = leaky(v.x + r.x); v.y = leaky(v.y + r.y); v.z = leaky(v.z + r.z); v.w = leaky(v.w + r.w);
    }
    Vec4<T>::st(out + i * 4, v);
  }
}

__device__ __forceinline__ float dact(float pre, float dz, int act) {
  if (act == RCFD_ACT_LEAKY) return pre > 0.f ? dz : kLeakySlope * dz;
  if (act == RCFD_ACT_SIGMOID) {
    float s = sigmoid_precise(pre);
    return dz * s * (1.f - s);
  }
  return dz;
}

// per-channel sums of dpre and dpre*xhat.  Thread t owns channel-vector (t % CVP); rows strided.
template <typename T>
__global__ void bn_bwd_reduce_kernel(const T* __restrict__ dz, const T* __restrict__ y,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                     double* __restrict__ sums, int64_t pixels, int C, int CVP, int act,
                                     int64_t rows_per_block) {
  __shared__ float red[2][NT * 4];
  const int CV = C >> 2;
  const int cv = threadIdx.x % CVP;
  const int r0 = threadIdx.x / CVP;
  const int rstep = NT / CVP;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  const int64_t pbeg = blockIdx.x * rows_per_block;
  int64_t pend = pbeg + rows_per_block;
  if (pend > pixels) pend = pixels;
  if (cv < CV) {
    const int c = cv * 4;
    const float4 sc = *reinterpret_cast<const float4*>(scale + c);
    const float4 sh = *reinterpret_cast<const float4*>(shift + c);
    const float4 mu = *reinterpret_cast<const float4*>(mean + c);
    const float4 is = *reinterpret_cast<const float4*>(invstd + c);
    for (int64_t p = pbeg + r0; p < pend; p += rstep) {
      const float4 g = Vec4<T>::ld(dz + p * C + c);
      const float4 v = Vec4<T>::ld(y + p * C + c);
      float d;
      d = dact(fmaf(v.x, sc.x, sh.x), g.x, act); s[0] += d; q[0] += d * (v.x - mu.x) * is.x;
      d = dact(fmaf(v.y, sc.y, sh.y), g.y, act); s[1] += d; q[1] += d * (v.y - mu.y) * is.y;
      d = dact(fmaf(v.z, sc.z, sh.z), g.z, act); s[2] += d; q[2] += d * (v.z - mu.z) * is.z;
      d = dact(fmaf(v.w, sc.w, sh.w), g.w, act); s[3] += d; q[3] += d * (v.w - mu.w) * is.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[0][threadIdx.x * 4 + j] = s[j];
    red[1][threadIdx.x * 4 + j] = q[j];
  }
  __syncthreads();
  // thread t < C reduces channel t over the rstep row-groups
  for (int c = threadIdx.x; c < C; c += NT) {
    const int v = c >> 2, j = c & 3;
    double a = 0.0, b = 0.0;
    for (int r = 0; r < rstep; ++r) {
      a += (double)red[0][(r * CVP + v) * 4 + j];
      b += (double)red[1][(r * CVP + v) * 4 + j];
    }
    atomicAdd(sums + c, a);
    atomicAdd(sums + C + c, b);
  }
}

template <typename T>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ dz, const T* __restrict__ y,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const double* __restrict__ sums, T* __restrict__ dy, int64_t nvec,
                                    int C, int act, float inv_count, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta) {
  const int CV = C >> 2;
  if (blockIdx.x == 0) {                       // parameter gradients ride along (one block, C values)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dbeta) dbeta[c] = (float)sums[c];
      if (dgamma) dgamma[c] = (float)sums[C + c];
    }
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV) * 4;
    const float4 g = Vec4<T>::ld(dz + i * 4);
    const float4 v = Vec4<T>::ld(y + i * 4);
    const float4 sc = *reinterpret_cast<const float4*>(scale + c);
    const float4 sh = *reinterpret_cast<const float4*>(shift + c);
    const float4 mu = *reinterpret_cast<const float4*>(mean + c);
    const float4 is = *reinterpret_cast<const float4*>(invstd + c);
    float4 o;
    float d;
    d = dact(fmaf(v.x, sc.x, sh.x), g.x, act);
    o.x = sc.x * (d - (float)sums[c + 0] * inv_count - (v.x - mu.x) * is.x * (float)sums[C + c + 0] * inv_count);
    d = dact(fmaf(v.y, sc.y, sh.y), g.y, act);
    o.y = sc.y * (d - (float)sums[c + 1] * inv_count - (v.y - mu.y) * is.y * (float)sums[C + c + 1] * inv_count);
    d = dact(fmaf(v.z, sc.z, sh.z), g.z, act);
    o.z = sc.z * (d - (float)sums[c + 2] * inv_count - (v.z - mu.z) * is.z * (float)sums[C + c + 2] * inv_count);
    d = dact(fmaf(v.w, sc.w, sh.w), g.w, act);
    o.w = sc.w * (d - (float)sums[c + 3] * inv_count - (v.w - mu.w) * is.w * (float)sums[C + c + 3] * inv_count);
    Vec4<T>::st(dy + i * 4, o);
  }
}

// ---- 16-byte-vector variants (C % V16<T>::N == 0): 8 bf16 / 4 float per thread and access
// The launch uses 256-thread blocks and C / N divides 256, so a thread meets the SAME channel group in every
// grid-stride iteration: per-channel parameters are loaded once, the loop body is 2 loads + 1 store.
template <typename T>
__global__ void bn_act_fwd_wide(const T* __restrict__ y, const float* __restrict__ scale,
                                const float* __restrict__ shift, const T* __restrict__ res, T* __restrict__ out,
                                int64_t nvec, int C, int act) {
  constexpr int N = V16<T>::N;
  const int CV = C / N;
  const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int c = (int)(i0 % CV) * N;
  float sc[N], sh[N];
#pragma unroll
  for (int k = 0; k < N; ++k) { sc[k] = scale ? scale[c + k] : 1.f; sh[k] = scale ? shift[c + k] : 0.f; }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = i0;
  for (; i + 3 * stride < nvec; i += 4 * stride) {          // grid sized for >= 4 vectors per thread, all loads in flight
    float v[4][N], r[4][N];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      V16<T>::ld(y + (i + u * stride) * N, v[u]);
      if (res) V16<T>::ld(res + (i + u * stride) * N, r[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        float x = apply_act(fmaf(v[u][k], sc[k], sh[k]), act, 0.f, 0.f);
        if (res) x = leaky(x + r[u][k]);
        v[u][k] = x;
      }
      V16<T>::st(out + (i + u * stride) * N, v[u]);
    }
  }
  for (; i < nvec; i += stride) {
    float v[N], r[N];
    V16<T>::ld(y + i * N, v);
    if (res) V16<T>::ld(res + i * N, r);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      float x = apply_act(fmaf(v[k], sc[k], sh[k]), act, 0.f, 0.f);
      if (res) x = leaky(x + r[k]);
      v[k] = x;
    }
    V16<T>::st(out + i * N, v);
  }
}

constexpr int BN_TRAIN_MAX_C = 1024;
// Training-mode BatchNorm forward in ONE pass over the tensor: every thread derives scale / shift of its own
// channel group from the batch sums (same fp64 formulas as bn_finalize_kernel, so the values are identical),
// block 0 also publishes scale / shift / mean / invstd for the backward pass and updates the running statistics.
template <typename T>
__global__ void bn_train_act_fwd_wide(const T* __restrict__ y, const double* __restrict__ sum, const double* __restrict__ sqsum,
                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                      float* __restrict__ rmean, float* __restrict__ rvar, float* __restrict__ scale,
                                      float* __restrict__ shift, float* __restrict__ smean, float* __restrict__ sinv,
                                      const T* __restrict__ res, T* __restrict__ out, int64_t nvec, int C, int act,
                                      double count, float eps, float momentum) {
  constexpr int N = V16<T>::N;
  const int CV = C / N;
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const double mean = sum[c] / count;
      double var = sqsum[c] / count - mean * mean;
      if (var < 0.0) var = 0.0;
      const double invstd = 1.0 / sqrt(var + (double)eps);
      scale[c] = (float)((double)gamma[c] * invstd);
      shift[c] = (float)((double)beta[c] - mean * (double)gamma[c] * invstd);
      smean[c] = (float)mean;
      sinv[c] = (float)invstd;
      if (rmean) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        rmean[c] = (float)((1.0 - momentum) * (double)rmean[c] + momentum * mean);
        rvar[c] = (float)((1.0 - momentum) * (double)rvar[c] + momentum * unbiased);
      }
    }
  }
  // scale / shift of every channel once per block (fp64 divide + sqrt: not per thread), then registers
  __shared__ float s_sc[BN_TRAIN_MAX_C], s_sh[BN_TRAIN_MAX_C];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double mean = sum[c] / count;
    double var = sqsum[c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double invstd = 1.0 / sqrt(var + (double)eps);
    s_sc[c] = (float)((double)gamma[c] * invstd);
    s_sh[c] = (float)((double)beta[c] - mean * (double)gamma[c] * invstd);
  }
  __syncthreads();
  const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i0 >= nvec) return;
  const int c0 = (int)(i0 % CV) * N;
  float sc[N], sh[N];
#pragma unroll
  for (int k = 0; k < N; ++k) { sc[k] = s_sc[c0 + k]; sh[k] = s_sh[c0 + k]; }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = i0;
  for (; i + 3 * stride < nvec; i += 4 * stride) {          // grid sized for >= 4 vectors per thread, all loads in flight
    float v[4][N], r[4][N];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      V16<T>::ld(y + (i + u * stride) * N, v[u]);
      if (res) V16<T>::ld(res + (i + u * stride) * N, r[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        float x = apply_act(fmaf(v[u][k], sc[k], sh[k]), act, 0.f, 0.f);
        if (res) x = leaky(x + r[u][k]);
        v[u][k] = x;
      }
      V16<T>::st(out + (i + u * stride) * N, v[u]);
    }
  }
  for (; i < nvec; i += stride) {
    float v[N], r[N];
    V16<T>::ld(y + i * N, v);
    if (res) V16<T>::ld(res + i * N, r);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      float x = apply_act(fmaf(v[k], sc[k], sh[k]), act, 0.f, 0.f);
      if (res) x = leaky(x + r[k]);
      v[k] = x;
    }
    V16<T>::st(out + i * N, v);
  }
}

template <typename T>
__global__ void bn_bwd_reduce_wide(const T* __restrict__ dz, const T* __restrict__ y, const float* __restrict__ scale,
                                   const float* __restrict__ shift, const float* __restrict__ mean,
                                   const float* __restrict__ invstd, double* __restrict__ sums, int64_t pixels, int C,
                                   int CVP, int act, int64_t rows_per_block) {
  constexpr int N = V16<T>::N;
  __shared__ float red[2][NT * N];
  const int CV = C / N;
  const int cv = threadIdx.x % CVP;
  const int r0 = threadIdx.x / CVP;
  const int rstep = NT / CVP;
  float s[N], q[N];
#pragma unroll
  for (int k = 0; k < N; ++k) { s[k] = 0.f; q[k] = 0.f; }
  const int64_t pbeg = blockIdx.x * rows_per_block;
  int64_t pend = pbeg + rows_per_block;
  if (pend > pixels) pend = pixels;
  if (cv < CV) {
    const int c = cv * N;
    float sc[N], sh[N], mu[N], is[N];
#pragma unroll
    for (int k = 0; k < N; ++k) { sc[k] = scale[c + k]; sh[k] = shift[c + k]; mu[k] = mean[c + k]; is[k] = invstd[c + k]; }
    int64_t p = pbeg + r0;
    for (; p + 3 * (int64_t)rstep < pend; p += 4 * (int64_t)rstep) {      // 8 independent 16-byte loads in flight per thread
      float g[4][N], v[4][N];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        V16<T>::ld(dz + (p + u * (int64_t)rstep) * C + c, g[u]);
        V16<T>::ld(y + (p + u * (int64_t)rstep) * C + c, v[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const float d = dact(fmaf(v[u][k], sc[k], sh[k]), g[u][k], act);
          s[k] += d;
          q[k] += d * (v[u][k] - mu[k]) * is[k];
        }
    }
    for (; p < pend; p += rstep) {
      float g[N], v[N];
      V16<T>::ld(dz + p * C + c, g);
      V16<T>::ld(y + p * C + c, v);
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const float d = dact(fmaf(v[k], sc[k], sh[k]), g[k], act);
        s[k] += d;
        q[k] += d * (v[k] - mu[k]) * is[k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    red[0][threadIdx.x * N + k] = s[k];
    red[1][threadIdx.x * N + k] = q[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += NT) {
    const int v = c / N, j = c % N;
    double a = 0.0, b = 0.0;
    for (int r = 0; r < rstep; ++r) {
      a += (double)red[0][(r * CVP + v) * N + j];
      b += (double)red[1][(r * CVP + v) * N + j];
    }
    atomicAdd(sums + c, a);
    atomicAdd(sums + C + c, b);
  }
}

template <typename T>
__global__ void bn_bwd_apply_wide(const T* __restrict__ dz, const T* __restrict__ y, const float* __restrict__ scale,
                                  const float* __restrict__ shift, const float* __restrict__ mean,
                                  const float* __restrict__ invstd, const double* __restrict__ sums, T* __restrict__ dy,
                                  int64_t nvec, int C, int act, float inv_count, float* __restrict__ dgamma,
                                  float* __restrict__ dbeta) {
  constexpr int N = V16<T>::N;
  const int CV = C / N;
  if (blockIdx.x == 0) {                       // parameter gradients ride along (one block, C values)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dbeta) dbeta[c] = (float)sums[c];
      if (dgamma) dgamma[c] = (float)sums[C + c];
    }
  }
  const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int c = (int)(i0 % CV) * N;
  float sc[N], sh[N], mu[N], k1[N], k2[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    sc[k] = scale[c + k]; sh[k] = shift[c + k]; mu[k] = mean[c + k];
    k1[k] = (float)sums[c + k] * inv_count;                     // mean of dpre
    k2[k] = invstd[c + k] * (float)sums[C + c + k] * inv_count;       // invstd * mean(dpre * xhat)
  }
  // the grid is sized for >= 4 vectors per thread (the 56 parameter loads above are per thread, not per vector: with one
  // vector per thread they made a 22x44 x 256 map take 19.6 us on 968 blocks against 5 us of traffic); 8 loads in flight
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = i0;
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    float g[4][N], v[4][N];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      V16<T>::ld(dz + (i + u * stride) * N, g[u]);
      V16<T>::ld(y + (i + u * stride) * N, v[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float o[N];
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const float d = dact(fmaf(v[u][k], sc[k], sh[k]), g[u][k], act);
        o[k] = sc[k] * (d - k1[k] - (v[u][k] - mu[k]) * k2[k]);
      }
      V16<T>::st(dy + (i + u * stride) * N, o);
    }
  }
  for (; i < nvec; i += stride) {
    float g[N], v[N], o[N];
    V16<T>::ld(dz + i * N, g);
    V16<T>::ld(y + i * N, v);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float d = dact(fmaf(v[k], sc[k], sh[k]), g[k], act);
      o[k] = sc[k] * (d - k1[k] - (v[k] - mu[k]) * k2[k]);
    }
    V16<T>::st(dy + i * N, o);
  }
}

// BatchNorm backward of a SMALL map in one launch: a CTA owns one 16-byte channel vector over ALL pixels, so the
// per-channel sums never leave the CTA (no second kernel, no grid barrier): pass 1 sums, block reduce, pass 2 re-reads the
// two tensors (L1 / L2 hits at these sizes) and writes dy.  POST: the gradient first goes through the LeakyReLU after a
// residual add (z = its output); the masked gradient is written to dzm for the shortcut branch.
constexpr int NT_SLICED = 512;
template <typename T, bool POST>
__global__ void __launch_bounds__(NT_SLICED)
bn_bwd_sliced_kernel(const T* __restrict__ dz, const T* __restrict__ y, const T* __restrict__ zpost, T* dzm,
                     const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                     const float* __restrict__ invstd, T* __restrict__ dy, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, int pixels, int C, int act, float inv_count) {
  constexpr int N = V16<T>::N;
  constexpr int U = 4;
  __shared__ double red[NT_SLICED / 32][2 * N];
  __shared__ float tot[2 * N];
  const int c = blockIdx.x * N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float sc[N], sh[N], mu[N], is[N], s[N], q[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    sc[k] = scale[c + k]; sh[k] = shift[c + k]; mu[k] = mean[c + k]; is[k] = invstd[c + k];
    s[k] = 0.f; q[k] = 0.f;
  }
  auto pass1 = [&](const float* g, const float* v) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float d = dact(fmaf(v[k], sc[k], sh[k]), g[k], act);
      s[k] += d;
      q[k] += d * (v[k] - mu[k]) * is[k];
    }
  };
  int p = tid;
  for (; p + (U - 1) * NT_SLICED < pixels; p += U * NT_SLICED) {
    float g[U][N], v[U][N];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t o = (size_t)(p + u * NT_SLICED) * C + c;
      V16<T>::ld(dz + o, g[u]);
      V16<T>::ld(y + o, v[u]);
      if (POST) {
        float z[N];
        V16<T>::ld(zpost + o, z);
#pragma unroll
        for (int k = 0; k < N; ++k) g[u][k] = to_f<T>(from_f<T>(z[k] > 0.f ? g[u][k] : kLeakySlope * g[u][k]));   // as stored
        V16<T>::st(dzm + o, g[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) pass1(g[u], v[u]);
  }
  for (; p < pixels; p += NT_SLICED) {
    float g[N], v[N];
    const size_t o = (size_t)p * C + c;
    V16<T>::ld(dz + o, g);
    V16<T>::ld(y + o, v);
    if (POST) {
      float z[N];
      V16<T>::ld(zpost + o, z);
#pragma unroll
      for (int k = 0; k < N; ++k) g[k] = to_f<T>(from_f<T>(z[k] > 0.f ? g[k] : kLeakySlope * g[k]));
      V16<T>::st(dzm + o, g);
    }
    pass1(g, v);
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s[k] += __shfl_xor_sync(0xffffffffu, s[k], off);
      q[k] += __shfl_xor_sync(0xffffffffu, q[k], off);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) { red[warp][k] = (double)s[k]; red[warp][N + k] = (double)q[k]; }
  }
  __syncthreads();
  if (tid < 2 * N) {
    double a = 0.0;
    for (int w = 0; w < NT_SLICED / 32; ++w) a += red[w][tid];
    tot[tid] = (float)a;
    if (tid < N) { if (dbeta) dbeta[c + tid] = (float)a; }
    else if (dgamma) dgamma[c + tid - N] = (float)a;
  }
  __syncthreads();
  float k1[N], k2[N];
#pragma unroll
  for (int k = 0; k < N; ++k) { k1[k] = tot[k] * inv_count; k2[k] = is[k] * tot[N + k] * inv_count; }
  // pass 2 (POST: dzm holds this thread's own masked gradients, read back in program order)
  const T* gsrc = POST ? dzm : dz;
  auto pass2 = [&](const float* g, const float* v, float* o) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float d = dact(fmaf(v[k], sc[k], sh[k]), g[k], act);
      o[k] = sc[k] * (d - k1[k] - (v[k] - mu[k]) * k2[k]);
    }
  };
  p = tid;
  for (; p + (U - 1) * NT_SLICED < pixels; p += U * NT_SLICED) {
    float g[U][N], v[U][N];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t o = (size_t)(p + u * NT_SLICED) * C + c;
      V16<T>::ld(gsrc + o, g[u]);
      V16<T>::ld(y + o, v[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float o[N];
      pass2(g[u], v[u], o);
      V16<T>::st(dy + (size_t)(p + u * NT_SLICED) * C + c, o);
    }
  }
  for (; p < pixels; p += NT_SLICED) {
    float g[N], v[N], o[N];
    const size_t off = (size_t)p * C + c;
    V16<T>::ld(gsrc + off, g);
    V16<T>::ld(y + off, v);
    pass2(g, v, o);
    V16<T>::st(dy + off, o);
  }
}

template <typename T>
__global__ void leaky_bwd_wide(const T* __restrict__ dout, const T* __restrict__ out, T* __restrict__ din, int64_t nvec) {
  constexpr int N = V16<T>::N;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (; i + 3 * stride < nvec; i += 4 * stride) {            // 8 loads in flight per thread
    float g[4][N], o[4][N];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      V16<T>::ld(dout + (i + u * stride) * N, g[u]);
      V16<T>::ld(out + (i + u * stride) * N, o[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < N; ++k) g[u][k] = o[u][k] > 0.f ? g[u][k] : kLeakySlope * g[u][k];
      V16<T>::st(din + (i + u * stride) * N, g[u]);
    }
  }
  for (; i < nvec; i += stride) {
    float g[N], o[N];
    V16<T>::ld(dout + i * N, g);
    V16<T>::ld(out + i * N, o);
#pragma unroll
    for (int k = 0; k < N; ++k) g[k] = o[k] > 0.f ? g[k] : kLeakySlope * g[k];
    V16<T>::st(din + i * N, g);
  }
}
template <typename T>
__global__ void add_inplace_wide(T* __restrict__ acc, const T* __restrict__ x, int64_t nvec) {
  constexpr int N = V16<T>::N;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (; i + 3 * stride < nvec; i += 4 * stride) {            // 8 loads in flight per thread
    float a[4][N], b[4][N];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      V16<T>::ld(acc + (i + u * stride) * N, a[u]);
      V16<T>::ld(x + (i + u * stride) * N, b[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < N; ++k) a[u][k] += b[u][k];
      V16<T>::st(acc + (i + u * stride) * N, a[u]);
    }
  }
  for (; i < nvec; i += stride) {
    float a[N], b[N];
    V16<T>::ld(acc + i * N, a);
    V16<T>::ld(x + i * N, b);
#pragma unroll
    for (int k = 0; k < N; ++k) a[k] += b[k];
    V16<T>::st(acc + i * N, a);
  }
}

// ------------------------------------------------------------------ gated fusion
template <typename T>
__global__ void gate_fwd_kernel(const T* __restrict__ y, const float* __restrict__ scale,
                                const float* __restrict__ shift, const T* __restrict__ img,
                                T* __restrict__ out, int64_t nvec, int C) {
  const int CV = C >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / CV;
    const int c = (int)(i - p * CV) * 4;
    float4 a = Vec4<T>::ld(y + p * 2 * C + c);
    float4 b = Vec4<T>::ld(y + p * 2 * C + C + c);
    if (scale) {
      const float4 sa = *reinterpret_cast<const float4*>(scale + c), ba = *reinterpret_cast<const float4*>(shift + c);
      const float4 sb = *reinterpret_cast<const float4*>(scale + C + c), bb = *reinterpret_cast<const float4*>(shift + C + c);
      a.x = fmaf(a.x, sa.x, ba.x); a.y = fmaf(a.y, sa.y, ba.y); a.z = fmaf(a.z, sa.z, ba.z); a.w = fmaf(a.w, sa.w, ba.w);
      b.x = fmaf(b.x, sb.x, bb.x); b.y = fmaf(b.y, sb.y, bb.y); b.z = fmaf(b.z, sb.z, bb.z); b.w = fmaf(b.w, sb.w, bb.w);
    }
    const float4 im = Vec4<T>::ld(img + p * C + c);
    float4 o;
    o.x = fmaf(sigmoid_precise(a.x), b.x, im.x);
    o.y = fmaf(sigmoid_precise(a.y), b.y, im.y);
    o.z = fmaf(sigmoid_precise(a.z), b.z, im.z);
    o.w = fmaf(sigmoid_precise(a.w), b.w, im.w);
    Vec4<T>::st(out + p * C + c, o);
  }
}

template <typename T>
__global__ void gate_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ y,
                                const float* __restrict__ scale, const float* __restrict__ shift,
                                T* __restrict__ dzy, int64_t nvec, int C) {
  const int CV = C >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / CV;
    const int c = (int)(i - p * CV) * 4;
    float4 a = Vec4<T>::ld(y + p * 2 * C + c);
    float4 b = Vec4<T>::ld(y + p * 2 * C + C + c);
    if (scale) {
      const float4 sa = *reinterpret_cast<const float4*>(scale + c), ba = *reinterpret_cast<const float4*>(shift + c);
      const float4 sb = *reinterpret_cast<const float4*>(scale + C + c), bb = *reinterpret_cast<const float4*>(shift + C + c);
      a.x = fmaf(a.x, sa.x, ba.x); a.y = fmaf(a.y, sa.y, ba.y); a.z = fmaf(a.z, sa.z, ba.z); a.w = fmaf(a.w, sa.w, ba.w);
      b.x = fmaf(b.x, sb.x, bb.x); b.y = fmaf(b.y, sb.y, bb.y); b.z = fmaf(b.z, sb.z, bb.z); b.w = fmaf(b.w, sb.w, bb.w);
    }
    const float4 g = Vec4<T>::ld(dout + p * C + c);
    float4 da, db;
    float s;
    s = sigmoid_precise(a.x); da.x = g.x * b.x * s * (1.f - s); db.x = g.x * s;
    s = sigmoid_precise(a.y); da.y = g.y * b.y * s * (1.f - s); db.y = g.y * s;
    s = sigmoid_precise(a.z); da.z = g.z * b.z * s * (1.f - s); db.z = g.z * s;
    s = sigmoid_precise(a.w); da.w = g.w * b.w * s * (1.f - s); db.w = g.w * s;
    Vec4<T>::st(dzy + p * 2 * C + c, da);
    Vec4<T>::st(dzy + p * 2 * C + C + c, db);
  }
}

// ------------------------------------------------------------------ max-pool 3x3 s2 p1
template <typename T>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, int N, int H, int W,
                                   int C, int HO, int WO) {
  const int CV = C >> 2;
  const int64_t total = (int64_t)N * HO * WO * CV;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    int64_t r = i / CV;
    const int ox = (int)(r % WO); r /= WO;
    const int oy = (int)(r % HO);
    const int n = (int)(r / HO);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int dy = 0; dy < 3; ++dy) {
      const int iy = oy * 2 - 1 + dy;
      if (iy < 0 || iy >= H) continue;
      for (int dx = 0; dx < 3; ++dx) {
        const int ix = ox * 2 - 1 + dx;
        if (ix < 0 || ix >= W) continue;
        const float4 v = Vec4<T>::ld(x + ((size_t)(n * H + iy) * W + ix) * C + cv * 4);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    Vec4<T>::st(out + i * 4, m);
  }
}

// Training forward: also records WHERE the maximum came from (window position dy * 3 + dx of the FIRST maximum in
// row-major scan order, like ATen; 255 = empty / all -inf), one byte per output element, so that the backward pass
// reads dout + one byte instead of re-scanning 9 inputs per window.  16-byte channel vectors.
template <typename T>
__global__ void maxpool_fwd_idx_kernel(const T* __restrict__ x, T* __restrict__ out, uint8_t* __restrict__ idx, int N, int H,
                                       int W, int C, int HO, int WO) {
  constexpr int NV = V16<T>::N;
  const int CV = C / NV;
  const int64_t total = (int64_t)N * HO * WO * CV;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    int64_t r = i / CV;
    const int ox = (int)(r % WO); r /= WO;
    const int oy = (int)(r % HO);
    const int n = (int)(r / HO);
    float best[NV];
    uint8_t arg[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) { best[c] = -INFINITY; arg[c] = 255; }
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int iy = oy * 2 - 1 + dy;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int ix = ox * 2 - 1 + dx;
        if (ix < 0 || ix >= W) continue;
        float v[NV];
        V16<T>::ld(x + (((size_t)(n * H + iy) * W + ix) * CV + cv) * NV, v);
#pragma unroll
        for (int c = 0; c < NV; ++c)
          if (v[c] > best[c]) { best[c] = v[c]; arg[c] = (uint8_t)(dy * 3 + dx); }
      }
    }
    V16<T>::st(out + i * NV, best);
#pragma unroll
    for (int c = 0; c < NV; c += 4)
      *reinterpret_cast<uint32_t*>(idx + i * NV + c) = (uint32_t)arg[c] | ((uint32_t)arg[c + 1] << 8) | ((uint32_t)arg[c + 2] << 16) |
                                                       ((uint32_t)arg[c + 3] << 24);
  }
}
// Backward from the recorded positions: one thread owns a 2x2 block of input pixels x 16 bytes of channels and looks at
// the <= 4 windows that touch it (dout vector + position bytes each).
template <typename T>
__global__ void maxpool_bwd_idx_kernel(const T* __restrict__ dout, const uint8_t* __restrict__ idx, T* __restrict__ dx, int N,
                                       int H, int W, int C, int HO, int WO) {
  constexpr int NV = V16<T>::N;
  const int CV = C / NV;
  const int HB = (H + 1) >> 1, WB = (W + 1) >> 1;
  const int64_t total = (int64_t)N * HB * WB * CV;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    int64_t r = i / CV;
    const int m = (int)(r % WB); r /= WB;
    const int k = (int)(r % HB);
    const int n = (int)(r / HB);
    float g[4][NV];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < NV; ++c) g[a][c] = 0.f;
#pragma unroll
    for (int wy = 0; wy < 2; ++wy) {
#pragma unroll
      for (int wx = 0; wx < 2; ++wx) {
        const int oy = k + wy, ox = m + wx;
        if (oy >= HO || ox >= WO) continue;
        const size_t o = (((size_t)(n * HO + oy) * WO + ox) * CV + cv) * NV;
        float d[NV];
        V16<T>::ld(dout + o, d);
        uint8_t arg[NV];
#pragma unroll
        for (int c = 0; c < NV; c += 4) {
          const uint32_t pk = *reinterpret_cast<const uint32_t*>(idx + o + c);
          arg[c] = pk & 0xff; arg[c + 1] = (pk >> 8) & 0xff; arg[c + 2] = (pk >> 16) & 0xff; arg[c + 3] = pk >> 24;
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int dy = a + 1 - 2 * wy;
          if (dy < 0) continue;
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const int dxx = b + 1 - 2 * wx;
            if (dxx < 0) continue;
#pragma unroll
            for (int c = 0; c < NV; ++c)
              if (arg[c] == dy * 3 + dxx) g[a * 2 + b][c] += d[c];
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int iy = 2 * k + a;
      if (iy >= H) continue;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int ix = 2 * m + b;
        if (ix >= W) continue;
        V16<T>::st(dx + (((size_t)(n * H + iy) * W + ix) * CV + cv) * NV, g[a * 2 + b]);
      }
    }
  }
}

// gather form: one thread owns a 2x2 block of input pixels x 16 bytes of channels and recomputes the arg-max
// (FIRST maximum in row-major window scan, like ATen) of the <= 4 windows that touch the block: rows 2k, 2k+1
// lie in windows k (window rows 1, 2) and k+1 (window row 0, odd input row only), same for columns.
template <typename T>
__global__ void maxpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dout,
                                   T* __restrict__ dx, int N, int H, int W, int C, int HO, int WO) {
  constexpr int NV = V16<T>::N;
  const int CV = C / NV;
  const int HB = (H + 1) >> 1, WB = (W + 1) >> 1;
  const int64_t total = (int64_t)N * HB * WB * CV;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    int64_t r = i / CV;
    const int m = (int)(r % WB); r /= WB;
    const int k = (int)(r % HB);
    const int n = (int)(r / HB);
    float g[4][NV];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < NV; ++c) g[a][c] = 0.f;
#pragma unroll
    for (int wy = 0; wy < 2; ++wy) {
#pragma unroll
      for (int wx = 0; wx < 2; ++wx) {
        const int oy = k + wy, ox = m + wx;
        if (oy >= HO || ox >= WO) continue;
        float best[NV];
        int arg[NV];
#pragma unroll
        for (int c = 0; c < NV; ++c) { best[c] = -INFINITY; arg[c] = -1; }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int yy = oy * 2 - 1 + dy;
          if (yy < 0 || yy >= H) continue;
#pragma unroll
          for (int dxx = 0; dxx < 3; ++dxx) {
            const int xx = ox * 2 - 1 + dxx;
            if (xx < 0 || xx >= W) continue;
            float v[NV];
            V16<T>::ld(x + (((size_t)(n * H + yy) * W + xx) * CV + cv) * NV, v);
#pragma unroll
            for (int c = 0; c < NV; ++c)
              if (v[c] > best[c]) { best[c] = v[c]; arg[c] = dy * 3 + dxx; }
          }
        }
        float d[NV];
        V16<T>::ld(dout + (((size_t)(n * HO + oy) * WO + ox) * CV + cv) * NV, d);
        // block pixel (a, b) sits at window position (dy, dx) = (a + 1 - 2 wy, b + 1 - 2 wx) when that is in [0, 3)
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int dy = a + 1 - 2 * wy;
          if (dy < 0) continue;
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const int dxx = b + 1 - 2 * wx;
            if (dxx < 0) continue;
#pragma unroll
            for (int c = 0; c < NV; ++c)
              if (arg[c] == dy * 3 + dxx) g[a * 2 + b][c] += d[c];
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int iy = 2 * k + a;
      if (iy >= H) continue;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int ix = 2 * m + b;
        if (ix >= W) continue;
        V16<T>::st(dx + (((size_t)(n * H + iy) * W + ix) * CV + cv) * NV, g[a * 2 + b]);
      }
    }
  }
}

// 4-channel fallback (C % 4 == 0 only): every input pixel looks at the <= 4 windows covering it.
template <typename T>
__global__ void maxpool_bwd_narrow_kernel(const T* __restrict__ x, const T* __restrict__ dout,
                                          T* __restrict__ dx, int N, int H, int W, int C, int HO, int WO) {
  const int CV = C >> 2;
  const int64_t total = (int64_t)N * H * W * CV;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    int64_t r = i / CV;
    const int ix = (int)(r % W); r /= W;
    const int iy = (int)(r % H);
    const int n = (int)(r / H);
    const float4 sv = Vec4<T>::ld(x + i * 4);
    const float self[4] = {sv.x, sv.y, sv.z, sv.w};
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    const int oy_lo = (iy + 1) / 2 - (((iy + 1) & 1) == 0 ? 1 : 0), oy_hi = (iy + 1) / 2;
    const int ox_lo = (ix + 1) / 2 - (((ix + 1) & 1) == 0 ? 1 : 0), ox_hi = (ix + 1) / 2;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      if (oy < 0 || oy >= HO) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        if (ox < 0 || ox >= WO) continue;
        bool win[4] = {true, true, true, true};
        for (int dy = 0; dy < 3; ++dy) {
          const int yy = oy * 2 - 1 + dy;
          if (yy < 0 || yy >= H) continue;
          for (int dxx = 0; dxx < 3; ++dxx) {
            const int xx = ox * 2 - 1 + dxx;
            if (xx < 0 || xx >= W || (yy == iy && xx == ix)) continue;
            const float4 vv = Vec4<T>::ld(x + (((size_t)(n * H + yy) * W + xx) * CV + cv) * 4);
            const float v[4] = {vv.x, vv.y, vv.z, vv.w};
            const bool before = (yy < iy) || (yy == iy && xx < ix);
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (v[c] > self[c] || (before && v[c] == self[c])) win[c] = false;
          }
        }
        const float4 dv = Vec4<T>::ld(dout + (((size_t)(n * HO + oy) * WO + ox) * CV + cv) * 4);
        const float d[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (win[c]) g[c] += d[c];
      }
    }
    Vec4<T>::st(dx + i * 4, make_float4(g[0], g[1], g[2], g[3]));
  }
}

// ------------------------------------------------------------------ nearest upsample backward
template <typename T>
__global__ void upsample_bwd_kernel(const T* __restrict__ dup, T* __restrict__ dsrc, int N, int HS,
                                    int WS, int HU, int WU, int C, float sch, float scw, int accumulate) {
  const int CV = C >> 2;
  const int64_t total = (int64_t)N * HS * WS * CV;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    int64_t r = i / CV;
    const int sx = (int)(r % WS); r /= WS;
    const int sy = (int)(r % HS);
    const int n = (int)(r / HS);
    int y0 = (int)floorf((float)sy / sch) - 2, y1 = (int)ceilf((float)(sy + 1) / sch) + 2;
    int x0 = (int)floorf((float)sx / scw) - 2, x1 = (int)ceilf((float)(sx + 1) / scw) + 2;
    if (y0 < 0) y0 = 0;
    if (x0 < 0) x0 = 0;
    if (y1 > HU - 1) y1 = HU - 1;
    if (x1 > WU - 1) x1 = WU - 1;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int y = y0; y <= y1; ++y) {
      if (nearest_src(y, sch, HS) != sy) continue;
      for (int x = x0; x <= x1; ++x) {
        if (nearest_src(x, scw, WS) != sx) continue;
        const float4 v = Vec4<T>::ld(dup + (((size_t)(n * HU + y) * WU + x) * CV + cv) * 4);
        g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
      }
    }
    if (accumulate) {
      const float4 o = Vec4<T>::ld(dsrc + i * 4);
      g.x += o.x; g.y += o.y; g.z += o.z; g.w += o.w;
    }
    Vec4<T>::st(dsrc + i * 4, g);
  }
}

// exact 2x case: the four up-sampled pixels of a source pixel, 16-byte vectors
template <typename T>
__global__ void upsample2x_bwd_wide(const T* __restrict__ dup, T* __restrict__ dsrc, int N, int HS, int WS, int C,
                                    int accumulate) {
  constexpr int NV = V16<T>::N;
  const int CV = C / NV;
  const int64_t total = (int64_t)N * HS * WS * CV;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  auto src = [&](int64_t i) {
    const int cv = (int)(i % CV);
    int64_t r = i / CV;
    const int sx = (int)(r % WS); r /= WS;
    const int sy = (int)(r % HS);
    const int n = (int)(r / HS);
    return dup + (((size_t)(n * 2 * HS + 2 * sy) * (2 * WS) + 2 * sx) * CV + cv) * NV;
  };
  auto finish = [&](int64_t i, float* a, const float* b, const float* c, const float* d) {
#pragma unroll
    for (int k = 0; k < NV; ++k) a[k] = ((a[k] + b[k]) + c[k]) + d[k];      // same order as the generic kernel
    if (accumulate) {
      float e[NV];
      V16<T>::ld(dsrc + i * NV, e);
#pragma unroll
      for (int k = 0; k < NV; ++k) a[k] += e[k];
    }
    V16<T>::st(dsrc + i * NV, a);
  };
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (; i + stride < total; i += 2 * stride) {               // two source pixels per trip: 8 loads in flight
    float a[2][NV], b[2][NV], c[2][NV], d[2][NV];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const T* base = src(i + u * stride);
      V16<T>::ld(base, a[u]);
      V16<T>::ld(base + C, b[u]);
      V16<T>::ld(base + (size_t)2 * WS * C, c[u]);
      V16<T>::ld(base + (size_t)2 * WS * C + C, d[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) finish(i + u * stride, a[u], b[u], c[u], d[u]);
  }
  for (; i < total; i += stride) {
    const T* base = src(i);
    float a[NV], b[NV], c[NV], d[NV];
    V16<T>::ld(base, a);
    V16<T>::ld(base + C, b);
    V16<T>::ld(base + (size_t)2 * WS * C, c);
    V16<T>::ld(base + (size_t)2 * WS * C + C, d);
    finish(i, a, b, c, d);
  }
}

// ------------------------------------------------------------------ small elementwise
template <typename T>
__global__ void leaky_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ out, T* __restrict__ din, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = to_f<T>(dout[i]);
    din[i] = from_f<T>(to_f<T>(out[i]) > 0.f ? g : kLeakySlope * g);
  }
}
template <typename T>
__global__ void add_inplace_kernel(T* __restrict__ acc, const T* __restrict__ x, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc[i] = from_f<T>(to_f<T>(acc[i]) + to_f<T>(x[i]));
}

template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int N, int C, int H, int W,
                                    int CP) {
  const int64_t total = (int64_t)N * CP * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % CP);
    int64_t r = i / CP;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const int n = (int)(r / H);
    dst[i] = from_f<T>(c < C ? src[((size_t)(n * C + c) * H + y) * W + x] : 0.f);
  }
}
// space-to-depth + layout + precision in one pass: thread = one output pixel (2x2 input patch, all channels)
template <typename T>
__global__ void nchw_to_s2d_kernel(const float* __restrict__ src, T* __restrict__ dst, int N, int C, int H, int W, int CP) {
  const int H2 = H >> 1, W2 = W >> 1;
  const int64_t total = (int64_t)N * H2 * W2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x2 = (int)(i % W2);
    int64_t r = i / W2;
    const int y2 = (int)(r % H2);
    const int n = (int)(r / H2);
    T* o = dst + i * CP;
    if (CP == 16 && C <= 4) {                 // the stems: whole pixel in registers, 16-byte stores
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < C) {
          const float* s = src + (((size_t)n * C + c) * H + 2 * y2) * W + 2 * x2;
          const float2 top = *reinterpret_cast<const float2*>(s);
          const float2 bot = *reinterpret_cast<const float2*>(s + W);
#pragma unroll
          for (int k = 0; k < 16; ++k) {      // static register indices: select instead of dynamic indexing
            if (k == 0 * C + c) v[k] = top.x;
            if (k == 1 * C + c) v[k] = top.y;
            if (k == 2 * C + c) v[k] = bot.x;
            if (k == 3 * C + c) v[k] = bot.y;
          }
        }
      }
      constexpr int N = V16<T>::N;
#pragma unroll
      for (int k = 0; k < 16; k += N) V16<T>::st(o + k, v + k);
      continue;
    }
    for (int c = 0; c < C; ++c) {
      const float* s = src + (((size_t)n * C + c) * H + 2 * y2) * W + 2 * x2;
      const float2 top = *reinterpret_cast<const float2*>(s);
      const float2 bot = *reinterpret_cast<const float2*>(s + W);
      o[0 * C + c] = from_f<T>(top.x);
      o[1 * C + c] = from_f<T>(top.y);
      o[2 * C + c] = from_f<T>(bot.x);
      o[3 * C + c] = from_f<T>(bot.y);
    }
    for (int c = 4 * C; c < CP; ++c) o[c] = from_f<T>(0.f);
  }
}
template <typename T>
__global__ void pack_stem_s2d_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int C, int CP) {
  const int64_t total = (int64_t)cout * 16 * CP;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % CP);
    int64_t r = i / CP;
    const int tap = (int)(r % 16);
    const int co = (int)(r / 16);
    float v = 0.f;
    if (ch < 4 * C) {
      const int ph = ch / C, c = ch - ph * C;
      const int rr = 2 * (tap >> 2) + (ph >> 1) - 1, ss = 2 * (tap & 3) + (ph & 1) - 1;
      if (rr >= 0 && rr < 7 && ss >= 0 && ss < 7) v = w[(((size_t)co * C + c) * 7 + rr) * 7 + ss];
    }
    out[i] = from_f<T>(v);
  }
}
__global__ void unpack_stem_s2d_kernel(const float* __restrict__ packed, float* __restrict__ g, int cout, int C, int CP) {
  const int64_t total = (int64_t)cout * C * 49;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ss = (int)(i % 7);
    int64_t r = i / 7;
    const int rr = (int)(r % 7); r /= 7;
    const int c = (int)(r % C);
    const int co = (int)(r / C);
    const int ty = (rr + 1) >> 1, dy = (rr + 1) & 1, tx = (ss + 1) >> 1, dx = (ss + 1) & 1;
    g[i] = packed[((size_t)co * 16 + ty * 4 + tx) * CP + (dy * 2 + dx) * C + c];
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int N, int C, int H, int W) {
  const int64_t total = (int64_t)N * C * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    int64_t r = i / W;
    const int y = (int)(r % H); r /= H;
    const int c = (int)(r % C);
    const int n = (int)(r / C);
    dst[i] = to_f<T>(src[((size_t)(n * H + y) * W + x) * C + c]);
  }
}

template <typename T>
__global__ void depth_head_bwd_kernel(const float* __restrict__ dd, const float* __restrict__ d,
                                      T* __restrict__ dl, float mn, float r, int64_t n, int CP) {
  constexpr int N = V16<T>::N;
  for (int64_t px = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; px < n; px += (int64_t)gridDim.x * blockDim.x) {
    const float dv = d[px];
    const float s = mn / dv - r;              // sigmoid(logit)
    const float g = -dd[px] * dv * dv / mn * s * (1.f - s);
    T* o = dl + px * CP;
    if (CP % N == 0) {                        // channel 0 carries the gradient, the padding channels are zero
      float v[N];
#pragma unroll
      for (int k = 0; k < N; ++k) v[k] = 0.f;
      for (int c = N; c < CP; c += N) V16<T>::st(o + c, v);
      v[0] = g;
      V16<T>::st(o, v);
    } else {
      o[0] = from_f<T>(g);
      for (int c = 1; c < CP; ++c) o[c] = from_f<T>(0.f);
    }
  }
}

// ------------------------------------------------------------------ masked L1 loss
__global__ void l1_accum_kernel(const float* __restrict__ out, const float* __restrict__ gt,
                                const float* __restrict__ lidar, double* __restrict__ accum, int64_t n) {
  float sg = 0.f, cg = 0.f, sl = 0.f, cl = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float o = out[i], l = lidar[i];
    const float g = l > 0.f ? 0.f : gt[i];
    if (g > 0.f) { sg += fabsf(o - g); cg += 1.f; }
    if (l > 0.f) { sl += fabsf(o - l); cl += 1.f; }
  }
  __shared__ double red[4][NT / 32];
  float v[4] = {sg, cg, sl, cl};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float x = v[j];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[j][threadIdx.x >> 5] = (double)x;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < NT / 32; ++w) t += red[threadIdx.x][w];
    atomicAdd(accum + threadIdx.x, t);
  }
}
__global__ void l1_finish_kernel(const float* __restrict__ out, const float* __restrict__ gt,
                                 const float* __restrict__ lidar, const double* __restrict__ accum,
                                 float w_lidar, float* __restrict__ loss, float* __restrict__ dout, int64_t n) {
  const float inv_g = (float)(1.0 / accum[1]);
  const float inv_l = (float)(1.0 / accum[3]);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float L = (float)(accum[0] / accum[1]);
    if (w_lidar > 0.f) L += w_lidar * (float)(accum[2] / accum[3]);
    loss[0] = L;
  }
  if (!dout) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float o = out[i], l = lidar[i];
    const float g = (w_lidar > 0.f && l > 0.f) ? 0.f : gt[i];
    float d = 0.f;
    if (g > 0.f) d += (o > g ? 1.f : (o < g ? -1.f : 0.f)) * inv_g;
    if (w_lidar > 0.f && l > 0.f) d += w_lidar * (o > l ? 1.f : (o < l ? -1.f : 0.f)) * inv_l;
    dout[i] = d;
  }
}

// ------------------------------------------------------------------ outlier removal
__global__ void max_kernel(const float* __restrict__ x, float* __restrict__ mx, int64_t n) {
  float m = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, x[i]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(mx), __float_as_int(m));   // values >= 0
}
// separable min filter on a shared-memory tile: 32 x 8 outputs per block, halo ks/2 (ks <= 15)
constexpr int OUT_TW = 32, OUT_TH = 8, OUT_MAXPAD = 7;
__global__ void outlier_kernel(const float* __restrict__ d, const float* __restrict__ mx,
                               float* __restrict__ out, int N, int H, int W, int ks, float thr) {
  __shared__ float tile[OUT_TH + 2 * OUT_MAXPAD][OUT_TW + 2 * OUT_MAXPAD + 1];
  __shared__ float rowmin[OUT_TH + 2 * OUT_MAXPAD][OUT_TW];
  const float fill = 10.f * mx[0];
  const int pad = ks / 2;
  const int n = blockIdx.z, y0 = blockIdx.y * OUT_TH, x0 = blockIdx.x * OUT_TW;
  const int th = OUT_TH + 2 * pad, tw = OUT_TW + 2 * pad;
  const float* img = d + (size_t)n * H * W;
  for (int i = threadIdx.x; i < th * tw; i += blockDim.x) {
    const int ty = i / tw, tx = i - ty * tw;
    const int yy = y0 + ty - pad, xx = x0 + tx - pad;
    float v = fill;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      v = img[(size_t)yy * W + xx];
      v = v > 0.f ? v : fill;
    }
    tile[ty][tx] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < th * OUT_TW; i += blockDim.x) {
    const int ty = i / OUT_TW, tx = i - ty * OUT_TW;
    float m = fill;
    for (int dx = 0; dx < ks; ++dx) m = fminf(m, tile[ty][tx + dx]);
    rowmin[ty][tx] = m;
  }
  __syncthreads();
  const int tx = threadIdx.x % OUT_TW, ty = threadIdx.x / OUT_TW;
  const int y = y0 + ty, x = x0 + tx;
  if (y < H && x < W) {
    float m = fill;
    for (int dy = 0; dy < ks; ++dy) m = fminf(m, rowmin[ty + dy][tx]);
    const float v = img[(size_t)y * W + x];
    out[(size_t)n * H * W + (size_t)y * W + x] = (m < v - thr) ? v * 0.f : v;
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float step_size, float b1, float b2,
                            float eps, float inv_sqrt_bc2) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);         // torch lerp form
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}

// ------------------------------------------------------------------ weight (un)packing
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin,
                                   int kh, int kw, int cin_off, int cin_cnt, int cpad, int mode) {
  const int taps = kh * kw;
  if (mode == 0) {            // out[co][tap][ci < cpad]
    const int64_t total = (int64_t)cout * taps * cpad;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int ci = (int)(i % cpad);
      int64_t r = i / cpad;
      const int tap = (int)(r % taps);
      const int co = (int)(r / taps);
      out[i] = from_f<T>(ci < cin_cnt ? w[((size_t)co * cin + cin_off + ci) * taps + tap] : 0.f);
    }
  } else {                    // out[ci][flipped tap][co < copad]
    const int copad = cpad > cout ? cpad : cout;
    const int64_t total = (int64_t)cin_cnt * taps * copad;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int co = (int)(i % copad);
      int64_t r = i / copad;
      const int tap = (int)(r % taps);
      const int ci = (int)(r / taps);
      // rows past the real input channels (a source stored with zero-padded channels) are zero
      out[i] = from_f<T>((co < cout && cin_off + ci < cin) ? w[((size_t)co * cin + cin_off + ci) * taps + (taps - 1 - tap)] : 0.f);
    }
  }
}
template <typename T>
__global__ void pack_up2x_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin) {
  const int64_t total = (int64_t)4 * cout * 4 * cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    int64_t r = i / cin;
    const int tap = (int)(r % 4); r /= 4;
    const int co = (int)(r % cout);
    const int ph = (int)(r / cout);
    const int a = ph >> 1, b = ph & 1, t = tap >> 1, u = tap & 1;
    // 3x3 rows folded onto low-res row t of phase a: a=0: {0},{1,2}; a=1: {0,1},{2}
    const int r0 = a == 0 ? (t == 0 ? 0 : 1) : (t == 0 ? 0 : 2), r1 = a == 0 ? (t == 0 ? 0 : 2) : (t == 0 ? 1 : 2);
    const int s0 = b == 0 ? (u == 0 ? 0 : 1) : (u == 0 ? 0 : 2), s1 = b == 0 ? (u == 0 ? 0 : 2) : (u == 0 ? 1 : 2);
    const float* wp = w + ((size_t)co * cin + ci) * 9;
    float acc = 0.f;
    for (int rr = r0; rr <= r1; ++rr)
      for (int ss = s0; ss <= s1; ++ss) acc += wp[rr * 3 + ss];
    out[i] = from_f<T>(acc);
  }
}
// phase weights of the stride-2 data gradient (conv_tma_dgrad_s2_supported): out[phase a*2+b][ci][tap t*2+u][co < copad]
template <typename T>
__global__ void pack_dgrad_s2_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin, int cin_off,
                                     int cin_cnt, int copad) {
  const int64_t total = (int64_t)16 * cin_cnt * copad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % copad);
    int64_t r = i / copad;
    const int tap = (int)(r % 4); r /= 4;
    const int ci = (int)(r % cin_cnt);
    const int ph = (int)(r / cin_cnt);
    out[i] = from_f<T>(dgrad_s2_weight(w, cout, cin, cin_off + ci, co, ph, tap));
  }
}
template <typename T>
__global__ void pack_upconv_dgrad_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin, int cin_off,
                                         int cin_cnt, int copad) {
  const int64_t total = (int64_t)cin_cnt * 16 * copad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % copad);
    const int64_t r = i / copad;
    out[i] = from_f<T>(upconv_dgrad_weight(w, cout, cin, cin_off + (int)(r / 16), co, (int)(r % 16)));
  }
}
__global__ void unpack_wgrad_kernel(const float* __restrict__ packed, float* __restrict__ g, int cout,
                                    int cin, int kh, int kw, int cin_off, int cin_cnt, int cpad, int accumulate) {
  const int taps = kh * kw;
  const int64_t total = (int64_t)cout * taps * cin_cnt;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin_cnt);
    int64_t r = i / cin_cnt;
    const int tap = (int)(r % taps);
    const int co = (int)(r / taps);
    const size_t o = ((size_t)co * cin + cin_off + ci) * taps + tap;
    const float v = packed[((size_t)co * taps + tap) * cpad + ci];
    g[o] = accumulate ? g[o] + v : v;
  }
}

}  // namespace
}  // namespace rcfd

using namespace rcfd;

#define DISPATCH_T(dtype, ...)                                   \
  if ((dtype) == RCFD_F32) { typedef float T; __VA_ARGS__; }     \
  else if ((dtype) == RCFD_BF16) { typedef bf16 T; __VA_ARGS__; } \
  else { set_error("bad dtype %d", (int)(dtype)); return RCFD_EINVAL; }

extern "C" {

int rcfd_bn_finalize(const double* sum, const double* sqsum, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, float* scale, float* shift,
                     float* save_mean, float* save_invstd, int32_t channels, int64_t count, float eps,
                     float momentum, void* stream) {
  RCFD_CHECK_ARG(sum && sqsum && gamma && beta && scale && shift && channels > 0 && count > 0, "bn_finalize: bad args");
  RCFD_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_finalize: running stats must come together");
  bn_finalize_kernel<<<ceil_div(channels, 128), 128, 0, (cudaStream_t)stream>>>(
      sum, sqsum, gamma, beta, running_mean, running_var, scale, shift, save_mean, save_invstd, channels,
      (double)count, eps, momentum);
  RCFD_CHECK_LAUNCH("bn_finalize");
  return RCFD_OK;
}

int rcfd_bn_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                 float* scale, float* shift, int32_t channels, float eps, void* stream) {
  RCFD_CHECK_ARG(gamma && beta && running_mean && running_var && scale && shift && channels > 0, "bn_fold: bad args");
  bn_fold_kernel<<<ceil_div(channels, 128), 128, 0, (cudaStream_t)stream>>>(gamma, beta, running_mean, running_var,
                                                                            scale, shift, channels, eps);
  RCFD_CHECK_LAUNCH("bn_fold");
  return RCFD_OK;
}

int rcfd_bn_act_fwd(const void* y, const float* scale, const float* shift, const void* residual, void* out,
                    int64_t pixels, int32_t channels, int32_t act, int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(y && out && pixels > 0 && channels > 0 && channels % 4 == 0, "bn_act_fwd: bad args (channels %% 4)");
  RCFD_CHECK_ARG((scale == nullptr) == (shift == nullptr), "bn_act_fwd: scale/shift");
  const int vw = dtype == RCFD_BF16 ? 8 : 4;
  if (channels % vw == 0 && NT % (channels / vw) == 0) {
    const int64_t nv = pixels * channels / vw;
    DISPATCH_T(dtype, (bn_act_fwd_wide<T><<<grid_for(nv, NT * g_bn_fwd_vectors_per_thread, 148 * 8), NT, 0, (cudaStream_t)stream>>>(
                          (const T*)y, scale, shift, (const T*)residual, (T*)out, nv, channels, act)));
    RCFD_CHECK_LAUNCH("bn_act_fwd");
    return RCFD_OK;
  }
  const int64_t nvec = pixels * channels / 4;
  DISPATCH_T(dtype, (bn_act_fwd_kernel<T><<<grid_for(nvec), NT, 0, (cudaStream_t)stream>>>(
                        (const T*)y, scale, shift, (const T*)residual, (T*)out, nvec, channels, act)));
  RCFD_CHECK_LAUNCH("bn_act_fwd");
  return RCFD_OK;
}

int rcfd_bn_train_act_fwd(const void* y, const double* sum, const double* sqsum, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, float* scale, float* shift, float* save_mean,
                          float* save_invstd, const void* residual, void* out, int64_t pixels, int32_t channels,
                          int32_t act, float eps, float momentum, int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(y && sum && sqsum && gamma && beta && scale && shift && save_mean && save_invstd && out && pixels > 0 &&
                     channels > 0 && channels % 4 == 0,
                 "bn_train_act_fwd: bad args");
  const int vw = dtype == RCFD_BF16 ? 8 : 4;
  if (channels % vw == 0 && NT % (channels / vw) == 0 && channels <= BN_TRAIN_MAX_C) {
    const int64_t nv = pixels * channels / vw;
    DISPATCH_T(dtype, (bn_train_act_fwd_wide<T><<<grid_for(nv, NT * g_bn_fwd_vectors_per_thread, 148 * 8), NT, 0, (cudaStream_t)stream>>>(
                          (const T*)y, sum, sqsum, gamma, beta, running_mean, running_var, scale, shift, save_mean,
                          save_invstd, (const T*)residual, (T*)out, nv, channels, act, (double)pixels, eps, momentum)));
    RCFD_CHECK_LAUNCH("bn_train_act_fwd");
    return RCFD_OK;
  }
  int rc = rcfd_bn_finalize(sum, sqsum, gamma, beta, running_mean, running_var, scale, shift, save_mean, save_invstd,
                            channels, pixels, eps, momentum, stream);
  if (rc != RCFD_OK) return rc;
  return rcfd_bn_act_fwd(y, scale, shift, residual, out, pixels, channels, act, dtype, stream);
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

static int bn_bwd_reduce_impl(const void* dz, const void* y, const float* scale, const float* shift,
                              const float* mean, const float* invstd, double* sums, int64_t pixels,
                              int32_t channels, int32_t act, int32_t dtype, void* stream, bool zero_first) {
  RCFD_CHECK_ARG(dz && y && scale && shift && mean && invstd && sums, "bn_bwd_reduce: null");
  RCFD_CHECK_ARG(channels % 4 == 0 && channels <= 1024 && channels > 0 && pixels > 0, "bn_bwd_reduce: channels");
  if (zero_first) {
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * channels, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("bn_bwd_reduce memset: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
  }
  const int vw = dtype == RCFD_BF16 ? 8 : 4;
  const bool wide = channels % vw == 0;
  const int CVP = next_pow2(channels / (wide ? vw : 4));
  const int rstep = NT / CVP;
  const int rpt = g_bn_reduce_rows_per_thread;
  int blocks = (int)((pixels + rstep * rpt - 1) / (rstep * rpt));
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  int64_t rpb = (pixels + blocks - 1) / blocks;
  blocks = (int)((pixels + rpb - 1) / rpb);
  if (wide) {
    DISPATCH_T(dtype, (bn_bwd_reduce_wide<T><<<blocks, NT, 0, (cudaStream_t)stream>>>(
                          (const T*)dz, (const T*)y, scale, shift, mean, invstd, sums, pixels, channels, CVP, act, rpb)));
  } else {
    DISPATCH_T(dtype, (bn_bwd_reduce_kernel<T><<<blocks, NT, 0, (cudaStream_t)stream>>>(
                          (const T*)dz, (const T*)y, scale, shift, mean, invstd, sums, pixels, channels, CVP, act, rpb)));
  }
  RCFD_CHECK_LAUNCH("bn_bwd_reduce");
  return RCFD_OK;
}

int rcfd_bn_act_bwd_reduce(const void* dz, const void* y, const float* scale, const float* shift,
                           const float* mean, const float* invstd, double* sums, int64_t pixels,
                           int32_t channels, int32_t act, int32_t dtype, void* stream) {
  return bn_bwd_reduce_impl(dz, y, scale, shift, mean, invstd, sums, pixels, channels, act, dtype, stream, true);
}

int rcfd_bn_act_bwd_reduce_acc(const void* dz, const void* y, const float* scale, const float* shift,
                               const float* mean, const float* invstd, double* sums, int64_t pixels,
                               int32_t channels, int32_t act, int32_t dtype, void* stream) {
  return bn_bwd_reduce_impl(dz, y, scale, shift, mean, invstd, sums, pixels, channels, act, dtype, stream, false);
}

int rcfd_bn_act_bwd_apply(const void* dz, const void* y, const float* scale, const float* shift,
                          const float* mean, const float* invstd, const double* sums, void* dy, float* dgamma,
                          float* dbeta, int64_t pixels, int32_t channels, int32_t act, int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(dz && y && scale && shift && mean && invstd && sums && dy, "bn_bwd_apply: null");
  RCFD_CHECK_ARG(channels % 4 == 0 && channels > 0 && pixels > 0, "bn_bwd_apply: channels");
  const int vw = dtype == RCFD_BF16 ? 8 : 4;
  if (channels % vw == 0 && NT % (channels / vw) == 0) {
    const int64_t nv = pixels * channels / vw;
    DISPATCH_T(dtype, (bn_bwd_apply_wide<T><<<grid_for(nv, NT * g_bn_vectors_per_thread, 148 * 8), NT, 0, (cudaStream_t)stream>>>(
                          (const T*)dz, (const T*)y, scale, shift, mean, invstd, sums, (T*)dy, nv, channels, act,
                          (float)(1.0 / (double)pixels), dgamma, dbeta)));
  } else {
    const int64_t nvec = pixels * channels / 4;
    DISPATCH_T(dtype, (bn_bwd_apply_kernel<T><<<grid_for(nvec), NT, 0, (cudaStream_t)stream>>>(
                          (const T*)dz, (const T*)y, scale, shift, mean, invstd, sums, (T*)dy, nvec, channels, act,
                          (float)(1.0 / (double)pixels), dgamma, dbeta)));
  }
  RCFD_CHECK_LAUNCH("bn_bwd_apply");
  return RCFD_OK;
}

int rcfd_bn_act_bwd_fused(const void* dz, const void* y, const void* post_z, void* dz_masked, const float* scale,
                          const float* shift, const float* mean, const float* invstd, void* dy, float* dgamma,
                          float* dbeta, int64_t pixels, int32_t channels, int32_t act, int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(dz && y && scale && shift && mean && invstd && dy, "bn_bwd_fused: null");
  RCFD_CHECK_ARG((post_z == nullptr) == (dz_masked == nullptr), "bn_bwd_fused: post_z and dz_masked go together");
  const int vw = dtype == RCFD_BF16 ? 8 : 4;
  RCFD_CHECK_ARG(channels > 0 && channels % vw == 0, "bn_bwd_fused: channels must be a multiple of %d", vw);
  RCFD_CHECK_ARG(pixels > 0 && pixels <= (1 << 20), "bn_bwd_fused: 1 .. 2^20 pixels (small maps; use reduce + apply)");
  const float inv_count = (float)(1.0 / (double)pixels);
  if (post_z) {
    DISPATCH_T(dtype, (bn_bwd_sliced_kernel<T, true><<<channels / vw, NT_SLICED, 0, (cudaStream_t)stream>>>(
                          (const T*)dz, (const T*)y, (const T*)post_z, (T*)dz_masked, scale, shift, mean, invstd, (T*)dy,
                          dgamma, dbeta, (int)pixels, channels, act, inv_count)));
  } else {
    DISPATCH_T(dtype, (bn_bwd_sliced_kernel<T, false><<<channels / vw, NT_SLICED, 0, (cudaStream_t)stream>>>(
                          (const T*)dz, (const T*)y, nullptr, nullptr, scale, shift, mean, invstd, (T*)dy, dgamma, dbeta,
                          (int)pixels, channels, act, inv_count)));
  }
  RCFD_CHECK_LAUNCH("bn_bwd_fused");
  return RCFD_OK;
}

int rcfd_gate_fuse_fwd(const void* y, const float* scale, const float* shift, const void* img, void* out,
                       int64_t pixels, int32_t channels, int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(y && img && out && pixels > 0 && channels > 0 && channels % 4 == 0, "gate_fwd: bad args");
  const int64_t nvec = pixels * channels / 4;
  DISPATCH_T(dtype, (gate_fwd_kernel<T><<<grid_for(nvec), NT, 0, (cudaStream_t)stream>>>(
                        (const T*)y, scale, shift, (const T*)img, (T*)out, nvec, channels)));
  RCFD_CHECK_LAUNCH("gate_fwd");
  return RCFD_OK;
}

int rcfd_gate_fuse_bwd(const void* dout, const void* y, const float* scale, const float* shift, void* dz_y,
                       int64_t pixels, int32_t channels, int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(dout && y && dz_y && pixels > 0 && channels > 0 && channels % 4 == 0, "gate_bwd: bad args");
  const int64_t nvec = pixels * channels / 4;
  DISPATCH_T(dtype, (gate_bwd_kernel<T><<<grid_for(nvec), NT, 0, (cudaStream_t)stream>>>(
                        (const T*)dout, (const T*)y, scale, shift, (T*)dz_y, nvec, channels)));
  RCFD_CHECK_LAUNCH("gate_bwd");
  return RCFD_OK;
}

int rcfd_maxpool3x3s2_fwd(const void* x, void* out, int32_t n, int32_t h, int32_t w, int32_t c, int32_t dtype,
                          void* stream) {
  RCFD_CHECK_ARG(x && out && n > 0 && h > 0 && w > 0 && c > 0 && c % 4 == 0, "maxpool_fwd: bad args");
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  const int64_t total = (int64_t)n * ho * wo * (c / 4);
  DISPATCH_T(dtype, (maxpool_fwd_kernel<T><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>((const T*)x, (T*)out, n, h,
                                                                                           w, c, ho, wo)));
  RCFD_CHECK_LAUNCH("maxpool_fwd");
  return RCFD_OK;
}

int rcfd_maxpool3x3s2_bwd(const void* x, const void* dout, void* dx, int32_t n, int32_t h, int32_t w, int32_t c,
                          int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(x && dout && dx && n > 0 && h > 0 && w > 0 && c > 0 && c % 4 == 0, "maxpool_bwd: bad args");
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  const int vw = dtype == RCFD_BF16 ? 8 : 4;
  if (c % vw == 0) {
    const int64_t total = (int64_t)n * ((h + 1) / 2) * ((w + 1) / 2) * (c / vw);
    DISPATCH_T(dtype, (maxpool_bwd_kernel<T><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>(
                          (const T*)x, (const T*)dout, (T*)dx, n, h, w, c, ho, wo)));
  } else {
    const int64_t total = (int64_t)n * h * w * (c / 4);
    DISPATCH_T(dtype, (maxpool_bwd_narrow_kernel<T><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>(
                          (const T*)x, (const T*)dout, (T*)dx, n, h, w, c, ho, wo)));
  }
  RCFD_CHECK_LAUNCH("maxpool_bwd");
  return RCFD_OK;
}

int rcfd_maxpool3x3s2_fwd_idx(const void* x, void* out, uint8_t* idx, int32_t n, int32_t h, int32_t w, int32_t c,
                              int32_t dtype, void* stream) {
  const int vw = dtype == RCFD_BF16 ? 8 : 4;
  RCFD_CHECK_ARG(x && out && idx && n > 0 && h > 0 && w > 0 && c > 0 && c % vw == 0, "maxpool_fwd_idx: bad args (channels %% 16 bytes)");
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  const int64_t total = (int64_t)n * ho * wo * (c / vw);
  DISPATCH_T(dtype, (maxpool_fwd_idx_kernel<T><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>((const T*)x, (T*)out, idx, n, h, w,
                                                                                               c, ho, wo)));
  RCFD_CHECK_LAUNCH("maxpool_fwd_idx");
  return RCFD_OK;
}

int rcfd_maxpool3x3s2_bwd_idx(const void* dout, const uint8_t* idx, void* dx, int32_t n, int32_t h, int32_t w, int32_t c,
                              int32_t dtype, void* stream) {
  const int vw = dtype == RCFD_BF16 ? 8 : 4;
  RCFD_CHECK_ARG(dout && idx && dx && n > 0 && h > 0 && w > 0 && c > 0 && c % vw == 0, "maxpool_bwd_idx: bad args (channels %% 16 bytes)");
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  const int64_t total = (int64_t)n * ((h + 1) / 2) * ((w + 1) / 2) * (c / vw);
  DISPATCH_T(dtype, (maxpool_bwd_idx_kernel<T><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>((const T*)dout, idx, (T*)dx, n, h,
                                                                                               w, c, ho, wo)));
  RCFD_CHECK_LAUNCH("maxpool_bwd_idx");
  return RCFD_OK;
}

int rcfd_upsample_nearest_bwd(const void* dup, void* dsrc, int32_t n, int32_t hs, int32_t ws, int32_t hu,
                              int32_t wu, int32_t c, int32_t accumulate, int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(dup && dsrc && n > 0 && hs > 0 && ws > 0 && hu > 0 && wu > 0 && c > 0 && c % 4 == 0,
                 "upsample_bwd: bad args (channels %% 4)");
  const int vw = dtype == RCFD_BF16 ? 8 : 4;
  if (hu == 2 * hs && wu == 2 * ws && c % vw == 0) {
    const int64_t total = (int64_t)n * hs * ws * (c / vw);
    DISPATCH_T(dtype, (upsample2x_bwd_wide<T><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>(
                          (const T*)dup, (T*)dsrc, n, hs, ws, c, accumulate)));
    RCFD_CHECK_LAUNCH("upsample2x_bwd");
    return RCFD_OK;
  }
  const int64_t total = (int64_t)n * hs * ws * (c / 4);
  const float sch = (float)hs / (float)hu, scw = (float)ws / (float)wu;
  DISPATCH_T(dtype, (upsample_bwd_kernel<T><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>(
                        (const T*)dup, (T*)dsrc, n, hs, ws, hu, wu, c, sch, scw, accumulate)));
  RCFD_CHECK_LAUNCH("upsample_bwd");
  return RCFD_OK;
}

int rcfd_leaky_bwd(const void* dout, const void* out, void* din, int64_t count, int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(dout && out && din && count > 0, "leaky_bwd: bad args");
  const int vw = dtype == RCFD_BF16 ? 8 : 4;
  if (count % vw == 0) {
    DISPATCH_T(dtype, (leaky_bwd_wide<T><<<grid_for(count / vw, NT * g_ew_vectors_per_thread, 148 * 8), NT, 0, (cudaStream_t)stream>>>(
                          (const T*)dout, (const T*)out, (T*)din, count / vw)));
  } else {
    DISPATCH_T(dtype, (leaky_bwd_kernel<T><<<grid_for(count), NT, 0, (cudaStream_t)stream>>>((const T*)dout, (const T*)out,
                                                                                            (T*)din, count)));
  }
  RCFD_CHECK_LAUNCH("leaky_bwd");
  return RCFD_OK;
}

int rcfd_add_inplace(void* acc, const void* x, int64_t count, int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(acc && x && count > 0, "add_inplace: bad args");
  const int vw = dtype == RCFD_BF16 ? 8 : 4;
  if (count % vw == 0) {
    DISPATCH_T(dtype, (add_inplace_wide<T><<<grid_for(count / vw, NT * g_ew_vectors_per_thread, 148 * 8), NT, 0, (cudaStream_t)stream>>>(
                          (T*)acc, (const T*)x, count / vw)));
  } else {
    DISPATCH_T(dtype, (add_inplace_kernel<T><<<grid_for(count), NT, 0, (cudaStream_t)stream>>>((T*)acc, (const T*)x, count)));
  }
  RCFD_CHECK_LAUNCH("add_inplace");
  return RCFD_OK;
}

int rcfd_nchw_to_nhwc(const float* src, void* dst, int32_t n, int32_t c, int32_t h, int32_t w, int32_t cpad,
                      int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(src && dst && n > 0 && c > 0 && h > 0 && w > 0 && cpad >= c, "nchw_to_nhwc: bad args");
  const int64_t total = (int64_t)n * cpad * h * w;
  DISPATCH_T(dtype, (nchw_to_nhwc_kernel<T><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>(src, (T*)dst, n, c, h, w, cpad)));
  RCFD_CHECK_LAUNCH("nchw_to_nhwc");
  return RCFD_OK;
}

int rcfd_nchw_to_s2d_nhwc(const float* src, void* dst, int32_t n, int32_t c, int32_t h, int32_t w, int32_t cpad,
                          int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(src && dst && n > 0 && c > 0 && h > 0 && w > 0 && (h % 2 == 0) && (w % 2 == 0) && cpad >= 4 * c,
                 "nchw_to_s2d: bad args (h, w even; cpad >= 4c)");
  const int64_t total = (int64_t)n * (h / 2) * (w / 2);
  DISPATCH_T(dtype, (nchw_to_s2d_kernel<T><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>(src, (T*)dst, n, c, h, w, cpad)));
  RCFD_CHECK_LAUNCH("nchw_to_s2d");
  return RCFD_OK;
}

int rcfd_pack_stem_s2d_weight(const float* w_oihw, void* packed, int32_t cout, int32_t c, int32_t cpad, int32_t dtype,
                              void* stream) {
  RCFD_CHECK_ARG(w_oihw && packed && cout > 0 && c > 0 && cpad >= 4 * c, "pack_stem_s2d: bad args");
  const int64_t total = (int64_t)cout * 16 * cpad;
  DISPATCH_T(dtype, (pack_stem_s2d_kernel<T><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>(w_oihw, (T*)packed, cout, c, cpad)));
  RCFD_CHECK_LAUNCH("pack_stem_s2d");
  return RCFD_OK;
}

int rcfd_unpack_stem_s2d_wgrad(const float* packed, float* g_oihw, int32_t cout, int32_t c, int32_t cpad, void* stream) {
  RCFD_CHECK_ARG(packed && g_oihw && cout > 0 && c > 0 && cpad >= 4 * c, "unpack_stem_s2d: bad args");
  const int64_t total = (int64_t)cout * c * 49;
  unpack_stem_s2d_kernel<<<grid_for(total), NT, 0, (cudaStream_t)stream>>>(packed, g_oihw, cout, c, cpad);
  RCFD_CHECK_LAUNCH("unpack_stem_s2d");
  return RCFD_OK;
}

int rcfd_nhwc_to_nchw(const void* src, float* dst, int32_t n, int32_t c, int32_t h, int32_t w, int32_t dtype,
                      void* stream) {
  RCFD_CHECK_ARG(src && dst && n > 0 && c > 0 && h > 0 && w > 0, "nhwc_to_nchw: bad args");
  const int64_t total = (int64_t)n * c * h * w;
  DISPATCH_T(dtype, (nhwc_to_nchw_kernel<T><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>((const T*)src, dst, n, c, h, w)));
  RCFD_CHECK_LAUNCH("nhwc_to_nchw");
  return RCFD_OK;
}

int rcfd_depth_head_bwd(const float* ddepth, const float* depth, void* dlogit, float min_depth, float min_over_max,
                        int64_t count, int32_t cpad, int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(ddepth && depth && dlogit && count > 0 && min_depth > 0.f && cpad >= 1, "depth_head_bwd: bad args");
  DISPATCH_T(dtype, (depth_head_bwd_kernel<T><<<grid_for(count), NT, 0, (cudaStream_t)stream>>>(
                        ddepth, depth, (T*)dlogit, min_depth, min_over_max, count, cpad)));
  RCFD_CHECK_LAUNCH("depth_head_bwd");
  return RCFD_OK;
}

int rcfd_masked_l1_loss(const float* out, const float* gt, const float* lidar, float w_lidar, double* accum,
                        float* loss, float* dout, int64_t count, void* stream) {
  RCFD_CHECK_ARG(out && gt && lidar && accum && loss && count > 0, "masked_l1_loss: bad args");
  RCFD_CHECK_ARG(w_lidar > 0.f, "masked_l1_loss: the fused kernel implements the w_lidar_loss > 0 branch");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(accum, 0, 4 * sizeof(double), st);
  if (e != cudaSuccess) { set_error("masked_l1 memset: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
  l1_accum_kernel<<<grid_for(count, NT, 148 * 4), NT, 0, st>>>(out, gt, lidar, accum, count);
  RCFD_CHECK_LAUNCH("l1_accum");
  l1_finish_kernel<<<grid_for(count), NT, 0, st>>>(out, gt, lidar, accum, w_lidar, loss, dout, count);
  RCFD_CHECK_LAUNCH("l1_finish");
  return RCFD_OK;
}

int rcfd_outlier_removal(const float* depth, float* out, float* scratch_max, int32_t n, int32_t h, int32_t w,
                         int32_t kernel_size, float threshold, void* stream) {
  RCFD_CHECK_ARG(depth && out && scratch_max && n > 0 && h > 0 && w > 0 && kernel_size > 0 && (kernel_size & 1) &&
                     kernel_size <= 2 * OUT_MAXPAD + 1,
                 "outlier_removal: bad args (odd kernel size <= 15)");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)n * h * w;
  cudaError_t e = cudaMemsetAsync(scratch_max, 0, sizeof(float), st);
  if (e != cudaSuccess) { set_error("outlier memset: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
  max_kernel<<<grid_for(total, NT, 148 * 2), NT, 0, st>>>(depth, scratch_max, total);
  RCFD_CHECK_LAUNCH("outlier_max");
  outlier_kernel<<<dim3(ceil_div(w, OUT_TW), ceil_div(h, OUT_TH), n), OUT_TW * OUT_TH, 0, st>>>(depth, scratch_max, out, n, h, w, kernel_size,
                                                                                          threshold);
  RCFD_CHECK_LAUNCH("outlier");
  return RCFD_OK;
}

int rcfd_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t count, float lr,
                   float beta1, float beta2, float eps, int32_t step, void* stream) {
  RCFD_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && count > 0 && step > 0, "adam: bad args");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<grid_for(count), NT, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, count,
                                                                (float)((double)lr / bc1), beta1, beta2, eps,
                                                                (float)(1.0 / sqrt(bc2)));
  RCFD_CHECK_LAUNCH("adam");
  return RCFD_OK;
}

int rcfd_pack_conv_weight(const float* w_oihw, void* packed, int32_t cout, int32_t cin, int32_t kh, int32_t kw,
                          int32_t cin_off, int32_t cin_cnt, int32_t cin_pad, int32_t mode, int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(w_oihw && packed && cout > 0 && cin > 0 && kh > 0 && kw > 0, "pack_weight: bad args");
  RCFD_CHECK_ARG(cin_off >= 0 && cin_cnt > 0 && cin_off < cin && (mode == 1 || cin_off + cin_cnt <= cin) && (mode == 0 || mode == 1),
                 "pack_weight: range");
  RCFD_CHECK_ARG(mode == 1 || cin_pad >= cin_cnt, "pack_weight: cin_pad < cin_cnt");
  const int64_t total = mode == 0 ? (int64_t)cout * kh * kw * cin_pad
                                  : (int64_t)cin_cnt * kh * kw * (cin_pad > cout ? cin_pad : cout);
  DISPATCH_T(dtype, (pack_weight_kernel<T><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>(
                        w_oihw, (T*)packed, cout, cin, kh, kw, cin_off, cin_cnt, cin_pad, mode)));
  RCFD_CHECK_LAUNCH("pack_weight");
  return RCFD_OK;
}

int rcfd_pack_upconv2x_weight(const float* w_oihw, void* packed, int32_t cout, int32_t cin, int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(w_oihw && packed && cout > 0 && cin > 0, "pack_upconv2x: bad args");
  const int64_t total = (int64_t)16 * cout * cin;
  DISPATCH_T(dtype, (pack_up2x_kernel<T><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>(w_oihw, (T*)packed, cout, cin)));
  RCFD_CHECK_LAUNCH("pack_upconv2x");
  return RCFD_OK;
}

int rcfd_pack_dgrad_s2_weight(const float* w_oihw, void* packed, int32_t cout, int32_t cin, int32_t cin_off,
                              int32_t cin_cnt, int32_t cout_pad, int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(w_oihw && packed && cout > 0 && cin_cnt > 0 && cin_off >= 0 && cin_off + cin_cnt <= cin && cout_pad >= cout,
                 "pack_dgrad_s2_weight: bad args");
  const int64_t total = (int64_t)16 * cin_cnt * cout_pad;
  DISPATCH_T(dtype, (pack_dgrad_s2_kernel<T><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>(
                        w_oihw, (T*)packed, cout, cin, cin_off, cin_cnt, cout_pad)));
  RCFD_CHECK_LAUNCH("pack_dgrad_s2_weight");
  return RCFD_OK;
}

int rcfd_pack_upconv2x_dgrad_weight(const float* w_oihw, void* packed, int32_t cout, int32_t cin, int32_t cin_off,
                                    int32_t cin_cnt, int32_t cout_pad, int32_t dtype, void* stream) {
  RCFD_CHECK_ARG(w_oihw && packed && cout > 0 && cin_cnt > 0 && cin_off >= 0 && cin_off + cin_cnt <= cin && cout_pad >= cout,
                 "pack_upconv2x_dgrad_weight: bad args");
  const int64_t total = (int64_t)cin_cnt * 16 * cout_pad;
  DISPATCH_T(dtype, (pack_upconv_dgrad_kernel<T><<<grid_for(total), NT, 0, (cudaStream_t)stream>>>(
                        w_oihw, (T*)packed, cout, cin, cin_off, cin_cnt, cout_pad)));
  RCFD_CHECK_LAUNCH("pack_upconv2x_dgrad_weight");
  return RCFD_OK;
}

int rcfd_unpack_conv_wgrad(const float* packed, float* g_oihw, int32_t cout, int32_t cin, int32_t kh, int32_t kw,
                           int32_t cin_off, int32_t cin_cnt, int32_t cin_pad, int32_t accumulate, void* stream) {
  RCFD_CHECK_ARG(packed && g_oihw && cout > 0 && cin > 0 && kh > 0 && kw > 0, "unpack_wgrad: bad args");
  RCFD_CHECK_ARG(cin_off >= 0 && cin_cnt > 0 && cin_off + cin_cnt <= cin && cin_pad >= cin_cnt, "unpack_wgrad: range");
  const int64_t total = (int64_t)cout * kh * kw * cin_cnt;
  unpack_wgrad_kernel<<<grid_for(total), NT, 0, (cudaStream_t)stream>>>(packed, g_oihw, cout, cin, kh, kw, cin_off,
                                                                        cin_cnt, cin_pad, accumulate);
  RCFD_CHECK_LAUNCH("unpack_wgrad");
  return RCFD_OK;
}

}  // extern "C"
