// Shared device/host helpers for librcfd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/rcfd.h"

namespace rcfd {

void set_error(const char* fmt, ...);
void note_kernel(const char* fmt, ...);   // name of the kernel the last conv / wgrad call launched (rcfd_last_kernel)

#define RCFD_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      ::rcfd::set_error(__VA_ARGS__);             \
      return RCFD_EINVAL;                         \
    }                                             \
  } while (0)

#define RCFD_CHECK_LAUNCH(name)                                                   \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      ::rcfd::set_error("%s: %s", name, cudaGetErrorString(e__));                 \
      return RCFD_ECUDA;                                                          \
    }                                                                             \
  } while (0)

constexpr float kLeakySlope = 0.2f;   // reference: src/net_utils.py:15

typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4-wide vector load/store of T as floats (pointer must be 4-element aligned)
template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ float4 ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void st(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Vec4<bf16> {
  static __device__ __forceinline__ float4 ld(const bf16* p) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ void st(bf16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
  }
};

// 16-byte vector of T viewed as floats: 4 x float or 8 x bf16 (pointer must be 16-byte aligned)
template <typename T> struct V16;
template <> struct V16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void ld(const float* p, float* f) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  static __device__ __forceinline__ void st(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct V16<bf16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void ld(const bf16* p, float* f) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  static __device__ __forceinline__ void st(bf16* p, const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float sigmoid_precise(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : kLeakySlope * x; }

__device__ __forceinline__ float apply_act(float v, int act, float p0, float p1) {
  switch (act) {
    case RCFD_ACT_LEAKY: return leaky(v);
    case RCFD_ACT_SIGMOID: return sigmoid_precise(v);
    case RCFD_ACT_DEPTH_HEAD: return p0 / (sigmoid_precise(v) + p1);
    default: return v;
  }
}

// ATen nearest-neighbour source index (UpSampleNearest: floorf(dst * scale), clamped)
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
  int s = (int)floorf((float)dst * scale);
  return s < in_size - 1 ? s : in_size - 1;
}

// Phase weight of the stride-2 / 3x3 / pad-1 data gradient: destination pixel (2i + a, 2j + b) = sum over the 2x2 taps
// (t, u) on dy rows i-1+a+t, columns j-1+b+u of W'[a,b][t,u] = w[r(a,t)][s(b,u)]:  a = 0: (t=0: none, t=1: r=1);
// a = 1: (t=0: r=2, t=1: r=0); same for columns.  w is OIHW [cout][cin][3][3]; (ci, co) = (row, column) of the dgrad GEMM.
__host__ __device__ inline float dgrad_s2_weight(const float* w, int cout, int cin, int ci, int co, int ph, int tap) {
  const int a = ph >> 1, b = ph & 1, t = tap >> 1, u = tap & 1;
  const int r = a == 0 ? (t == 0 ? -1 : 1) : (t == 0 ? 2 : 0);
  const int s = b == 0 ? (u == 0 ? -1 : 1) : (u == 0 ? 2 : 0);
  if (r < 0 || s < 0 || co >= cout) return 0.f;
  return w[((size_t)co * cin + ci) * 9 + r * 3 + s];
}

// Data gradient of `3x3 / pad-1 conv after exact 2x nearest up-sampling` w.r.t. the LOW-RES source, as one 4x4 / stride-2 /
// pad-1 convolution over dy:  dsrc[i][j] = sum_{k,l in 0..3} W4[k][l] . dy[2i-1+k][2j-1+l],  W4[k][l] = sum of w[r][s] over the
// taps with k + r in {2, 3} and l + s in {2, 3} (the up-sampled pixels 2i, 2i+1 that source pixel i feeds).
__host__ __device__ inline float upconv_dgrad_weight(const float* w, int cout, int cin, int ci, int co, int tap) {
  if (co >= cout) return 0.f;
  const int k = tap >> 2, l = tap & 3;
  const int r0 = k == 0 ? 2 : (k == 1 ? 1 : 0), r1 = k == 0 ? 2 : (k == 1 ? 2 : (k == 2 ? 1 : 0));
  const int s0 = l == 0 ? 2 : (l == 1 ? 1 : 0), s1 = l == 0 ? 2 : (l == 1 ? 2 : (l == 2 ? 1 : 0));
  const float* wp = w + ((size_t)co * cin + ci) * 9;
  float acc = 0.f;
  for (int r = r0; r <= r1; ++r)
    for (int s = s0; s <= s1; ++s) acc += wp[r * 3 + s];
  return acc;
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// index of the current device (function attributes such as the dynamic shared-memory limit are per device)
inline int cur_dev() {
  int d = 0;
  cudaGetDevice(&d);
  return d & 15;
}

}  // namespace rcfd
