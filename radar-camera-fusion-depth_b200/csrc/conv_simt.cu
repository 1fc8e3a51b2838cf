// SIMT engine of the implicit-GEMM convolution: fp32 FMA, fp32 accumulate.
// This is the full-precision ("parity", RCFD_F32) product path and the engine for shapes
// the tcgen05 kernel does not take (3- and 2-channel 7x7 stems, cout = 1 head).
// Replaces cuDNN conv fwd / dgrad / wgrad reached from src/net_utils.py:85.
#include "conv_common.cuh"

namespace rcfd {

int make_conv_kp(const rcfd_conv_desc* d, ConvKP* p) {
  RCFD_CHECK_ARG(d != nullptr, "conv: null descriptor");
  RCFD_CHECK_ARG(d->n > 0 && d->ho > 0 && d->wo > 0 && d->cout > 0, "conv: bad output shape");
  RCFD_CHECK_ARG(d->kh > 0 && d->kw > 0 && d->stride > 0 && d->pad >= 0, "conv: bad filter");
  RCFD_CHECK_ARG(d->in_dilation == 1 || (d->in_dilation == 2 && d->stride == 1),
                 "conv: in_dilation must be 1, or 2 with stride 1");
  RCFD_CHECK_ARG(d->src0 && d->c0 > 0 && d->h0 > 0 && d->w0 > 0, "conv: bad src0");
  RCFD_CHECK_ARG(d->c1 >= 0 && (d->c1 == 0 || d->src1), "conv: bad src1");
  RCFD_CHECK_ARG(d->hin > 0 && d->win > 0, "conv: bad logical input extent");
  RCFD_CHECK_ARG(d->weight && d->dst, "conv: null weight/dst");
  RCFD_CHECK_ARG(d->dtype == RCFD_F32 || d->dtype == RCFD_BF16, "conv: bad dtype");
  RCFD_CHECK_ARG((d->scale == nullptr) == (d->shift == nullptr), "conv: scale/shift must come together");
  RCFD_CHECK_ARG((d->stats_sum == nullptr) == (d->stats_sqsum == nullptr), "conv: stats pointers must come together");
  RCFD_CHECK_ARG((int64_t)d->n * d->ho * d->wo < (int64_t)1 << 31, "conv: too many output pixels");
  p->n = d->n; p->ho = d->ho; p->wo = d->wo; p->cout = d->cout;
  p->kh = d->kh; p->kw = d->kw; p->stride = d->stride; p->pad = d->pad; p->dil = d->in_dilation;
  p->hin = d->hin; p->win = d->win;
  p->src0 = d->src0; p->h0 = d->h0; p->w0 = d->w0; p->c0 = d->c0;
  p->up = (d->h0 != d->hin || d->w0 != d->win) ? 1 : 0;
  p->sch = (float)d->h0 / (float)d->hin;   // ATen compute_scales_value<float>(in, out)
  p->scw = (float)d->w0 / (float)d->win;
  p->src1 = d->src1; p->c1 = d->c1;
  p->ctot = d->c0 + d->c1;
  p->K = d->kh * d->kw * p->ctot;
  p->M = d->n * d->ho * d->wo;
  p->weight = d->weight; p->weight_up2x = d->weight_up2x; p->dst = d->dst;
  p->scale = d->scale; p->shift = d->shift;
  p->act = d->act; p->p0 = d->act_p0; p->p1 = d->act_p1;
  p->residual = d->residual;
  p->ssum = d->stats_sum; p->ssq = d->stats_sqsum;
  p->accumulate = d->accumulate; p->dst_f32 = d->dst_f32;
  return RCFD_OK;
}

namespace {

constexpr int BM = 64, BN = 64, BK = 32, NT = 256, PADS = 4;

// ------------------------------------------------------------------ forward / dgrad
template <typename T, int VEC>
__global__ void __launch_bounds__(NT) conv_simt_kernel(const ConvKP p) {
  __shared__ __align__(16) float As[BK][BM + PADS];
  __shared__ __align__(16) float Bs[BK][BN + PADS];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;

  // this thread's A-gather row
  const int lm = tid & 63;
  const int gm = m0 + lm;
  const bool mvalid = gm < p.M;
  int pn = 0, oy = 0, ox = 0;
  if (mvalid) {
    pn = gm / (p.ho * p.wo);
    int rem = gm - pn * p.ho * p.wo;
    oy = rem / p.wo;
    ox = rem - oy * p.wo;
  }
  const T* W = reinterpret_cast<const T*>(p.weight);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    if (VEC == 4) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int kv = (tid >> 6) + 4 * i;
        const int k = k0 + kv * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mvalid && k < p.K) {
          const int tap = k / p.ctot, c = k - tap * p.ctot;
          const T* ptr;
          if (conv_src_ptr<T>(p, pn, oy, ox, tap, c, ptr)) a = Vec4<T>::ld(ptr);
        }
        As[kv * 4 + 0][lm] = a.x; As[kv * 4 + 1][lm] = a.y; As[kv * 4 + 2][lm] = a.z; As[kv * 4 + 3][lm] = a.w;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        const int gn = n0 + lm;
        if (gn < p.cout && k < p.K) b = Vec4<T>::ld(W + (size_t)gn * p.K + k);
        Bs[kv * 4 + 0][lm] = b.x; Bs[kv * 4 + 1][lm] = b.y; Bs[kv * 4 + 2][lm] = b.z; Bs[kv * 4 + 3][lm] = b.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int kk = (tid >> 6) + 4 * i;
        const int k = k0 + kk;
        float a = 0.f, b = 0.f;
        if (mvalid && k < p.K) {
          const int tap = k / p.ctot, c = k - tap * p.ctot;
          const T* ptr;
          if (conv_src_ptr<T>(p, pn, oy, ox, tap, c, ptr)) a = to_f<T>(*ptr);
        }
        const int gn = n0 + lm;
        if (gn < p.cout && k < p.K) b = to_f<T>(W[(size_t)gn * p.K + k]);
        As[kk][lm] = a;
        Bs[kk][lm] = b;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- BatchNorm statistics of the raw accumulators (training mode)
  if (p.ssum != nullptr) {
    float* red_s = &As[0][0];   // [16][64]
    float* red_q = &Bs[0][0];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (m0 + ty * 4 + i < p.M) {
          s += acc[i][j];
          q += acc[i][j] * acc[i][j];
        }
      }
      red_s[ty * 64 + tx * 4 + j] = s;
      red_q[ty * 64 + tx * 4 + j] = q;
    }
    __syncthreads();
    if (tid < 64 && n0 + tid < p.cout) {
      double s = 0.0, q = 0.0;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        s += (double)red_s[r * 64 + tid];
        q += (double)red_q[r * 64 + tid];
      }
      atomicAdd(p.ssum + n0 + tid, s);
      atomicAdd(p.ssq + n0 + tid, q);
    }
  }

  // ---- epilogue: affine (folded BN) -> activation -> residual add + leaky -> store
  const T* R = reinterpret_cast<const T*>(p.residual);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.cout) continue;
      float v = acc[i][j];
      if (p.scale) v = fmaf(v, p.scale[n], p.shift[n]);
      v = apply_act(v, p.act, p.p0, p.p1);
      const size_t o = (size_t)m * p.cout + n;
      if (R) v = leaky(v + to_f<T>(R[o]));
      if (p.dst_f32) {
        float* D = reinterpret_cast<float*>(p.dst);
        D[o] = p.accumulate ? D[o] + v : v;
      } else {
        T* D = reinterpret_cast<T*>(p.dst);
        D[o] = from_f<T>(p.accumulate ? to_f<T>(D[o]) + v : v);
      }
    }
  }
}

// ------------------------------------------------------------------ wgrad
// dw[co][k] += sum_{m in this split} dy[m][co] * A[m][k]     (dw is zeroed by the host call)
constexpr int WP = 32;   // pixels per reduction step
template <typename T, int VEC>
__global__ void __launch_bounds__(NT) conv_wgrad_simt_kernel(const ConvKP p, float* __restrict__ dw,
                                                             int pixels_per_split) {
  __shared__ __align__(16) float Ds[WP][BM + PADS];   // dy  [pixel][cout]
  __shared__ __align__(16) float As[WP][BN + PADS];   // act [pixel][k]
  const int tid = threadIdx.x;
  const int kt0 = blockIdx.x * BN, co0 = blockIdx.y * BM;
  const int mbeg = blockIdx.z * pixels_per_split;
  const int mend = min(p.M, mbeg + pixels_per_split);
  const int tx = tid & 15, ty = tid >> 4;
  const T* DY = reinterpret_cast<const T*>(p.dst);

  // fixed K coordinates of this thread's gather column(s)
  const int kv = tid & 15;                 // VEC==4: vector column; VEC==1 uses (tid & 63)
  int tap4 = 0, c4 = 0;
  bool kvalid4 = false;
  if (VEC == 4) {
    const int k = kt0 + kv * 4;
    kvalid4 = k < p.K;
    if (kvalid4) { tap4 = k / p.ctot; c4 = k - tap4 * p.ctot; }
  }
  int tap1 = 0, c1 = 0;
  bool kvalid1 = false;
  if (VEC == 1) {
    const int k = kt0 + (tid & 63);
    kvalid1 = k < p.K;
    if (kvalid1) { tap1 = k / p.ctot; c1 = k - tap1 * p.ctot; }
  }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int mb = mbeg; mb < mend; mb += WP) {
    // dy tile: 32 pixels x 64 couts
    if (p.cout % 4 == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int pix = (tid >> 4) + 16 * i, cv = tid & 15;
        const int m = mb + pix, co = co0 + cv * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < mend && co < p.cout) v = Vec4<T>::ld(DY + (size_t)m * p.cout + co);
        *reinterpret_cast<float4*>(&Ds[pix][cv * 4]) = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int pix = (tid >> 6) + 4 * i, cl = tid & 63;
        const int m = mb + pix, co = co0 + cl;
        Ds[pix][cl] = (m < mend && co < p.cout) ? to_f<T>(DY[(size_t)m * p.cout + co]) : 0.f;
      }
    }
    // gathered activation tile: 32 pixels x 64 k
    if (VEC == 4) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int pix = (tid >> 4) + 16 * i;
        const int m = mb + pix;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < mend && kvalid4) {
          const int pn = m / (p.ho * p.wo);
          const int rem = m - pn * p.ho * p.wo;
          const int oy = rem / p.wo, ox = rem - oy * p.wo;
          const T* ptr;
          if (conv_src_ptr<T>(p, pn, oy, ox, tap4, c4, ptr)) a = Vec4<T>::ld(ptr);
        }
        *reinterpret_cast<float4*>(&As[pix][kv * 4]) = a;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int pix = (tid >> 6) + 4 * i;
        const int m = mb + pix;
        float a = 0.f;
        if (m < mend && kvalid1) {
          const int pn = m / (p.ho * p.wo);
          const int rem = m - pn * p.ho * p.wo;
          const int oy = rem / p.wo, ox = rem - oy * p.wo;
          const T* ptr;
          if (conv_src_ptr<T>(p, pn, oy, ox, tap1, c1, ptr)) a = to_f<T>(*ptr);
        }
        As[pix][tid & 63] = a;
      }
    }
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < WP; ++pp) {
      const float4 a = *reinterpret_cast<const float4*>(&Ds[pp][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&As[pp][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= p.cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kt0 + tx * 4 + j;
      if (k < p.K) atomicAdd(dw + (size_t)co * p.K + k, acc[i][j]);
    }
  }
}

}  // namespace

int conv_simt_launch(const ConvKP& p, int dtype, cudaStream_t st) {
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.cout, BN));
  const bool vec = (p.c0 % 4 == 0) && (p.c1 % 4 == 0);
  if (dtype == RCFD_F32) {
    if (vec) conv_simt_kernel<float, 4><<<grid, NT, 0, st>>>(p);
    else conv_simt_kernel<float, 1><<<grid, NT, 0, st>>>(p);
  } else {
    if (vec) conv_simt_kernel<bf16, 4><<<grid, NT, 0, st>>>(p);
    else conv_simt_kernel<bf16, 1><<<grid, NT, 0, st>>>(p);
  }
  RCFD_CHECK_LAUNCH("conv_simt");
  return RCFD_OK;
}

int conv_wgrad_simt_launch(const ConvKP& p, float* dw, int dtype, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)p.cout * p.K * sizeof(float), st);
  if (e != cudaSuccess) { set_error("wgrad memset: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
  const int gx = ceil_div(p.K, BN), gy = ceil_div(p.cout, BM);
  int splits = (148 * 4 + gx * gy - 1) / (gx * gy);
  const int max_splits = (p.M + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int pps = (p.M + splits - 1) / splits;
  pps = ((pps + WP - 1) / WP) * WP;
  splits = (p.M + pps - 1) / pps;
  dim3 grid(gx, gy, splits);
  const bool vec = (p.c0 % 4 == 0) && (p.c1 % 4 == 0);
  if (dtype == RCFD_F32) {
    if (vec) conv_wgrad_simt_kernel<float, 4><<<grid, NT, 0, st>>>(p, dw, pps);
    else conv_wgrad_simt_kernel<float, 1><<<grid, NT, 0, st>>>(p, dw, pps);
  } else {
    if (vec) conv_wgrad_simt_kernel<bf16, 4><<<grid, NT, 0, st>>>(p, dw, pps);
    else conv_wgrad_simt_kernel<bf16, 1><<<grid, NT, 0, st>>>(p, dw, pps);
  }
  RCFD_CHECK_LAUNCH("conv_wgrad_simt");
  return RCFD_OK;
}

}  // namespace rcfd
