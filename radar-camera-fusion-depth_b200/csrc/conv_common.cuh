// Implicit-GEMM convolution: shared geometry + gather addressing used by both engines
// (SIMT fp32-FMA in conv_simt.cu, tcgen05 in conv_tc.cu).
//
// GEMM view:  M = n*ho*wo output pixels,  N = cout,  K = kh*kw*(c0+c1) ordered (tap, channel),
// channels ordered [src0 | src1] exactly like torch.cat([upsampled, skip], 1)
// (reference: src/net_utils.py:565).  Nearest up-sampling of src0 (src/net_utils.py:196)
// and zero-insertion (dgrad of stride-2 convs) are address maps, never materialised.
#pragma once
#include "common.cuh"

namespace rcfd {

struct ConvKP {
  int n, ho, wo, cout;
  int kh, kw, stride, pad, dil;
  int hin, win;
  const void* src0; int h0, w0, c0; int up; float sch, scw;
  const void* src1; int c1;
  int ctot, K, M;
  const void* weight;
  const void* weight_up2x;   // optional sub-pixel phase weights (rcfd_pack_upconv2x_weight)
  void* dst;
  const float* scale; const float* shift;
  int act; float p0, p1;
  const void* residual;
  double* ssum; double* ssq;
  int accumulate, dst_f32;
};

int make_conv_kp(const rcfd_conv_desc* d, ConvKP* p);   // validates; defined in conv_simt.cu

// Address of the element feeding output pixel (n, oy, ox), filter tap `tap`, concatenated
// channel `c`; returns false when the tap falls in the zero padding / an inserted zero.
template <typename T>
__device__ __forceinline__ bool conv_src_ptr(const ConvKP& p, int n, int oy, int ox, int tap, int c,
                                             const T*& ptr) {
  int r = tap / p.kw, s = tap - r * p.kw;
  int iy = oy * p.stride - p.pad + r;
  int ix = ox * p.stride - p.pad + s;
  if (p.dil == 2) {
    if ((iy | ix) & 1) return false;
    iy >>= 1;
    ix >>= 1;
  }
  if (iy < 0 || iy >= p.hin || ix < 0 || ix >= p.win) return false;
  if (c < p.c0) {
    int sy = iy, sx = ix;
    if (p.up) {
      sy = nearest_src(iy, p.sch, p.h0);
      sx = nearest_src(ix, p.scw, p.w0);
    }
    ptr = reinterpret_cast<const T*>(p.src0) + ((size_t)(n * p.h0 + sy) * p.w0 + sx) * p.c0 + c;
  } else {
    ptr = reinterpret_cast<const T*>(p.src1) + ((size_t)(n * p.hin + iy) * p.win + ix) * p.c1 + (c - p.c0);
  }
  return true;
}

}  // namespace rcfd
