// tcgen05 weight gradient of the implicit-GEMM convolution (bf16 operands, fp32 accumulate):
//
//     dW[k][co] = sum over output pixels m of  A[m][k] * dY[m][co]          k = (tap, channel)
//
// i.e. a GEMM whose REDUCTION dimension is the pixel index.  Both operands are read exactly as
// they lie in HBM (NHWC: channels contiguous), which makes them MN-major for the tensor core:
//   A^T tile: 64 pixels x 128 k-values, dY tile: 64 pixels x BN channels, both stored in shared
//   memory as [pixel row][64 elements = 128 B] blocks with the 128-byte swizzle; the UMMA
//   descriptors walk 8-row groups with SBO = 1024 B and 64-element blocks with LBO = 8192 B.
// One CTA owns a (128 k-values) x (BN channels) accumulator in TMEM and a contiguous range of
// pixels (split-K over pixels across blockIdx.z); partial sums are reduced with coalesced
// fp32 red.global.add into the packed gradient.  The gather (taps, padding, stride, nearest
// up-sampling, concat) is the same address arithmetic as the forward engine.
#include "tc_common.cuh"

namespace rcfd {
namespace {

using namespace tc;
constexpr int PB = 64;           // pixels per stage (reduction depth: 4 x UMMA_K)
constexpr int NPROD = 128;
constexpr int NTHREADS = 160;
constexpr int BLK = 64 * 128;    // one [64 pixel][64 element] block

template <int BN>
struct WgCfg {
  static constexpr int STAGES = BN >= 256 ? 3 : 4;
  static constexpr int A_BYTES = 2 * BLK;
  static constexpr int B_BYTES = (BN / 64) * BLK;
  static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS) wgrad_tc_kernel(const ConvKP p, float* __restrict__ dw, int pixels_per_split) {
  typedef WgCfg<BN> C;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + C::STAGES * C::A_BYTES;
  const uint32_t sBar = sB + C::STAGES * C::B_BYTES;
  const uint32_t sTmem = sBar + 8 * (2 * C::STAGES + 1);
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (base - smem_u32(smem_raw)) + (sTmem - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k0 = blockIdx.x * TM, n0 = blockIdx.y * BN;
  const int mbeg = blockIdx.z * pixels_per_split;
  const int mend = min(p.M, mbeg + pixels_per_split);
  const int num_pb = (mend - mbeg + PB - 1) / PB;

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(sBar + 8 * s, NPROD);
      mbar_init(sBar + 8 * (C::STAGES + s), 1);
    }
    mbar_init(sBar + 8 * (2 * C::STAGES), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmem), "n"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // =========================================================== PRODUCER
    // Thread t serves pixel row (t >> 1) of the 64-pixel stage and one 64-element k block
    // (t & 1): 8 consecutive 16-byte chunks; the (pixel, tap) -> address computation is done once
    // per tap the 64 k-values touch (one when C >= 64).
    const int half = tid & 1;
    const int prow = tid >> 1;
    const uint32_t swz = (uint32_t)(prow & 7);
    const bf16* S0 = reinterpret_cast<const bf16*>(p.src0);
    const bf16* S1 = reinterpret_cast<const bf16*>(p.src1);
    const bf16* DY = reinterpret_cast<const bf16*>(p.dst);
    // fixed (tap row, tap col, channel) of this thread's first chunk: k = k0 + half * 64
    const int kmine = k0 + half * 64;
    const int tap0 = kmine / p.ctot;
    const int tc0 = kmine - tap0 * p.ctot;
    const int tr0 = tap0 / p.kw, ts0 = tap0 - tr0 * p.kw;
    // running pixel coordinates of this thread's row (advance by 64 pixels per stage)
    int pn, oy, ox;
    {
      const int m = mbeg + prow;
      pn = m / (p.ho * p.wo);
      const int rem = m - pn * p.ho * p.wo;
      oy = rem / p.wo;
      ox = rem - oy * p.wo;
    }
    constexpr int CH = BN / 8;     // 16-byte chunks per dY pixel row
    for (int pb = 0; pb < num_pb; ++pb) {
      const int s = pb % C::STAGES;
      if (pb >= C::STAGES) mbar_wait(sBar + 8 * (C::STAGES + s), ((pb / C::STAGES) & 1) ^ 1);
      const int mb = mbeg + pb * PB;
      const bool mvalid = (mb + prow) < mend;
      const int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
      const uint32_t a_st = sA + s * C::A_BYTES + (uint32_t)half * BLK + (uint32_t)prow * 128u;
      int cc = tc0, cs = ts0, cr = tr0;
      int prev_r = -1, prev_s = -1, prev_src = -1;
      const bf16* base = S0;
      bool ok = false;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int from1 = cc >= p.c0 ? 1 : 0;
        if (cr != prev_r || cs != prev_s || from1 != prev_src) {
          prev_r = cr; prev_s = cs; prev_src = from1;
          const int iy = iy0 + cr, ix = ix0 + cs;
          ok = mvalid && cr < p.kh && iy >= 0 && iy < p.hin && ix >= 0 && ix < p.win;
          base = S0;
          if (ok) {
            if (!from1) {
              int sy = iy, sx = ix;
              if (p.up) {
                sy = nearest_src(iy, p.sch, p.h0);
                sx = nearest_src(ix, p.scw, p.w0);
              }
              base = S0 + ((size_t)(pn * p.h0 + sy) * p.w0 + sx) * p.c0;
            } else {
              base = S1 + ((size_t)(pn * p.hin + iy) * p.win + ix) * p.c1 - p.c0;
            }
          }
        }
        cp_async16(a_st + (((uint32_t)j ^ swz) << 4), ok ? base + cc : S0, ok ? 16u : 0u);
        cc += 8;
        if (cc >= p.ctot) {
          cc = 0;
          if (++cs == p.kw) { cs = 0; ++cr; }
        }
      }
      // advance this row by one stage (64 pixels)
      ox += PB;
      while (ox >= p.wo) {
        ox -= p.wo;
        if (++oy == p.ho) { oy = 0; ++pn; }
      }
      const uint32_t b_st = sB + s * C::B_BYTES;
      for (int i = tid; i < PB * CH; i += NPROD) {
        const int pix = i / CH, ch = i - pix * CH;
        const int blk = ch >> 3, jj = ch & 7;
        const int mm = mb + pix, co = n0 + ch * 8;
        const bool ok = (mm < mend) && (co < p.cout);
        const bf16* src = ok ? DY + (size_t)mm * p.cout + co : DY;
        cp_async16(b_st + (uint32_t)blk * BLK + (uint32_t)pix * 128u + (((uint32_t)jj ^ (uint32_t)(pix & 7)) << 4), src,
                   ok ? 16u : 0u);
      }
      // publish the stage: the barrier fires when this thread's copies have landed (no thread-side
      // wait, so up to STAGES stages of loads stay in flight); the MMA warp does the proxy fence.
      cp_async_mbar_arrive_noinc(sBar + 8 * s);
    }

    // =========================================================== EPILOGUE: TMEM -> red.global.add
    if (num_pb > 0) {
      mbar_wait(sBar + 8 * (2 * C::STAGES), 0);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
      const int k = k0 + tid;
#pragma unroll 1
      for (int cb = 0; cb < BN; cb += 16) {
        float v[16];
        tmem_ld16(trow + cb, v);
        if (k < p.K) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int co = n0 + cb + i;
            if (co < p.cout) atomicAdd(dw + (size_t)co * p.K + k, v[i]);
          }
        }
      }
    }
    tc_fence_before();
  } else {
    // =========================================================== MMA ISSUER
    const uint32_t idesc = umma_idesc_ex(TM, BN, /*a MN-major*/ 1, /*b MN-major*/ 1);
    for (int pb = 0; pb < num_pb; ++pb) {
      const int s = pb % C::STAGES;
      mbar_wait(sBar + 8 * s, (pb / C::STAGES) & 1);
      fence_proxy_async();      // generic-proxy (cp.async) writes -> async-proxy (UMMA) reads
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_st = sA + s * C::A_BYTES, b_st = sB + s * C::B_BYTES;
#pragma unroll
        for (int kk = 0; kk < PB / 16; ++kk) {
          umma_f16(tmem_base, umma_desc(a_st + kk * 2048, BLK, 1024, 2), umma_desc(b_st + kk * 2048, BLK, 1024, 2), idesc,
                   (uint32_t)((pb | kk) != 0));
        }
        umma_commit(sBar + 8 * (C::STAGES + s));
        if (pb == num_pb - 1) umma_commit(sBar + 8 * (2 * C::STAGES));
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
  }
}

template <int BN>
int launch_wg(const ConvKP& p, float* dw, cudaStream_t st) {
  note_kernel("wgrad_tc_kernel<%d>", BN);
  typedef WgCfg<BN> C;
  static bool attr_set_dev[16] = {}; bool& attr_set = attr_set_dev[cur_dev()];   // per device: the attribute belongs to the device's copy of the kernel
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) { set_error("wgrad_tc: smem attribute: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
    attr_set = true;
  }
  const int gx = ceil_div(p.K, TM), gy = ceil_div(p.cout, BN);
  // one wave of resident CTAs: two per SM fit only for the 64-wide tile (96 KB of stages), the wider ones take an SM each
  const int capacity = 148 * (C::SMEM <= 113 * 1024 ? 2 : 1);
  int splits = capacity / (gx * gy);
  const int max_splits = (p.M + 4 * PB - 1) / (4 * PB);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int pps = (p.M + splits - 1) / splits;
  pps = ((pps + PB - 1) / PB) * PB;
  splits = (p.M + pps - 1) / pps;
  dim3 grid(gx, gy, splits);
  wgrad_tc_kernel<BN><<<grid, NTHREADS, C::SMEM, st>>>(p, dw, pps);
  RCFD_CHECK_LAUNCH("wgrad_tc");
  return RCFD_OK;
}

}  // namespace

bool wgrad_tc_supported(const ConvKP& p, int dtype) {
  return dtype == RCFD_BF16 && p.c0 % 8 == 0 && p.c1 % 8 == 0 && p.cout % 8 == 0 && p.dil == 1;
}

int wgrad_tc_launch(const ConvKP& p, float* dw, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)p.cout * p.K * sizeof(float), st);
  if (e != cudaSuccess) { set_error("wgrad_tc memset: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
  if (p.cout > 128) return launch_wg<256>(p, dw, st);
  if (p.cout > 64) return launch_wg<128>(p, dw, st);
  return launch_wg<64>(p, dw, st);
}

}  // namespace rcfd
