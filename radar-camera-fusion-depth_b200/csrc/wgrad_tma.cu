// TMA + tcgen05 weight gradient (bf16 operands, fp32 accumulation in TMEM):
//
//     dW[k][co] = sum over output pixels m of  A[m][k] * dY[m][co],      k = (tap, channel)
//
// for every convolution whose im2col rows are boxes of the NHWC input (stride 1 / 2, no
// up-sampling): each 64-pixel reduction step is fed by 4-D TMA tile loads -- (128 / bkc) boxes
// of the shifted input (one per (tap, channel chunk) of this CTA's 128 k-values) and
// (BN / bnb) boxes of dY -- that land in shared memory as [pixel][channel] blocks, which is
// exactly the MN-major operand layout tcgen05.mma takes (both operands are read the way
// they lie in HBM; nothing is transposed or materialised).  Out-of-image pixels of partial
// tiles are zero-filled by the TMA unit in dY, so they contribute nothing.
// grid = (k tiles of 128, cout tiles of BN, pixel splits); every CTA owns one TMEM
// accumulator and walks its share of the 64-pixel tiles; partial sums are merged with
// coalesced fp32 red.global.add.
#include "tma_common.cuh"

namespace rcfd {
namespace {

using namespace tc;
using namespace tma;
constexpr int PB = 64;            // pixels per reduction step
constexpr int NTHREADS = 192;     // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue

#ifdef RCFD_TRACE
__device__ long long g_wtrace[148 * 32];
#define TRACE(slot) do { const int b_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z); if (b_ < 148) g_wtrace[b_ * 32 + (slot)] = clock64(); } while (0)
#define TRACE_VAL(slot, v) do { const int b_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z); if (b_ < 148) g_wtrace[b_ * 32 + (slot)] = (long long)(v); } while (0)
#else
#define TRACE(slot) do { } while (0)
#define TRACE_VAL(slot, v) do { } while (0)
#endif

struct WgTmaP {
  int n, ho, wo, cout;
  int kh, kw, stride, pad;
  int c0, c1, bkc;                // channels per TMA box of the input (16 / 32 / 64)
  int bnb;                        // channels per TMA box of dY (16 / 32 / 64)
  int tw, th;                     // pixel tile (tw * th = 64)
  int tiles_x, tiles_y, num_ptiles;
  int K;
};

template <int BN>
struct WgTmaCfg {
  static constexpr int A_BYTES = PB * 128 * 2;                 // 64 pixels x 128 k
  static constexpr int B_BYTES = PB * BN * 2;
  static constexpr int STAGE = A_BYTES + (B_BYTES < 1024 ? 1024 : B_BYTES);
  static constexpr int STAGES = BN >= 256 ? 4 : 6;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int SMEM = STAGES * STAGE + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
wgrad_tma_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                 const __grid_constant__ CUtensorMap map_dy, const WgTmaP p, float* __restrict__ dw) {
  typedef WgTmaCfg<BN> C;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sStage = base;
  const uint32_t sBar = base + C::STAGES * C::STAGE;          // full[S], empty[S], accum
  const uint32_t sTmem = sBar + 8 * (2 * C::STAGES + 1);
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (base - smem_u32(smem_raw)) + (sTmem - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k0 = blockIdx.x * TM, n0 = blockIdx.y * BN;
  const int split = blockIdx.z, nsplit = gridDim.z;
  const int my_tiles = (p.num_ptiles - split + nsplit - 1) / nsplit;    // tiles split, split+nsplit, ...
  if (tid == 0) { TRACE(0); TRACE_VAL(12, my_tiles); }

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(sBar + 8 * s, 1);
      mbar_init(sBar + 8 * (C::STAGES + s), 1);
    }
    mbar_init(sBar + 8 * (2 * C::STAGES), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmem), "n"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) TRACE(1);

  const int ctot = p.c0 + p.c1;
  const int ablocks = 128 / p.bkc;                  // input boxes per stage
  const int a_blk_bytes = PB * p.bkc * 2;
  const int nblocks = BN / p.bnb;                   // dY boxes per stage
  const int b_blk_bytes = PB * p.bnb * 2;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a0) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
      if (p.c1 > 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a1) : "memory");
      // how many of this CTA's input boxes / dY boxes are inside K / cout (fixed for the kernel)
      int a_live = 0;
      for (int b = 0; b < ablocks; ++b) a_live += (k0 + b * p.bkc) < p.K ? 1 : 0;
      int b_live = 0;
      for (int b = 0; b < nblocks; ++b) b_live += (n0 + b * p.bnb) < p.cout ? 1 : 0;
      const uint32_t tx_bytes = (uint32_t)(a_live * a_blk_bytes + b_live * b_blk_bytes);
      TRACE(2);
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % C::STAGES;
        if (i >= C::STAGES) mbar_wait(sBar + 8 * (C::STAGES + s), ((i / C::STAGES) & 1) ^ 1);
        int sp = split + i * nsplit;
        const int tx = sp % p.tiles_x; sp /= p.tiles_x;
        const int ty = sp % p.tiles_y;
        const int img = sp / p.tiles_y;
        const int ox0 = tx * p.tw, oy0 = ty * p.th;
        const uint32_t full = sBar + 8 * s;
        const uint32_t a_dst = sStage + s * C::STAGE, b_dst = a_dst + C::A_BYTES;
        mbar_expect_tx(full, tx_bytes);
        for (int b = 0; b < ablocks; ++b) {
          const int k = k0 + b * p.bkc;
          if (k >= p.K) break;
          const int tap = k / ctot, c = k - tap * ctot;
          const int tr = tap / p.kw, ts = tap - tr * p.kw;
          const int ix = ox0 * p.stride - p.pad + ts, iy = oy0 * p.stride - p.pad + tr;
          if (c < p.c0) tma_load_4d(a_dst + b * a_blk_bytes, &map_a0, full, c, ix, iy, img);
          else tma_load_4d(a_dst + b * a_blk_bytes, &map_a1, full, c - p.c0, ix, iy, img);
        }
        for (int b = 0; b < nblocks; ++b) {
          const int co = n0 + b * p.bnb;
          if (co >= p.cout) break;
          tma_load_4d(b_dst + b * b_blk_bytes, &map_dy, full, co, ox0, oy0, img);
        }
      }
      TRACE(3);
    }
  } else if (warp == 1) {
    // both operands MN-major; swizzle span = box row bytes
    const uint32_t idesc = umma_idesc_ex(TM, BN, 1, 1);
    const uint32_t la = p.bkc == 64 ? 2u : (p.bkc == 32 ? 4u : 6u), sbo_a = (uint32_t)(8 * p.bkc * 2);
    const uint32_t lb = p.bnb == 64 ? 2u : (p.bnb == 32 ? 4u : 6u), sbo_b = (uint32_t)(8 * p.bnb * 2);
    for (int i = 0; i < my_tiles; ++i) {
      const int s = i % C::STAGES;
      mbar_wait(sBar + 8 * s, (i / C::STAGES) & 1);
      tc_fence_after();
      if (lane == 0 && i == 0) TRACE(4);
      if (lane == 0 && i == 1) TRACE(14);
      if (lane == 0 && i == 2) TRACE(15);
      if (lane == 0 && i == my_tiles - 1) TRACE(6);
      if (lane == 0) {
        const uint32_t a_st = sStage + s * C::STAGE, b_st = a_st + C::A_BYTES;
#pragma unroll
        for (int kk = 0; kk < PB / 16; ++kk) {
          umma_f16(tmem_base, umma_desc(a_st + kk * 2 * sbo_a, (uint32_t)a_blk_bytes, sbo_a, la),
                   umma_desc(b_st + kk * 2 * sbo_b, (uint32_t)b_blk_bytes, sbo_b, lb), idesc, (uint32_t)((i | kk) != 0));
        }
        umma_commit(sBar + 8 * (C::STAGES + s));
        if (i == my_tiles - 1) umma_commit(sBar + 8 * (2 * C::STAGES));
      }
      __syncwarp();
    }
    tc_fence_before();
  } else {
    // =========================================================== EPILOGUE: TMEM -> red.global.add
    if (my_tiles > 0) {
      const int q = warp & 3;
      mbar_wait(sBar + 8 * (2 * C::STAGES), 0);
      tc_fence_after();
      if (tid == 64) TRACE(7);
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
      const int k = k0 + q * 32 + lane;
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
    if (tid == 64) TRACE(10);
    tc_fence_before();
  }
  __syncthreads();
  if (tid == 0) TRACE(11);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

#ifdef RCFD_TRACE
}  // namespace
extern "C" int rcfd_debug_read_wtrace(long long* host_out, int n) {
  return cudaMemcpyFromSymbol(host_out, g_wtrace, sizeof(long long) * n) == cudaSuccess ? 0 : -2;
}
namespace {
#endif

template <int BN>
int launch_wg_tma(const WgTmaP& t, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& dy, float* dw,
                  cudaStream_t st) {
  note_kernel("wgrad_tma_kernel<%d>", BN);
  typedef WgTmaCfg<BN> C;
  static bool attr_set_dev[16] = {}; bool& attr_set = attr_set_dev[cur_dev()];   // per device: the attribute belongs to the device's copy of the kernel
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) { set_error("wgrad_tma: smem attribute: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
    attr_set = true;
  }
  const int gx = ceil_div(t.K, TM), gy = ceil_div(t.cout, BN);
  // one wave: every CTA occupies a whole SM (shared memory), so a 149th CTA would run alone after the others (ncu: 162 CTAs
  // for 256 -> 256 @22x44 = 1.09 waves, SMs busy 50 % of the kernel's duration)
  int splits = num_sms() / (gx * gy);
  const int max_splits = (t.num_ptiles + 3) / 4;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  dim3 grid(gx, gy, splits);
  wgrad_tma_kernel<BN><<<grid, NTHREADS, C::SMEM, st>>>(a0, a1, dy, t, dw);
  RCFD_CHECK_LAUNCH("wgrad_tma");
  return RCFD_OK;
}

}  // namespace

bool wgrad_tma_supported(const ConvKP& p, int dtype) {
  if (dtype != RCFD_BF16) return false;
  if (p.up || p.dil != 1) return false;
  if (p.stride != 1 && p.stride != 2) return false;
  if (p.c0 % 16 != 0 || p.c1 % 16 != 0) return false;
  if (p.cout % 16 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(p.src0) & 15) || (p.c1 > 0 && (reinterpret_cast<uintptr_t>(p.src1) & 15)) ||
      (reinterpret_cast<uintptr_t>(p.dst) & 15))
    return false;
  return get_encode() != nullptr;
}

int wgrad_tma_launch(const ConvKP& p, float* dw, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)p.cout * p.K * sizeof(float), st);
  if (e != cudaSuccess) { set_error("wgrad_tma memset: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
  WgTmaP t;
  t.n = p.n; t.ho = p.ho; t.wo = p.wo; t.cout = p.cout;
  t.kh = p.kh; t.kw = p.kw; t.stride = p.stride; t.pad = p.pad;
  t.c0 = p.c0; t.c1 = p.c1; t.K = p.K;
  auto gcd_ok = [&](int b) { return p.c0 % b == 0 && p.c1 % b == 0; };
  t.bkc = gcd_ok(64) ? 64 : (gcd_ok(32) ? 32 : 16);
  const int bn = p.cout > 128 ? 256 : (p.cout > 64 ? 128 : (p.cout > 32 ? 64 : (p.cout > 16 ? 32 : 16)));
  t.bnb = bn >= 64 ? 64 : bn;
  if (p.cout % t.bnb != 0) t.bnb = p.cout % 32 == 0 ? 32 : 16;     // dY boxes must tile cout exactly enough
  int best_tw = 8;
  long best_cov = -1;
  for (int tw = 32; tw >= 8; tw >>= 1) {
    const int th = PB / tw;
    const long cov = (long)ceil_div(p.wo, tw) * tw * ceil_div(p.ho, th) * th;
    if (best_cov < 0 || cov < best_cov) { best_cov = cov; best_tw = tw; }
  }
  t.tw = best_tw; t.th = PB / best_tw;
  t.tiles_x = ceil_div(p.wo, t.tw); t.tiles_y = ceil_div(p.ho, t.th);
  t.num_ptiles = p.n * t.tiles_x * t.tiles_y;
  alignas(64) CUtensorMap a0, a1, dy;
  if (!make_act_map(&a0, p.src0, p.n, p.hin, p.win, p.c0, t.bkc, t.tw, t.th, p.stride) ||
      !make_act_map(&a1, p.c1 > 0 ? p.src1 : p.src0, p.n, p.hin, p.win, p.c1 > 0 ? p.c1 : p.c0, t.bkc, t.tw, t.th, p.stride) ||
      !make_act_map(&dy, p.dst, p.n, p.ho, p.wo, p.cout, t.bnb, t.tw, t.th, 1)) {
    set_error("wgrad_tma: cuTensorMapEncodeTiled failed");
    return RCFD_ECUDA;
  }
  // the dY box width may be narrower than min(64, BN) when cout is not a multiple of it
  if (bn % t.bnb != 0) { set_error("wgrad_tma: internal tiling error"); return RCFD_EINVAL; }
  switch (bn) {
    case 256: return launch_wg_tma<256>(t, a0, a1, dy, dw, st);
    case 128: return launch_wg_tma<128>(t, a0, a1, dy, dw, st);
    case 64: return launch_wg_tma<64>(t, a0, a1, dy, dw, st);
    case 32: return launch_wg_tma<32>(t, a0, a1, dy, dw, st);
    default: return launch_wg_tma<16>(t, a0, a1, dy, dw, st);
  }
}

}  // namespace rcfd
