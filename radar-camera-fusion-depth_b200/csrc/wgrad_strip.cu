// Row-streaming weight gradient of the big-spatial 3x3 / stride-1 convolutions (bf16 operands, fp32
// accumulation in TMEM):
//
//     dW[co][r,s][ci] = sum over output pixels (y, x) of  dY[y][x][co] * X[y-1+r][x-1+s][ci]
//
// The gather / per-tap TMA weight-gradient engines fetch the input once per filter tap.  Here a
// persistent CTA walks DOWN a 128-pixel-wide column strip exactly like conv_strip.cu:
//   * warp 0 streams rows: one TMA box of the input row (CIN channels x 130 pixels) into a ring,
//     one TMA box of the dY row (cout channels x 128 pixels) into a second ring -- every element
//     is fetched ONCE per CTA;
//   * both operands are used MN-major, i.e. exactly as they lie in shared memory ([pixel][channel],
//     the reduction index K = pixel).  Filter tap (r, s) is ring row (y-1+r) with the descriptor
//     start address shifted by s pixels; one tcgen05.mma (M = 128) covers several taps at once by
//     using a leading-dimension byte offset of ONE PIXEL between its channel blocks:
//       CIN = 32:  M = 4 blocks x 32 channels = taps s = 0, 1, 2 (+ one ignored block),
//       CIN = 64:  M = 2 blocks x 64 channels = taps (0, 1) and (2, ignored);
//       CIN = 16:  M = 8 blocks x 16 channels = taps s = 0 .. 3 of the 4x4 stems (+ four ignored blocks);
//   * the accumulators (one per filter row / tap group) stay in TMEM for the whole kernel; the
//     epilogue warps add them to dW with coalesced fp32 red.global.add once at the end.
//
// Variant KS = 4: the 4x4 / stride-1 / pad-2 form of the 7x7 / stride-2 stems on their space-to-depth input (16 channels):
// four filter-row accumulators, three halo rows, window origin two pixels left of the strip.
//
// Variant UP: the convolution sits behind an exact 2x nearest up-sampling (decoder "deconv" convs).
// Streamed over LOW-RES input rows with the sub-pixel decomposition of conv_strip_up_kernel:
//     G[a,b][t,u][ci][co] = sum_{i,j} X[i-1+a+t][j-1+b+u][ci] * dY[2i+a][2j+b][co]
// (the four dY phases are fetched with element-stride-2 TMA boxes), then a small fold kernel sums
// the 16 G matrices into the nine dW taps (transpose of rcfd_pack_upconv2x_weight).
#include "tma_common.cuh"

namespace rcfd {
namespace {

using namespace tc;
using namespace tma;
constexpr int NTHREADS = 192;      // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
constexpr int SW = 128;            // strip width in output pixels (= K extent of one row step)

struct WgStripP {
  int n, h, w, cout;               // h, w: the grid the kernel walks (output grid; low-res grid when UP)
  int strips, rows_per_chunk, chunks_per_col, num_items;
  int ld_co, ld_tap, coff;         // element strides of the destination: dst[co * ld_co + tap * ld_tap + coff + ci]
};

template <int BN, int CIN, bool UP, int KS>
struct WgStripCfg {
  static constexpr int RB = CIN * 2;                                        // bytes per input pixel
  static constexpr int PAD = KS / 2;                                        // 3x3: 1, 4x4 stem: 2
  static constexpr int HALO_W = SW + KS - 1;                                // pixels one TMA box of an input row brings
  static constexpr int G = (KS == 3 && CIN == 64) ? 2 : 1;                  // tap groups per filter row (plain)
  // pixels the M = 128 window touches: 128 / CIN channel blocks one pixel apart, shifted by 2 (G - 1); the blocks past
  // tap KS - 1 are ignored but still read (CIN = 32: pixel 130, CIN = 16: pixels 131 .. 134)
  static constexpr int REACH = SW + 128 / CIN - 1 + 2 * (G - 1);
  static constexpr int ROWBUF = (((REACH > HALO_W ? REACH : HALO_W) * RB + 1023) / 1024) * 1024;
  static constexpr int DB = BN * 2;                                         // bytes per dY pixel
  static constexpr int DYBUF = SW * DB < 1024 ? 1024 : SW * DB;
  static constexpr int NB = UP ? 4 : 1;                                     // dY buffers per row step (phases)
  static constexpr int NRX = (UP && BN == 64) ? 4 : KS + 2;                 // ring of input rows
  static constexpr int NRD = (UP && BN == 64) ? 2 : 3;                      // ring of dY row steps
  static constexpr int MT = UP ? 8 : KS * G;                                // accumulator tiles
  static constexpr int COLS = MT * BN;
  static constexpr int TMEM_COLS = COLS <= 32 ? 32 : (COLS <= 64 ? 64 : (COLS <= 128 ? 128 : (COLS <= 256 ? 256 : 512)));
  static constexpr int SMEM = NRX * ROWBUF + NRD * NB * DYBUF + 1024 + 256;
  static constexpr uint32_t LAYOUT_A = RB == 128 ? 2u : (RB == 64 ? 4u : 6u);
  static constexpr uint32_t LAYOUT_B = DB == 128 ? 2u : (DB == 64 ? 4u : 6u);
  static_assert(COLS <= 512, "accumulators exceed TMEM");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static_assert(!UP || (CIN == 64 && KS == 3), "the up-sampling variant is instantiated for 64 input channels, 3x3");
  static_assert(KS == 3 || (KS == 4 && CIN == 16), "4x4 taps: the 16-channel space-to-depth stems only");
  static_assert(128 / CIN + 2 * (G - 1) >= KS, "the channel blocks of one filter row must cover its taps");
};

template <int BN, int CIN, bool UP, int KS>
__global__ void __launch_bounds__(NTHREADS)
wgrad_strip_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, const WgStripP p,
                   float* __restrict__ dst) {
  typedef WgStripCfg<BN, CIN, UP, KS> C;
  constexpr int NRX = C::NRX, NRD = C::NRD;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sRing = base;
  const uint32_t sDy = base + NRX * C::ROWBUF;
  const uint32_t sBar = sDy + NRD * C::NB * C::DYBUF;     // xfull[NRX], xempty[NRX], dfull[NRD], dempty[NRD], done
  const uint32_t bXF = sBar, bXE = sBar + 8 * NRX, bDF = sBar + 8 * 2 * NRX, bDE = bDF + 8 * NRD, bDone = bDE + 8 * NRD;
  const uint32_t sTmem = bDone + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (sTmem - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < NRX; ++s) { mbar_init(bXF + 8 * s, 1); mbar_init(bXE + 8 * s, 1); }
    for (int s = 0; s < NRD; ++s) { mbar_init(bDF + 8 * s, 1); mbar_init(bDE + 8 * s, 1); }
    mbar_init(bDone, 1);
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

  if (warp == 0) {
    // =========================================================== TMA PRODUCER
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
      uint32_t L = 0, D = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int ck = item % p.chunks_per_col;
        const int col = item / p.chunks_per_col;
        const int strip = col % p.strips;
        const int img = col / p.strips;
        const int y0 = ck * p.rows_per_chunk;
        const int rows = min(p.rows_per_chunk, p.h - y0);
        for (int j = 0; j < rows + KS - 1; ++j, ++L) {
          const int s = L % NRX;
          if (L >= (uint32_t)NRX) mbar_wait(bXE + 8 * s, ((L / NRX) & 1) ^ 1);
          mbar_expect_tx(bXF + 8 * s, (uint32_t)(C::HALO_W * C::RB));
          tma_load_4d(sRing + s * C::ROWBUF, &map_x, bXF + 8 * s, 0, strip * SW - C::PAD, y0 - C::PAD + j, img);
          if (j >= KS - 1) {
            const int t = j - (KS - 1);
            const int ds = D % NRD;
            if (D >= (uint32_t)NRD) mbar_wait(bDE + 8 * ds, ((D / NRD) & 1) ^ 1);
            mbar_expect_tx(bDF + 8 * ds, (uint32_t)(C::NB * SW * C::DB));
            const uint32_t dbuf = sDy + ds * C::NB * C::DYBUF;
            if (UP) {
#pragma unroll
              for (int ph = 0; ph < 4; ++ph)
                tma_load_4d(dbuf + ph * C::DYBUF, &map_dy, bDF + 8 * ds, 0, 2 * strip * SW + (ph & 1),
                            2 * (y0 + t) + (ph >> 1), img);
            } else {
              tma_load_4d(dbuf, &map_dy, bDF + 8 * ds, 0, strip * SW, y0 + t, img);
            }
            ++D;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================== MMA ISSUER
    const uint32_t idesc = umma_idesc_ex(TM, BN, /*a MN-major*/ 1, /*b MN-major*/ 1);
    constexpr uint32_t sbo_a = 8 * C::RB, sbo_b = 8 * C::DB;
    uint32_t L = 0, D = 0;
    bool fresh = true;               // the first row step of this CTA overwrites the accumulators
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int ck = item % p.chunks_per_col;
      const int y0 = ck * p.rows_per_chunk;
      const int rows = min(p.rows_per_chunk, p.h - y0);
#pragma unroll
      for (int i = 0; i < KS - 1; ++i) mbar_wait(bXF + 8 * ((L + i) % NRX), ((L + i) / NRX) & 1);
      for (int t = 0; t < rows; ++t, ++D) {
        const uint32_t Lnew = L + t + KS - 1;
        mbar_wait(bXF + 8 * (Lnew % NRX), (Lnew / NRX) & 1);
        const int ds = D % NRD;
        mbar_wait(bDF + 8 * ds, (D / NRD) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t dbuf = sDy + ds * C::NB * C::DYBUF;
          const uint32_t acc0 = fresh ? 0u : 1u;
#pragma unroll
          for (int mt = 0; mt < C::MT; ++mt) {
            // plain: mt = r * G + g   -> ring row r, window shift 2 g, dY buffer 0
            // up   : mt = (a*2+b)*2+t2 -> ring row a + t2, window shift b, dY buffer a*2+b
            const int xr = UP ? ((mt >> 2) + (mt & 1)) : (mt / C::G);
            const int shift = UP ? ((mt >> 1) & 1) : 2 * (mt % C::G);
            const int bb = UP ? (mt >> 1) : 0;
            const uint32_t a0 = sRing + ((L + t + xr) % NRX) * C::ROWBUF + shift * C::RB;
            const uint32_t b0 = dbuf + bb * C::DYBUF;
            const uint32_t d_tmem = tmem_base + mt * BN;
#pragma unroll
            for (int kk = 0; kk < SW / 16; ++kk) {
              umma_f16(d_tmem, umma_desc(a0 + kk * 2 * sbo_a, (uint32_t)C::RB, sbo_a, C::LAYOUT_A),
                       umma_desc(b0 + kk * 2 * sbo_b, (uint32_t)C::DYBUF, sbo_b, C::LAYOUT_B), idesc, kk == 0 ? acc0 : 1u);
            }
          }
          umma_commit(bDE + 8 * ds);                         // dY row step consumed
          umma_commit(bXE + 8 * ((L + t) % NRX));            // oldest input row no longer needed
          if (t == rows - 1) {
#pragma unroll
            for (int i = 1; i < KS; ++i) umma_commit(bXE + 8 * ((L + t + i) % NRX));
          }
        }
        fresh = false;
        __syncwarp();
      }
      L += rows + KS - 1;
    }
    if (lane == 0) umma_commit(bDone);
    __syncwarp();
    tc_fence_before();
  } else {
    // =========================================================== EPILOGUE: TMEM -> red.global.add (once)
    if ((int)blockIdx.x < p.num_items) {
      const int q = warp & 3;
      const int row = q * 32 + lane;                 // accumulator row = (channel block, channel)
      const int blk = row / CIN, ci = row - blk * CIN;
      mbar_wait(bDone, 0);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < C::MT; ++mt) {
        int tap;
        bool valid;
        if (UP) {                                     // tap16 = ((a*2+b)*2 + t2)*2 + u
          tap = mt * 2 + blk;
          valid = true;
        } else {
          const int r = mt / C::G, s = 2 * (mt % C::G) + blk;
          tap = r * KS + s;
          valid = s < KS;
        }
        const uint32_t trow = tmem_base + mt * BN + ((uint32_t)(q * 32) << 16);
        float* drow = dst + (size_t)tap * p.ld_tap + p.coff + ci;
#pragma unroll 1
        for (int cb = 0; cb < BN; cb += 16) {
          float v[16];
          tmem_ld16(trow + cb, v);
          if (valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int co = cb + i;
              if (co < p.cout) atomicAdd(drow + (size_t)co * p.ld_co, v[i]);
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

// G[co][16 = (a, b, t, u)][ci]  ->  dW[co][r*3+s][coff + ci] (+=): dW[r][s] = sum_{a,b} G[a,b][t(a,r)][u(b,s)]
// with t(0, .) = {0, 1, 1}, t(1, .) = {0, 0, 1} (rows of rcfd_pack_upconv2x_weight transposed).
__global__ void wgrad_up_fold_kernel(const float* __restrict__ g, float* __restrict__ dw, int cout, int cin, int ld_co,
                                     int ld_tap, int coff) {
  const int total = cout * 9 * cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % cin;
    const int tap = (i / cin) % 9;
    const int co = i / (9 * cin);
    const int r = tap / 3, s = tap - r * 3;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int t = a == 0 ? (r >= 1) : (r >= 2);
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int u = b == 0 ? (s >= 1) : (s >= 2);
        acc += g[((size_t)co * 16 + ((a * 2 + b) * 2 + t) * 2 + u) * cin + ci];
      }
    }
    dw[(size_t)co * ld_co + tap * ld_tap + coff + ci] += acc;
  }
}

inline bool make_x_row_map(CUtensorMap* m, const void* ptr, int n, int h, int w, int c, int halo_w) {
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  cuuint32_t box[4] = {(cuuint32_t)c, (cuuint32_t)halo_w, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(c), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// dY rows: BN channels x 128 pixels (every pixel, or every second pixel for the sub-pixel phases)
inline bool make_dy_row_map(CUtensorMap* m, const void* ptr, int n, int h, int w, int c, int bn, int xstride) {
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  cuuint32_t box[4] = {(cuuint32_t)bn, (cuuint32_t)(SW * xstride), 1, 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)xstride, 1, 1};
  return get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bn), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

void plan_items(WgStripP& t, int ctas) {
  const int cols = t.n * t.strips;
  // every chunk re-reads 2 halo rows and the CTA flushes its accumulators once at the end
  plan_row_chunks(t.h, cols, ctas, 8, 8, &t.rows_per_chunk, &t.chunks_per_col);
  t.num_items = cols * t.chunks_per_col;
}

template <int BN, int CIN, bool UP, int KS = 3>
int launch_wg_strip(const void* x, int n, int hx, int wx, const void* dy, int hd, int wd, int cout, WgStripP& t, float* dst,
                    cudaStream_t st) {
  note_kernel("wgrad_strip_kernel<%d,%d,%d,%d>", BN, CIN, (int)UP, KS);
  typedef WgStripCfg<BN, CIN, UP, KS> C;
  static bool attr_set_dev[16] = {}; bool& attr_set = attr_set_dev[cur_dev()];   // per device: the attribute belongs to the device's copy of the kernel
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_strip_kernel<BN, CIN, UP, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) { set_error("wgrad_strip: smem attribute: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
    attr_set = true;
  }
  alignas(64) CUtensorMap mx, md;
  if (!make_x_row_map(&mx, x, n, hx, wx, CIN, C::HALO_W) || !make_dy_row_map(&md, dy, n, hd, wd, cout, BN, UP ? 2 : 1)) {
    set_error("wgrad_strip: cuTensorMapEncodeTiled failed");
    return RCFD_ECUDA;
  }
  // two CTAs per SM hide each other's barrier round trips when shared memory and TMEM allow it
  const int per_sm = (C::SMEM <= 112 * 1024 && C::TMEM_COLS <= 256) ? 2 : 1;
  int ctas = num_sms() * per_sm;
  plan_items(t, ctas);
  if (ctas > t.num_items) ctas = t.num_items;
  wgrad_strip_kernel<BN, CIN, UP, KS><<<ctas, NTHREADS, C::SMEM, st>>>(mx, md, t, dst);
  RCFD_CHECK_LAUNCH("wgrad_strip");
  return RCFD_OK;
}

template <bool UP>
int dispatch_wg_strip(int cin, int bn, const void* x, int n, int hx, int wx, const void* dy, int hd, int wd, int cout,
                      WgStripP& t, float* dst, cudaStream_t st) {
  if (cin == 64) {
    switch (bn) {
      case 64: return launch_wg_strip<64, 64, UP>(x, n, hx, wx, dy, hd, wd, cout, t, dst, st);
      case 32: return launch_wg_strip<32, 64, UP>(x, n, hx, wx, dy, hd, wd, cout, t, dst, st);
      default: return launch_wg_strip<16, 64, UP>(x, n, hx, wx, dy, hd, wd, cout, t, dst, st);
    }
  }
  if (!UP && cin == 16) {            // 4x4 space-to-depth stems
    switch (bn) {
      case 32: return launch_wg_strip<32, 16, false, 4>(x, n, hx, wx, dy, hd, wd, cout, t, dst, st);
      case 16: return launch_wg_strip<16, 16, false, 4>(x, n, hx, wx, dy, hd, wd, cout, t, dst, st);
      default: break;
    }
  } else if (!UP) {
    switch (bn) {
      case 64: return launch_wg_strip<64, 32, false>(x, n, hx, wx, dy, hd, wd, cout, t, dst, st);
      case 32: return launch_wg_strip<32, 32, false>(x, n, hx, wx, dy, hd, wd, cout, t, dst, st);
      default: return launch_wg_strip<16, 32, false>(x, n, hx, wx, dy, hd, wd, cout, t, dst, st);
    }
  }
  set_error("wgrad_strip: unsupported channel count");
  return RCFD_EUNSUPPORTED;
}

bool chan_ok(int c) { return c == 32 || c == 64; }

}  // namespace

// plain: 3x3 / stride 1 / pad 1, one or two sources with 32 or 64 channels each, cout in {16, 32, 64}
// stem : 4x4 / stride 1 / pad 2 on ONE 16-channel (space-to-depth) source, output cropped to the input grid, cout in {16, 32}
bool wgrad_strip_supported(const ConvKP& p, int dtype) {
  if (dtype != RCFD_BF16 || p.dil != 1) return false;
  const bool stem = p.kh == 4 && p.kw == 4 && p.stride == 1 && p.pad == 2 && !p.up && p.c0 == 16 && p.c1 == 0 &&
                    (p.cout == 16 || p.cout == 32);
  if (!stem && (p.kh != 3 || p.kw != 3 || p.stride != 1 || p.pad != 1)) return false;
  if (p.ho != p.hin || p.wo != p.win) return false;
  if (p.cout != 16 && p.cout != 32 && p.cout != 64) return false;
  if ((reinterpret_cast<uintptr_t>(p.src0) & 15) || (reinterpret_cast<uintptr_t>(p.dst) & 15)) return false;
  if (p.c1 > 0 && (reinterpret_cast<uintptr_t>(p.src1) & 15)) return false;
  if (get_encode() == nullptr) return false;
  if (p.up) return p.c1 == 0 && p.c0 == 64 && p.hin == 2 * p.h0 && p.win == 2 * p.w0;
  return stem || (chan_ok(p.c0) && (p.c1 == 0 || chan_ok(p.c1)));
}

int64_t wgrad_strip_workspace(const ConvKP& p, int dtype) {
  if (!wgrad_strip_supported(p, dtype) || !p.up) return 0;
  return (int64_t)p.cout * 16 * p.c0 * (int64_t)sizeof(float);
}

// worth it where the 128-wide strips fit the row reasonably and there are enough rows to stream
bool wgrad_strip_preferred(const ConvKP& p, int dtype) {
  if (!wgrad_strip_supported(p, dtype)) return false;
  const int w = p.up ? p.w0 : p.wo, h = p.up ? p.h0 : p.ho;
  const int strips = ceil_div(w, SW);
  return h >= 64 && (long)strips * SW * 2 <= (long)w * 3;
}

int wgrad_strip_launch(const ConvKP& p, float* dw, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)p.cout * p.K * sizeof(float), st);
  if (e != cudaSuccess) { set_error("wgrad_strip memset: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
  const int bn = p.cout;
  WgStripP t;
  t.n = p.n; t.cout = p.cout;
  if (p.up) {
    const int64_t need = wgrad_strip_workspace(p, RCFD_BF16);
    RCFD_CHECK_ARG(workspace != nullptr && workspace_bytes >= need, "wgrad_strip: workspace of %lld bytes required",
                   (long long)need);
    e = cudaMemsetAsync(workspace, 0, (size_t)need, st);
    if (e != cudaSuccess) { set_error("wgrad_strip memset: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
    t.h = p.h0; t.w = p.w0; t.strips = ceil_div(p.w0, SW);
    t.ld_co = 16 * p.c0; t.ld_tap = p.c0; t.coff = 0;
    int rc = dispatch_wg_strip<true>(p.c0, bn, p.src0, p.n, p.h0, p.w0, p.dst, p.ho, p.wo, p.cout, t,
                                     reinterpret_cast<float*>(workspace), st);
    if (rc != RCFD_OK) return rc;
    const int total = p.cout * 9 * p.c0;
    wgrad_up_fold_kernel<<<ceil_div(total, 256), 256, 0, st>>>(reinterpret_cast<const float*>(workspace), dw, p.cout, p.c0,
                                                               p.K, p.ctot, 0);
    RCFD_CHECK_LAUNCH("wgrad_up_fold");
    return RCFD_OK;
  }
  t.h = p.ho; t.w = p.wo; t.strips = ceil_div(p.wo, SW);
  t.ld_co = p.K; t.ld_tap = p.ctot; t.coff = 0;
  int rc = dispatch_wg_strip<false>(p.c0, bn, p.src0, p.n, p.hin, p.win, p.dst, p.ho, p.wo, p.cout, t, dw, st);
  if (rc != RCFD_OK || p.c1 == 0) return rc;
  t.coff = p.c0;
  return dispatch_wg_strip<false>(p.c1, bn, p.src1, p.n, p.hin, p.win, p.dst, p.ho, p.wo, p.cout, t, dw, st);
}

}  // namespace rcfd
