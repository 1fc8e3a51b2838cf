// Row-streaming 3x3 convolution: TMA + tcgen05 with the input kept in a rolling window of
// shared-memory rows, so every input pixel is fetched ONCE per CTA instead of once per tap.
//
// The generic TMA engine (conv_tma.cu) issues one box per filter tap: the A operand is
// re-read 9x from L2 and the TMA unit (about one <=128-byte box row per 5 cycles) becomes the
// bound.  For the big-spatial 3x3 / stride-1 layers (FusionNet decoder at 176x352 and 352x704
// and their dgrads) this kernel instead walks DOWN a 128-pixel-wide column strip:
//   * warp 0 streams input rows: one TMA box (C channels x 130 pixels x 1 row) per image row
//     into a ring of row buffers (zero rows above / below the image come from TMA OOB fill,
//     the left / right halo pixel likewise);
//   * the 3x3 weights of this CTA's cout tile live in shared memory for the whole kernel;
//   * warp 1 issues, per OUTPUT row, 9 x (C/16) tcgen05.mma (M = 128 pixels, N = cout tile,
//     K = 16): tap (r, s) reads ring row (y-1+r) through a shared-memory descriptor whose start
//     address is simply shifted by s pixels -- the 128/64-byte swizzle is a function of the
//     absolute shared-memory address, so the shifted window needs no data movement;
//   * warps 2-5 drain the double-buffered TMEM accumulator (BN statistics / folded BN +
//     activation + residual, 16-byte stores) while the next row is being multiplied.
// HBM/L2 traffic per CTA: input rows once (+2 halo rows per row chunk, +2 halo pixels per
// strip), output once -- the convolution's algorithmic minimum.
#include "tma_common.cuh"

namespace rcfd {
namespace {

using namespace tc;
using namespace tma;
constexpr int NTHREADS = 192;
constexpr int NEPI = 128;
constexpr int SW = 128;            // strip width in output pixels == UMMA_M
constexpr int HALO_W = SW + 2;
constexpr int NR = 6;              // ring of input rows

struct StripP {
  int n, h, w, cin, cout;
  int strips, rows_per_chunk, chunks_per_col, num_items;
  int desc_mode;                   // 0: base_offset 0 (address-based swizzle), 1: base_offset = (start >> 7) & 7
  void* dst;
  const float* scale; const float* shift;
  int act; float p0, p1;
  const void* residual;
  double* ssum; double* ssq;
  int accumulate, dst_f32;
};

template <int BN, int CIN, int KS = 3, int CIN1 = 0>
struct StripCfg {
  static constexpr int RB = CIN * 2;                                        // bytes per pixel row, source 0
  static constexpr int RB1 = CIN1 * 2;                                      // ... of the concat partner (0: none)
  static constexpr int HALO = SW + KS - 1;                                  // pixels per ring row
  static constexpr int ROWBUF0 = ((HALO * RB + 1023) / 1024) * 1024;
  static constexpr int ROWBUF1 = CIN1 ? ((HALO * RB1 + 1023) / 1024) * 1024 : 0;
  static constexpr int ROWBUF = ROWBUF0 + ROWBUF1;
  static constexpr int W_TAP0 = ((BN * RB + 1023) / 1024) * 1024;
  static constexpr int W_TAP1 = CIN1 ? ((BN * RB1 + 1023) / 1024) * 1024 : 0;
  static constexpr int W_TAP = W_TAP0 + W_TAP1;
  static constexpr int W_BYTES = KS * KS * W_TAP;
  static constexpr int RED_BYTES = 2 * 4 * BN * 2 * 4;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int RING = CIN1 ? 4 : NR;                                // two sources: the weights take the room
  static constexpr int SMEM = RING * ROWBUF + W_BYTES + RED_BYTES + 1024 + 256;
  static constexpr uint32_t LAYOUT = RB == 128 ? 2u : (RB == 64 ? 4u : 6u);
  static constexpr uint32_t LAYOUT1 = RB1 == 128 ? 2u : (RB1 == 64 ? 4u : 6u);
};

__device__ __forceinline__ uint64_t strip_desc(uint32_t saddr, uint32_t sbo, uint32_t layout, int mode) {
  uint64_t d = umma_desc(saddr, 16, sbo, layout);
  if (mode == 1) d |= (uint64_t)((saddr >> 7) & 7u) << 49;
  return d;
}


// ---- epilogue helpers shared by the two kernels -------------------------------------------------------
// 16 accumulator columns of one pixel -> folded BN / activation / residual / accumulate -> global
__device__ __forceinline__ void strip_store16(float (&v)[16], const StripP& p, int nb, size_t o, bool vector_epilogue) {
  const bf16* R = reinterpret_cast<const bf16*>(p.residual);
  bf16* D = reinterpret_cast<bf16*>(p.dst);
  if (!vector_epilogue) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int n = nb + i;
      if (n < p.cout) {
        float x = v[i];
        if (p.scale) x = fmaf(x, __ldg(p.scale + n), __ldg(p.shift + n));
        x = apply_act(x, p.act, p.p0, p.p1);
        if (R) x = leaky(x + __bfloat162float(R[o + i]));
        if (p.dst_f32) {
          float* Df = reinterpret_cast<float*>(p.dst);
          Df[o + i] = p.accumulate ? Df[o + i] + x : x;
        } else {
          D[o + i] = __float2bfloat16_rn(p.accumulate ? __bfloat162float(D[o + i]) + x : x);
        }
      }
    }
  } else if (nb < p.cout) {
    if (p.scale) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], __ldg(p.scale + nb + i), __ldg(p.shift + nb + i));
    }
    if (p.act == RCFD_ACT_LEAKY) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = leaky(v[i]);
    } else if (p.act == RCFD_ACT_SIGMOID) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = sigmoid_precise(v[i]);
    }
    if (R) {
      const uint4 r0 = *reinterpret_cast<const uint4*>(R + o);
      const uint4 r1 = *reinterpret_cast<const uint4*>(R + o + 8);
      const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr[i]));
        v[2 * i] = leaky(v[2 * i] + f.x);
        v[2 * i + 1] = leaky(v[2 * i + 1] + f.y);
      }
    }
    if (p.accumulate) {
      const uint4 r0 = *reinterpret_cast<const uint4*>(D + o);
      const uint4 r1 = *reinterpret_cast<const uint4*>(D + o + 8);
      const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr[i]));
        v[2 * i] += f.x;
        v[2 * i + 1] += f.y;
      }
    }
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      pk[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(D + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(D + o + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

// BatchNorm batch statistics.  Every epilogue thread owns ONE pixel column of the strip (its TMEM lane) and
// sees all BN channels of it, row after row: the per-channel sum / sum of squares are kept in REGISTERS for
// the whole life of the CTA (no shuffles, barriers or atomics per row) and reduced once at the end:
// a transposing butterfly over the 32 lanes (16 channels at a time), shared memory over the 4 warps, then
// one fp64 atomicAdd per channel per CTA.
template <int BN>
__device__ __forceinline__ void strip_flush_stats(float (&ss)[BN], float (&sq)[BN], float* red, int q, int lane, int etid,
                                                  int n0, const StripP& p) {
#pragma unroll
  for (int cb = 0; cb < BN; cb += 16) {
    float s16[16], q16[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      s16[i] = ss[cb + i] + __shfl_xor_sync(0xffffffffu, ss[cb + i], 16);
      q16[i] = sq[cb + i] + __shfl_xor_sync(0xffffffffu, sq[cb + i], 16);
    }
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1) {
      const bool hi = (lane & w) != 0;
#pragma unroll
      for (int i = 0; i < w; ++i) {
        const float send_s = hi ? s16[i] : s16[i + w];
        const float keep_s = hi ? s16[i + w] : s16[i];
        s16[i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, w);
        const float send_q = hi ? q16[i] : q16[i + w];
        const float keep_q = hi ? q16[i + w] : q16[i];
        q16[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, w);
      }
    }
    if (lane < 16) {
      red[(q * BN + cb + lane) * 2 + 0] = s16[0];
      red[(q * BN + cb + lane) * 2 + 1] = q16[0];
    }
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  for (int c = etid; c < BN; c += NEPI) {
    if (n0 + c < p.cout) {
      double s = 0.0, qq = 0.0;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        s += (double)red[(w * BN + c) * 2 + 0];
        qq += (double)red[(w * BN + c) * 2 + 1];
      }
      atomicAdd(p.ssum + n0 + c, s);
      atomicAdd(p.ssq + n0 + c, qq);
    }
  }
}

// CIN1 > 0: second source tensor (the skip connection of torch.cat([x, skip], 1), reference src/net_utils.py:565):
// every ring slot holds the row of both sources, every tap multiplies both channel groups.
template <int BN, int CIN, bool STATS, int KS, int CIN1>
__global__ void __launch_bounds__(NTHREADS)
conv_strip_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                  const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_w1, const StripP p) {
  typedef StripCfg<BN, CIN, KS, CIN1> C;
  constexpr int NR = C::RING;
  constexpr int PAD = KS / 2;            // 3x3 / pad 1, or the 4x4 / pad 2 window of the space-to-depth stems
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sRing = base;
  const uint32_t sW = base + NR * C::ROWBUF;
  const uint32_t sRed = sW + C::W_BYTES;
  const uint32_t sBar = sRed + C::RED_BYTES;   // full[NR], empty[NR], tfull[2], tempty[2], wbar
  const uint32_t sTmem = sBar + 8 * (2 * NR + 5);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (sTmem - base));
  float* red = reinterpret_cast<float*>(gen_base + (sRed - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * BN;
  const uint32_t wbar = sBar + 8 * (2 * NR + 4);

  if (tid == 0) {
    for (int s = 0; s < NR; ++s) {
      mbar_init(sBar + 8 * s, 1);
      mbar_init(sBar + 8 * (NR + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(sBar + 8 * (2 * NR + a), 1);
      mbar_init(sBar + 8 * (2 * NR + 2 + a), NEPI);
    }
    mbar_init(wbar, 1);
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
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
      if constexpr (CIN1 > 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
      }
      // resident weights: 9 taps x [BN][CIN]
      mbar_expect_tx(wbar, (uint32_t)(KS * KS * BN * (C::RB + C::RB1)));
      for (int tap = 0; tap < KS * KS; ++tap) {
        tma_load_2d(sW + tap * C::W_TAP, &map_w, wbar, tap * p.cin, n0);
        if constexpr (CIN1 > 0) tma_load_2d(sW + tap * C::W_TAP + C::W_TAP0, &map_w1, wbar, tap * p.cin + CIN, n0);
      }
      uint32_t L = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int ck = item % p.chunks_per_col;
        int col = item / p.chunks_per_col;
        const int strip = col % p.strips;
        const int img = col / p.strips;
        const int y0 = ck * p.rows_per_chunk;
        const int rows = min(p.rows_per_chunk, p.h - y0);
        for (int j = 0; j < rows + KS - 1; ++j, ++L) {
          const int s = L % NR;
          if (L >= (uint32_t)NR) mbar_wait(sBar + 8 * (NR + s), ((L / NR) & 1) ^ 1);
          mbar_expect_tx(sBar + 8 * s, (uint32_t)(C::HALO * (C::RB + C::RB1)));
          tma_load_4d(sRing + s * C::ROWBUF, &map_x, sBar + 8 * s, 0, strip * SW - PAD, y0 - PAD + j, img);
          if constexpr (CIN1 > 0)
            tma_load_4d(sRing + s * C::ROWBUF + C::ROWBUF0, &map_x1, sBar + 8 * s, 0, strip * SW - PAD, y0 - PAD + j, img);
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================== MMA ISSUER
    const uint32_t idesc = umma_idesc(BN);
    constexpr uint32_t sbo = 8 * C::RB;
    mbar_wait(wbar, 0);
    uint32_t L = 0, orow = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int ck = item % p.chunks_per_col;
      const int y0 = ck * p.rows_per_chunk;
      const int rows = min(p.rows_per_chunk, p.h - y0);
      // input rows L .. L+rows+KS-2 of this chunk; output row t uses L+t .. L+t+KS-1
#pragma unroll
      for (int j = 0; j < KS - 1; ++j) mbar_wait(sBar + 8 * ((L + j) % NR), ((L + j) / NR) & 1);
      for (int t = 0; t < rows; ++t, ++orow) {
        const uint32_t Lnew = L + t + KS - 1;
        mbar_wait(sBar + 8 * (Lnew % NR), (Lnew / NR) & 1);
        const uint32_t acc = orow & 1;
        if (orow >= 2) mbar_wait(sBar + 8 * (2 * NR + 2 + acc), ((orow >> 1) & 1) ^ 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t d_tmem = tmem_base + acc * BN;
#pragma unroll
          for (int r = 0; r < KS; ++r) {
            const uint32_t rowbuf = sRing + ((L + t + r) % NR) * C::ROWBUF;
#pragma unroll
            for (int s = 0; s < KS; ++s) {
              const uint32_t a0 = rowbuf + s * C::RB;
              const uint32_t b0 = sW + (r * KS + s) * C::W_TAP;
#pragma unroll
              for (int k = 0; k < CIN / 16; ++k) {
                umma_f16(d_tmem, strip_desc(a0 + k * 32, sbo, C::LAYOUT, p.desc_mode), umma_desc(b0 + k * 32, 16, sbo, C::LAYOUT),
                         idesc, (uint32_t)((r | s | k) != 0));
              }
              if constexpr (CIN1 > 0) {
                constexpr uint32_t sbo1 = 8 * C::RB1;
                const uint32_t a1 = rowbuf + C::ROWBUF0 + s * C::RB1;
                const uint32_t b1 = b0 + C::W_TAP0;
#pragma unroll
                for (int k = 0; k < CIN1 / 16; ++k) {
                  umma_f16(d_tmem, strip_desc(a1 + k * 32, sbo1, C::LAYOUT1, p.desc_mode),
                           umma_desc(b1 + k * 32, 16, sbo1, C::LAYOUT1), idesc, 1u);
                }
              }
            }
          }
          umma_commit(sBar + 8 * (2 * NR + acc));                  // accumulator of this output row complete
          umma_commit(sBar + 8 * (NR + (L + t) % NR));             // oldest input row no longer needed
          if (t == rows - 1) {                                      // chunk done: release its last KS-1 rows too
#pragma unroll
            for (int j = 1; j < KS; ++j) umma_commit(sBar + 8 * (NR + (L + t + j) % NR));
          }
        }
        __syncwarp();
      }
      L += rows + KS - 1;
    }
    tc_fence_before();
  } else {
    // =========================================================== EPILOGUE (warps 2..5)
    const int q = warp & 3;
    const int r = q * 32 + lane;               // pixel within the strip == TMEM lane
    const bool vector_epilogue = (p.cout % 16 == 0) && !p.dst_f32 && p.act != RCFD_ACT_DEPTH_HEAD;
    const int etid = tid - 64;
    float ss[STATS ? BN : 1], sq[STATS ? BN : 1];
    if constexpr (STATS) {
#pragma unroll
      for (int i = 0; i < BN; ++i) { ss[i] = 0.f; sq[i] = 0.f; }
    }
    uint32_t orow = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int ck = item % p.chunks_per_col;
      int col = item / p.chunks_per_col;
      const int strip = col % p.strips;
      const int img = col / p.strips;
      const int y0 = ck * p.rows_per_chunk;
      const int rows = min(p.rows_per_chunk, p.h - y0);
      const int ox = strip * SW + r;
      const bool mvalid = ox < p.w;
      for (int t = 0; t < rows; ++t, ++orow) {
        const uint32_t acc = orow & 1;
        const size_t gm = ((size_t)img * p.h + (y0 + t)) * p.w + ox;
        mbar_wait(sBar + 8 * (2 * NR + acc), (orow >> 1) & 1);
        tc_fence_after();
        const uint32_t trow = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
        if constexpr (STATS) {
#pragma unroll
          for (int cb = 0; cb < BN; cb += 16) {          // unrolled: ss / sq must be indexed statically
            float v[16];
            tmem_ld16(trow + cb, v);
            if (mvalid) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                ss[cb + i] += v[i];
                sq[cb + i] = fmaf(v[i], v[i], sq[cb + i]);
              }
              strip_store16(v, p, n0 + cb, gm * p.cout + n0 + cb, vector_epilogue);
            }
          }
        } else {
#pragma unroll 1
          for (int cb = 0; cb < BN; cb += 16) {
            float v[16];
            tmem_ld16(trow + cb, v);
            if (mvalid) strip_store16(v, p, n0 + cb, gm * p.cout + n0 + cb, vector_epilogue);
          }
        }
        tc_fence_before();
        mbar_arrive(sBar + 8 * (2 * NR + 2 + acc));
      }
    }
    if constexpr (STATS) strip_flush_stats<BN>(ss, sq, red, q, lane, etid, n0, p);
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

constexpr int NRU = 4;             // ring depth of the up-sampling variant (16 weight blocks need the room)

// Sub-pixel phases of `3x3 after 2x nearest up-sampling`: phase (a, b), tap (t2, u) reads the input window
// (row a + t2, column shift b + u) -- 9 distinct windows for 16 (phase, tap) pairs.  Pairs that share a window are issued
// as ONE tcgen05.mma whose B operand stacks their weight blocks along N and whose D spans their (adjacent) phase
// accumulators: 10 MMAs per 16-channel chunk instead of 16, i.e. 37 % fewer fetches of the 128-pixel A tile from shared
// memory (the unit that binds this kernel, profiles/r1_ncu_summary.md).  Phase accumulators sit in TMEM in the order
// 0, 1, 3, 2 so that three of the four two-phase windows are adjacent; the fourth is split.
// Weight blocks in shared memory, slot -> (phase, tap):
//   0-3 centre window (1,1): (0,3) (1,2) (3,0) (2,1)      4-5 window (0,1): (0,1) (1,0)      6-7 window (1,2): (1,3) (3,1)
//   8-9 window (2,1): (3,2) (2,3)      10 / 11 window (1,0): (0,2) / (2,0)      12-15 corners: (0,0) (1,1) (2,2) (3,3)
__device__ __constant__ int8_t kUpSlotPhase[16] = {0, 1, 3, 2, 0, 1, 1, 3, 3, 2, 0, 2, 0, 1, 2, 3};
__device__ __constant__ int8_t kUpSlotTap[16] = {3, 2, 0, 1, 1, 0, 3, 1, 2, 3, 2, 0, 0, 1, 2, 3};
// MMA list: window row, window column shift, first weight slot, blocks stacked along N, first TMEM phase position
__device__ __constant__ int8_t kUpMma[10][5] = {{1, 1, 0, 4, 0}, {0, 1, 4, 2, 0}, {1, 2, 6, 2, 1}, {2, 1, 8, 2, 2}, {1, 0, 10, 1, 0},
                                              {1, 0, 11, 1, 3}, {0, 0, 12, 1, 0}, {0, 2, 13, 1, 1}, {2, 0, 14, 1, 3}, {2, 2, 15, 1, 2}};
__device__ __forceinline__ int up_phase_pos(int ph) { return ph < 2 ? ph : 5 - ph; }     // TMEM order 0, 1, 3, 2

template <int BN, int CIN>
struct StripUpCfg {
  static constexpr int RB = CIN * 2;
  static constexpr int ROWBUF = ((HALO_W * RB + 1023) / 1024) * 1024;
  static constexpr int W_TAP = ((BN * RB + 1023) / 1024) * 1024;
  static_assert(W_TAP == BN * RB, "stacked weight blocks must be contiguous (merged sub-pixel MMAs)");
  static constexpr int W_BYTES = 16 * W_TAP;
  static constexpr int RED_BYTES = 2 * 4 * BN * 2 * 4;
  static constexpr int TMEM_COLS = 8 * BN;                   // 2 buffers x 4 phases
  static constexpr int SMEM = NRU * ROWBUF + W_BYTES + RED_BYTES + 1024 + 256;
  static constexpr uint32_t LAYOUT = RB == 128 ? 2u : (RB == 64 ? 4u : 6u);
};

// 3x3 conv behind an exact 2x nearest up-sampling, streamed over LOW-RES rows: per low-res row
// the four sub-pixel phases (2x2 taps each, summed weights) are accumulated into four TMEM
// buffers and written to output rows 2i, 2i+1 / columns 2j, 2j+1.
template <int BN, int CIN, bool STATS>
__global__ void __launch_bounds__(NTHREADS)
conv_strip_up_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const StripP p) {
  typedef StripUpCfg<BN, CIN> C;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sRing = base;
  const uint32_t sW = base + NRU * C::ROWBUF;
  const uint32_t sRed = sW + C::W_BYTES;
  const uint32_t sBar = sRed + C::RED_BYTES;   // full[NRU], empty[NRU], tfull[2], tempty[2], wbar
  const uint32_t sTmem = sBar + 8 * (2 * NRU + 5);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (sTmem - base));
  float* red = reinterpret_cast<float*>(gen_base + (sRed - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * BN;
  const uint32_t wbar = sBar + 8 * (2 * NRU + 4);

  if (tid == 0) {
    for (int s = 0; s < NRU; ++s) {
      mbar_init(sBar + 8 * s, 1);
      mbar_init(sBar + 8 * (NRU + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(sBar + 8 * (2 * NRU + a), 1);
      mbar_init(sBar + 8 * (2 * NRU + 2 + a), NEPI);
    }
    mbar_init(wbar, 1);
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
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
      // resident weights: 4 phases x 4 taps x [BN][CIN]  (rcfd_pack_upconv2x_weight layout [ph][cout][tap][cin])
      mbar_expect_tx(wbar, (uint32_t)(16 * BN * C::RB));
#pragma unroll
      for (int slot = 0; slot < 16; ++slot)
        tma_load_2d(sW + slot * C::W_TAP, &map_w, wbar, kUpSlotTap[slot] * p.cin, kUpSlotPhase[slot] * p.cout + n0);
      uint32_t L = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int ck = item % p.chunks_per_col;
        int col = item / p.chunks_per_col;
        const int strip = col % p.strips;
        const int img = col / p.strips;
        const int y0 = ck * p.rows_per_chunk;
        const int rows = min(p.rows_per_chunk, p.h - y0);
        for (int j = 0; j < rows + 2; ++j, ++L) {
          const int s = L % NRU;
          if (L >= (uint32_t)NRU) mbar_wait(sBar + 8 * (NRU + s), ((L / NRU) & 1) ^ 1);
          mbar_expect_tx(sBar + 8 * s, (uint32_t)(HALO_W * C::RB));
          tma_load_4d(sRing + s * C::ROWBUF, &map_x, sBar + 8 * s, 0, strip * SW - 1, y0 - 1 + j, img);
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================== MMA ISSUER
    const uint32_t idesc = umma_idesc(BN);
    constexpr uint32_t sbo = 8 * C::RB;
    mbar_wait(wbar, 0);
    uint32_t L = 0, orow = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int ck = item % p.chunks_per_col;
      const int y0 = ck * p.rows_per_chunk;
      const int rows = min(p.rows_per_chunk, p.h - y0);
      // input rows L .. L+rows+1 of this chunk; output row t uses L+t, L+t+1, L+t+2
      mbar_wait(sBar + 8 * (L % NRU), (L / NRU) & 1);
      mbar_wait(sBar + 8 * ((L + 1) % NRU), ((L + 1) / NRU) & 1);
      for (int t = 0; t < rows; ++t, ++orow) {
        const uint32_t Lnew = L + t + 2;
        mbar_wait(sBar + 8 * (Lnew % NRU), (Lnew / NRU) & 1);
        const uint32_t acc = orow & 1;
        if (orow >= 2) mbar_wait(sBar + 8 * (2 * NRU + 2 + acc), ((orow >> 1) & 1) ^ 1);
        tc_fence_after();
        if (lane == 0) {
          // phase (a, b): output pixel (2i+a, 2j+b) = 2x2 taps on low-res rows i-1+a, i+a and columns j-1+b, j+b;
          // issued window by window (see kUpMma); the centre window comes first and overwrites all four accumulators
#pragma unroll
          for (int k = 0; k < CIN / 16; ++k) {
#pragma unroll
            for (int m = 0; m < 10; ++m) {
              const int wr = kUpMma[m][0], wc = kUpMma[m][1], slot = kUpMma[m][2], nb = kUpMma[m][3], pos = kUpMma[m][4];
              const uint32_t a0 = sRing + ((L + t + wr) % NRU) * C::ROWBUF + wc * C::RB;
              const uint32_t b0 = sW + slot * C::W_TAP;
              umma_f16(tmem_base + (acc * 4 + pos) * BN, umma_desc(a0 + k * 32, 16, sbo, C::LAYOUT),
                       umma_desc(b0 + k * 32, 16, sbo, C::LAYOUT), umma_idesc(nb * BN), (uint32_t)((m | k) != 0));
            }
          }
          umma_commit(sBar + 8 * (2 * NRU + acc));                  // accumulator of this output row complete
          umma_commit(sBar + 8 * (NRU + (L + t) % NRU));             // oldest input row no longer needed
          if (t == rows - 1) {                                      // chunk done: release its last two rows too
            umma_commit(sBar + 8 * (NRU + (L + t + 1) % NRU));
            umma_commit(sBar + 8 * (NRU + (L + t + 2) % NRU));
          }
        }
        __syncwarp();
      }
      L += rows + 2;
    }
    tc_fence_before();
  } else {
    // =========================================================== EPILOGUE (warps 2..5)
    const int q = warp & 3;
    const int r = q * 32 + lane;               // pixel within the strip == TMEM lane
    const bool vector_epilogue = (p.cout % 16 == 0) && !p.dst_f32 && p.act != RCFD_ACT_DEPTH_HEAD;
    const int etid = tid - 64;
    float ss[STATS ? BN : 1], sq[STATS ? BN : 1];
    if constexpr (STATS) {
#pragma unroll
      for (int i = 0; i < BN; ++i) { ss[i] = 0.f; sq[i] = 0.f; }
    }
    uint32_t orow = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int ck = item % p.chunks_per_col;
      int col = item / p.chunks_per_col;
      const int strip = col % p.strips;
      const int img = col / p.strips;
      const int y0 = ck * p.rows_per_chunk;
      const int rows = min(p.rows_per_chunk, p.h - y0);
      const int ox = strip * SW + r;
      const bool mvalid = ox < p.w;
      for (int t = 0; t < rows; ++t, ++orow) {
        const uint32_t acc = orow & 1;
        mbar_wait(sBar + 8 * (2 * NRU + acc), (orow >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int ph = 0; ph < 4; ++ph) {
          const size_t gm = ((size_t)img * (2 * p.h) + (2 * (y0 + t) + (ph >> 1))) * (2 * p.w) + (2 * ox + (ph & 1));
          const uint32_t trow = tmem_base + (acc * 4 + up_phase_pos(ph)) * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll
          for (int cb = 0; cb < BN; cb += 16) {
            float v[16];
            tmem_ld16(trow + cb, v);
            if constexpr (STATS) {
              if (mvalid) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  ss[cb + i] += v[i];
                  sq[cb + i] = fmaf(v[i], v[i], sq[cb + i]);
                }
              }
            }
            if (mvalid) strip_store16(v, p, n0 + cb, gm * p.cout + n0 + cb, vector_epilogue);
          }
        }
        tc_fence_before();
        mbar_arrive(sBar + 8 * (2 * NRU + 2 + acc));
      }
    }
    if constexpr (STATS) strip_flush_stats<BN>(ss, sq, red, q, lane, etid, n0, p);
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}


// =====================================================================================================================
// Input-stationary variant of the row-streaming 3x3 conv (single source).
//
// In the kernel above every output row issues 9 x CIN/16 MMAs of N = cout-tile, and each of them fetches its own
// 128-pixel A window from shared memory: for cout tiles of 16-64 that operand fetch, not the MMA pipe, is the binding
// unit (ncu: l1tex__data_pipe_tc_wavefronts_mem_shared ~50 % of peak at 19 % tensor-pipe activity).  Here the roles
// are turned around: an INPUT row i (window shifted by the horizontal tap s) is multiplied ONCE with the weights of
// the three vertical taps stacked along N ([W(r=2,s); W(r=1,s); W(r=0,s)], N = 3 x BN), and the three column blocks
// of D land in the accumulators of output rows i-1, i, i+1, which live in a ring of S TMEM slots at columns
// (o mod S) * BN.  3 x CIN/16 MMAs per row instead of 9 x CIN/16, one third of the A fetches.  Accumulators are always
// accumulated into: the epilogue zeroes a slot (tcgen05.st) after draining it, before handing it back.
// Chunk borders: input row j of a chunk only feeds the outputs t = j - r that lie inside the chunk, i.e. a sub-range of
// the stacked blocks; a run that would wrap around the slot ring is split into two MMAs.
constexpr int IS_SLOTS = 8;
constexpr int IS_RING = 4;

template <int BN, int CIN>
struct StripIsCfg {
  static constexpr int RB = CIN * 2;
  static constexpr int HALO = SW + 2;
  static constexpr int ROWBUF = ((HALO * RB + 1023) / 1024) * 1024;
  static constexpr int W_BLK = BN * RB;                    // one (r, s) weight block; multiples of the swizzle atom (8 rows)
  static constexpr int W_BYTES = ((9 * W_BLK + 1023) / 1024) * 1024;
  static constexpr int RED_BYTES = 4 * BN * 2 * 4;
  static constexpr int TMEM_COLS = IS_SLOTS * BN;          // 128 / 256 / 512
  static constexpr int SMEM = IS_RING * ROWBUF + W_BYTES + RED_BYTES + 1024 + 256;
  static constexpr uint32_t LAYOUT = RB == 128 ? 2u : (RB == 64 ? 4u : 6u);
};

__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int BN, int CIN, bool STATS>
__global__ void __launch_bounds__(NTHREADS)
conv_strip_is_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const StripP p) {
  typedef StripIsCfg<BN, CIN> C;
  constexpr int NRI = IS_RING, S = IS_SLOTS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sRing = base;
  const uint32_t sW = base + NRI * C::ROWBUF;
  const uint32_t sRed = sW + C::W_BYTES;
  const uint32_t sBar = sRed + C::RED_BYTES;   // full[NRI], empty[NRI], tfull[S], tempty[S], wbar
  const uint32_t bFull = sBar, bEmpty = sBar + 8 * NRI, bTfull = sBar + 8 * 2 * NRI, bTempty = bTfull + 8 * S;
  const uint32_t wbar = bTempty + 8 * S;
  const uint32_t sTmem = wbar + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (sTmem - base));
  float* red = reinterpret_cast<float*>(gen_base + (sRed - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * BN;

  if (tid == 0) {
    for (int s = 0; s < NRI; ++s) {
      mbar_init(bFull + 8 * s, 1);
      mbar_init(bEmpty + 8 * s, 1);
    }
    for (int a = 0; a < S; ++a) {
      mbar_init(bTfull + 8 * a, 1);
      mbar_init(bTempty + 8 * a, NEPI);
    }
    mbar_init(wbar, 1);
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
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
      // resident weights, stacked per horizontal tap s as [r = 2; r = 1; r = 0] blocks of [BN][CIN]
      mbar_expect_tx(wbar, (uint32_t)(9 * C::W_BLK));
      for (int s = 0; s < 3; ++s)
        for (int r = 0; r < 3; ++r)
          tma_load_2d(sW + (s * 3 + (2 - r)) * C::W_BLK, &map_w, wbar, (r * 3 + s) * p.cin, n0);
      uint32_t L = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int ck = item % p.chunks_per_col;
        int col = item / p.chunks_per_col;
        const int strip = col % p.strips;
        const int img = col / p.strips;
        const int y0 = ck * p.rows_per_chunk;
        const int rows = min(p.rows_per_chunk, p.h - y0);
        for (int j = 0; j < rows + 2; ++j, ++L) {
          const int s = L % NRI;
          if (L >= (uint32_t)NRI) mbar_wait(bEmpty + 8 * s, ((L / NRI) & 1) ^ 1);
          mbar_expect_tx(bFull + 8 * s, (uint32_t)(C::HALO * C::RB));
          tma_load_4d(sRing + s * C::ROWBUF, &map_x, bFull + 8 * s, 0, strip * SW - 1, y0 - 1 + j, img);
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================== MMA ISSUER
    constexpr uint32_t sbo = 8 * C::RB;
    mbar_wait(wbar, 0);
    uint32_t L = 0, O = 0;                       // running input-row / output-row counters of this CTA
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int ck = item % p.chunks_per_col;
      const int y0 = ck * p.rows_per_chunk;
      const int rows = min(p.rows_per_chunk, p.h - y0);
      for (int j = 0; j < rows + 2; ++j, ++L) {
        // input row j feeds outputs t = j - r, 0 <= t < rows
        const int r_lo = max(0, j - rows + 1), r_hi = min(2, j);
        if (j < rows) {                          // output t = j is touched for the first time: its slot must be free (and zeroed)
          const uint32_t o = O + j;
          mbar_wait(bTempty + 8 * (o % S), (o / S) & 1);
        }
        mbar_wait(bFull + 8 * (L % NRI), (L / NRI) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t rowbuf = sRing + (L % NRI) * C::ROWBUF;
          // outputs in ascending order: o_first = O + j - r_hi ... o_last = O + j - r_lo; stacked blocks 2 - r, ascending too
          const uint32_t o_first = O + j - r_hi;
          const int nblk = r_hi - r_lo + 1;
          const int slot0 = o_first % S;
          const int run0 = min(nblk, S - slot0);            // blocks before the slot ring wraps
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const uint32_t a0 = rowbuf + s * C::RB;
            const uint32_t b0 = sW + (s * 3 + (2 - r_hi)) * C::W_BLK;
#pragma unroll
            for (int k = 0; k < CIN / 16; ++k) {
              umma_f16(tmem_base + slot0 * BN, strip_desc(a0 + k * 32, sbo, C::LAYOUT, 0), umma_desc(b0 + k * 32, 16, sbo, C::LAYOUT),
                       umma_idesc(run0 * BN), 1u);
              if (run0 < nblk)
                umma_f16(tmem_base, strip_desc(a0 + k * 32, sbo, C::LAYOUT, 0),
                         umma_desc(b0 + run0 * C::W_BLK + k * 32, 16, sbo, C::LAYOUT), umma_idesc((nblk - run0) * BN), 1u);
            }
          }
          umma_commit(bEmpty + 8 * (L % NRI));                       // this input row has been consumed
          if (j >= 2) umma_commit(bTfull + 8 * ((O + j - 2) % S));   // output t = j - 2 received its last contribution
        }
        __syncwarp();
      }
      O += rows;
    }
    tc_fence_before();
  } else {
    // =========================================================== EPILOGUE (warps 2..5)
    const int q = warp & 3;
    const int r = q * 32 + lane;               // pixel within the strip == TMEM lane
    const bool vector_epilogue = (p.cout % 16 == 0) && !p.dst_f32 && p.act != RCFD_ACT_DEPTH_HEAD;
    const int etid = tid - 64;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float ss[STATS ? BN : 1], sq[STATS ? BN : 1];
    if constexpr (STATS) {
#pragma unroll
      for (int i = 0; i < BN; ++i) { ss[i] = 0.f; sq[i] = 0.f; }
    }
    // all slots start zeroed and free
#pragma unroll 1
    for (int a = 0; a < S; ++a) {
#pragma unroll
      for (int cb = 0; cb < BN; cb += 16) tmem_st16_zero(lane_base + a * BN + cb);
    }
    tmem_wait_st();
    tc_fence_before();
    for (int a = 0; a < S; ++a) mbar_arrive(bTempty + 8 * a);
    uint32_t O = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int ck = item % p.chunks_per_col;
      int col = item / p.chunks_per_col;
      const int strip = col % p.strips;
      const int img = col / p.strips;
      const int y0 = ck * p.rows_per_chunk;
      const int rows = min(p.rows_per_chunk, p.h - y0);
      const int ox = strip * SW + r;
      const bool mvalid = ox < p.w;
      for (int t = 0; t < rows; ++t, ++O) {
        const uint32_t slot = O % S;
        const size_t gm = ((size_t)img * p.h + (y0 + t)) * p.w + ox;
        mbar_wait(bTfull + 8 * slot, (O / S) & 1);
        tc_fence_after();
        const uint32_t trow = lane_base + slot * BN;
        if constexpr (STATS) {
#pragma unroll
          for (int cb = 0; cb < BN; cb += 16) {
            float v[16];
            tmem_ld16(trow + cb, v);
            tmem_st16_zero(trow + cb);
            if (mvalid) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                ss[cb + i] += v[i];
                sq[cb + i] = fmaf(v[i], v[i], sq[cb + i]);
              }
              strip_store16(v, p, n0 + cb, gm * p.cout + n0 + cb, vector_epilogue);
            }
          }
        } else {
#pragma unroll 1
          for (int cb = 0; cb < BN; cb += 16) {
            float v[16];
            tmem_ld16(trow + cb, v);
            tmem_st16_zero(trow + cb);
            if (mvalid) strip_store16(v, p, n0 + cb, gm * p.cout + n0 + cb, vector_epilogue);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bTempty + 8 * slot);
      }
    }
    if constexpr (STATS) strip_flush_stats<BN>(ss, sq, red, q, lane, etid, n0, p);
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}


// =====================================================================================================================
// Input-stationary variant of the sub-pixel up-conv (3x3 behind an exact 2x nearest up-sampling).
//
// Low-res input row j feeds the low-res output rows t = j - r, r = a + t2 in {0, 1, 2} (phase row a, vertical tap t2):
// r = 2 -> (a=1, t2=1), r = 1 -> (a=0, t2=1) and (a=1, t2=0), r = 0 -> (a=0, t2=0).  For a fixed phase column b and
// horizontal tap u (input window shifted by c = b + u pixels) those four (a, t2) weight blocks are stacked along N in the
// order of the accumulators they feed -- [t=j-2: a=1 | t=j-1: a=0 | t=j-1: a=1 | t=j: a=0] -- and issued as ONE MMA
// (N = 4 x BN): 4 MMAs per input row and 16-channel chunk (c = 0: b=0; c = 1: b=0 and b=1; c = 2: b=1) instead of 10.
// TMEM: two rings (one per phase column b) of S slots, a slot = [a=0 | a=1] x BN columns of one output row.
constexpr int UIS_SLOTS = 4;

template <int BN, int CIN>
struct StripUpIsCfg {
  static constexpr int RB = CIN * 2;
  static constexpr int HALO = SW + 2;
  static constexpr int ROWBUF = ((HALO * RB + 1023) / 1024) * 1024;
  static constexpr int W_BLK = BN * RB;
  static constexpr int W_BYTES = ((16 * W_BLK + 1023) / 1024) * 1024;
  static constexpr int RED_BYTES = 4 * BN * 2 * 4;
  static constexpr int TMEM_COLS = UIS_SLOTS * 4 * BN;     // 2 rings x S slots x 2 phases rows x BN
  static constexpr int SMEM = IS_RING * ROWBUF + W_BYTES + RED_BYTES + 1024 + 256;
  static constexpr uint32_t LAYOUT = RB == 128 ? 2u : (RB == 64 ? 4u : 6u);
};

template <int BN, int CIN, bool STATS>
__global__ void __launch_bounds__(NTHREADS)
conv_strip_up_is_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const StripP p) {
  typedef StripUpIsCfg<BN, CIN> C;
  constexpr int NRI = IS_RING, S = UIS_SLOTS;
  static_assert(C::TMEM_COLS <= 512, "sub-pixel input-stationary kernel: cout tile too wide for TMEM");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sRing = base;
  const uint32_t sW = base + NRI * C::ROWBUF;
  const uint32_t sRed = sW + C::W_BYTES;
  const uint32_t sBar = sRed + C::RED_BYTES;
  const uint32_t bFull = sBar, bEmpty = sBar + 8 * NRI, bTfull = sBar + 8 * 2 * NRI, bTempty = bTfull + 8 * S;
  const uint32_t wbar = bTempty + 8 * S;
  const uint32_t sTmem = wbar + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (sTmem - base));
  float* red = reinterpret_cast<float*>(gen_base + (sRed - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * BN;

  if (tid == 0) {
    for (int s = 0; s < NRI; ++s) {
      mbar_init(bFull + 8 * s, 1);
      mbar_init(bEmpty + 8 * s, 1);
    }
    for (int a = 0; a < S; ++a) {
      mbar_init(bTfull + 8 * a, 1);
      mbar_init(bTempty + 8 * a, NEPI);
    }
    mbar_init(wbar, 1);
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
  constexpr uint32_t RING1 = S * 2 * BN;            // first column of the b = 1 ring

  if (warp == 0) {
    // =========================================================== TMA PRODUCER
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
      // weights (rcfd_pack_upconv2x_weight: [phase = a*2+b][cout][tap = t2*2+u][cin]); group g = b*2+u holds the four
      // (a, t2) blocks in accumulator order (a,t2) = (1,1) (0,1) (1,0) (0,0)
      mbar_expect_tx(wbar, (uint32_t)(16 * C::W_BLK));
      for (int g = 0; g < 4; ++g) {
        const int b = g >> 1, u = g & 1;
        for (int blk = 0; blk < 4; ++blk) {
          const int a = (blk == 0 || blk == 2) ? 1 : 0, t2 = blk < 2 ? 1 : 0;
          tma_load_2d(sW + (g * 4 + blk) * C::W_BLK, &map_w, wbar, (t2 * 2 + u) * p.cin, (a * 2 + b) * p.cout + n0);
        }
      }
      uint32_t L = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int ck = item % p.chunks_per_col;
        int col = item / p.chunks_per_col;
        const int strip = col % p.strips;
        const int img = col / p.strips;
        const int y0 = ck * p.rows_per_chunk;
        const int rows = min(p.rows_per_chunk, p.h - y0);
        for (int j = 0; j < rows + 2; ++j, ++L) {
          const int s = L % NRI;
          if (L >= (uint32_t)NRI) mbar_wait(bEmpty + 8 * s, ((L / NRI) & 1) ^ 1);
          mbar_expect_tx(bFull + 8 * s, (uint32_t)(C::HALO * C::RB));
          tma_load_4d(sRing + s * C::ROWBUF, &map_x, bFull + 8 * s, 0, strip * SW - 1, y0 - 1 + j, img);
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================== MMA ISSUER
    constexpr uint32_t sbo = 8 * C::RB;
    mbar_wait(wbar, 0);
    uint32_t L = 0, O = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int ck = item % p.chunks_per_col;
      const int y0 = ck * p.rows_per_chunk;
      const int rows = min(p.rows_per_chunk, p.h - y0);
      for (int j = 0; j < rows + 2; ++j, ++L) {
        if (j < rows) {                          // output row t = j is touched for the first time
          const uint32_t o = O + j;
          mbar_wait(bTempty + 8 * (o % S), (o / S) & 1);
        }
        mbar_wait(bFull + 8 * (L % NRI), (L / NRI) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t rowbuf = sRing + (L % NRI) * C::ROWBUF;
          // blocks 0..3 <-> (t, a) = (j-2, 1) (j-1, 0) (j-1, 1) (j, 0); the valid ones are a contiguous range
          const int b_lo = j >= 2 ? 0 : (j >= 1 ? 1 : 3);
          const int b_hi = j < rows ? 3 : (j - 1 < rows ? 2 : 0);
          // column (within a ring) of block blk
          uint32_t colv[4];
#pragma unroll
          for (int blk = 0; blk < 4; ++blk) {
            const int t = blk == 0 ? j - 2 : (blk == 3 ? j : j - 1);
            const int a = (blk == 0 || blk == 2) ? 1 : 0;
            colv[blk] = (uint32_t)(((O + t) % S) * 2 * BN + a * BN);      // only used for valid blocks
          }
          // first run: from b_lo while columns stay contiguous; a second run after the ring wrap
          int run0 = 1;
          while (b_lo + run0 <= b_hi && colv[b_lo + run0] == colv[b_lo + run0 - 1] + BN) ++run0;
          const int nblk = b_hi - b_lo + 1;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int b = g >> 1, u = g & 1;
            const uint32_t a0 = rowbuf + (b + u) * C::RB;
            const uint32_t w0 = sW + (g * 4 + b_lo) * C::W_BLK;
            const uint32_t d0 = tmem_base + (b ? RING1 : 0u);
#pragma unroll
            for (int k = 0; k < CIN / 16; ++k) {
              umma_f16(d0 + colv[b_lo], umma_desc(a0 + k * 32, 16, sbo, C::LAYOUT), umma_desc(w0 + k * 32, 16, sbo, C::LAYOUT),
                       umma_idesc(run0 * BN), 1u);
              if (run0 < nblk)
                umma_f16(d0 + colv[b_lo + run0], umma_desc(a0 + k * 32, 16, sbo, C::LAYOUT),
                         umma_desc(w0 + run0 * C::W_BLK + k * 32, 16, sbo, C::LAYOUT), umma_idesc((nblk - run0) * BN), 1u);
            }
          }
          umma_commit(bEmpty + 8 * (L % NRI));
          if (j >= 2) umma_commit(bTfull + 8 * ((O + j - 2) % S));
        }
        __syncwarp();
      }
      O += rows;
    }
    tc_fence_before();
  } else {
    // =========================================================== EPILOGUE (warps 2..5)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const bool vector_epilogue = (p.cout % 16 == 0) && !p.dst_f32 && p.act != RCFD_ACT_DEPTH_HEAD;
    const int etid = tid - 64;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float ss[STATS ? BN : 1], sq[STATS ? BN : 1];
    if constexpr (STATS) {
#pragma unroll
      for (int i = 0; i < BN; ++i) { ss[i] = 0.f; sq[i] = 0.f; }
    }
#pragma unroll 1
    for (int cb = 0; cb < C::TMEM_COLS; cb += 16) tmem_st16_zero(lane_base + cb);
    tmem_wait_st();
    tc_fence_before();
    for (int a = 0; a < S; ++a) mbar_arrive(bTempty + 8 * a);
    uint32_t O = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int ck = item % p.chunks_per_col;
      int col = item / p.chunks_per_col;
      const int strip = col % p.strips;
      const int img = col / p.strips;
      const int y0 = ck * p.rows_per_chunk;
      const int rows = min(p.rows_per_chunk, p.h - y0);
      const int ox = strip * SW + r;
      const bool mvalid = ox < p.w;
      for (int t = 0; t < rows; ++t, ++O) {
        const uint32_t slot = O % S;
        mbar_wait(bTfull + 8 * slot, (O / S) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int ph = 0; ph < 4; ++ph) {
          const int a = ph >> 1, b = ph & 1;
          const size_t gm = ((size_t)img * (2 * p.h) + (2 * (y0 + t) + a)) * (2 * p.w) + (2 * ox + b);
          const uint32_t trow = lane_base + (b ? RING1 : 0u) + slot * 2 * BN + a * BN;
#pragma unroll
          for (int cb = 0; cb < BN; cb += 16) {
            float v[16];
            tmem_ld16(trow + cb, v);
            tmem_st16_zero(trow + cb);
            if (mvalid) {
              if constexpr (STATS) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  ss[cb + i] += v[i];
                  sq[cb + i] = fmaf(v[i], v[i], sq[cb + i]);
                }
              }
              strip_store16(v, p, n0 + cb, gm * p.cout + n0 + cb, vector_epilogue);
            }
          }
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bTempty + 8 * slot);
      }
    }
    if constexpr (STATS) strip_flush_stats<BN>(ss, sq, red, q, lane, etid, n0, p);
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

inline bool make_row_map(CUtensorMap* m, const void* ptr, int n, int h, int w, int c, int halo = HALO_W) {
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  cuuint32_t box[4] = {(cuuint32_t)c, (cuuint32_t)halo, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(c), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int CIN, bool STATS, int KS, int CIN1>
int launch_strip_s(const ConvKP& k, StripP& t, cudaStream_t st) {
  note_kernel("conv_strip_kernel<%d,%d,%d,%d,%d>", BN, CIN, (int)STATS, KS, CIN1);
  typedef StripCfg<BN, CIN, KS, CIN1> C;
  static_assert(C::SMEM <= 227 * 1024, "row-streaming configuration exceeds shared memory");
  static int per_sm_dev[16] = {}; int& per_sm = per_sm_dev[cur_dev()];                 // resident CTAs per SM (shared memory AND registers: the statistics variant is wide)
  if (per_sm == 0) {
    cudaError_t e = cudaFuncSetAttribute(conv_strip_kernel<BN, CIN, STATS, KS, CIN1>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) { set_error("conv_strip: smem attribute: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
    // two CTAs per SM when shared memory (227 KB) and the register file (64 K) both allow it
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, conv_strip_kernel<BN, CIN, STATS, KS, CIN1>);
    if (e != cudaSuccess) { set_error("conv_strip: attributes: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
    const int regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * NTHREADS;
    per_sm = (C::SMEM <= 112 * 1024 && 2 * regs_per_cta <= 65536) ? 2 : 1;
  }
  alignas(64) CUtensorMap mx, mw, mx1, mw1;
  if (!make_row_map(&mx, k.src0, k.n, k.hin, k.win, k.c0, C::HALO) || !make_w_map(&mw, k.weight, k.cout, k.K, CIN, BN)) {
    set_error("conv_strip: cuTensorMapEncodeTiled failed");
    return RCFD_ECUDA;
  }
  if (CIN1 > 0) {
    if (!make_row_map(&mx1, k.src1, k.n, k.hin, k.win, k.c1, C::HALO) || !make_w_map(&mw1, k.weight, k.cout, k.K, CIN1, BN)) {
      set_error("conv_strip: cuTensorMapEncodeTiled failed (second source)");
      return RCFD_ECUDA;
    }
  } else {
    mx1 = mx;
    mw1 = mw;
  }
  const int ntile = ceil_div(k.cout, BN);
  int ctas = num_sms() * per_sm / ntile;                   // two CTAs per SM hide each other's barrier latencies
  if (ctas < 1) ctas = 1;
  plan_row_chunks(t.h, t.n * t.strips, ctas, 6, 4, &t.rows_per_chunk, &t.chunks_per_col);
  t.num_items = t.n * t.strips * t.chunks_per_col;
  if (ctas > t.num_items) ctas = t.num_items;
  dim3 grid(ctas, ntile);
  conv_strip_kernel<BN, CIN, STATS, KS, CIN1><<<grid, NTHREADS, C::SMEM, st>>>(mx, mw, mx1, mw1, t);
  RCFD_CHECK_LAUNCH("conv_strip");
  return RCFD_OK;
}

template <int BN, int CIN, int KS = 3, int CIN1 = 0>
int launch_strip(const ConvKP& k, StripP& t, cudaStream_t st) {
  return t.ssum != nullptr ? launch_strip_s<BN, CIN, true, KS, CIN1>(k, t, st) : launch_strip_s<BN, CIN, false, KS, CIN1>(k, t, st);
}


template <int BN, int CIN, bool STATS>
int launch_strip_is_s(const ConvKP& k, StripP& t, cudaStream_t st) {
  note_kernel("conv_strip_is_kernel<%d,%d,%d>", BN, CIN, (int)STATS);
  typedef StripIsCfg<BN, CIN> C;
  static_assert(C::SMEM <= 227 * 1024, "row-streaming configuration exceeds shared memory");
  static int per_sm_dev[16] = {}; int& per_sm = per_sm_dev[cur_dev()];
  if (per_sm == 0) {
    cudaError_t e = cudaFuncSetAttribute(conv_strip_is_kernel<BN, CIN, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) { set_error("conv_strip_is: smem attribute: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, conv_strip_is_kernel<BN, CIN, STATS>);
    if (e != cudaSuccess) { set_error("conv_strip_is: attributes: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
    const int regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * NTHREADS;
    per_sm = (C::SMEM <= 112 * 1024 && 2 * regs_per_cta <= 65536 && 2 * C::TMEM_COLS <= 512) ? 2 : 1;
  }
  alignas(64) CUtensorMap mx, mw;
  if (!make_row_map(&mx, k.src0, k.n, k.hin, k.win, k.c0, C::HALO) || !make_w_map(&mw, k.weight, k.cout, k.K, CIN, BN)) {
    set_error("conv_strip_is: cuTensorMapEncodeTiled failed");
    return RCFD_ECUDA;
  }
  const int ntile = ceil_div(k.cout, BN);
  int ctas = num_sms() * per_sm / ntile;
  if (ctas < 1) ctas = 1;
  plan_row_chunks(t.h, t.n * t.strips, ctas, 6, 4, &t.rows_per_chunk, &t.chunks_per_col);
  t.num_items = t.n * t.strips * t.chunks_per_col;
  if (ctas > t.num_items) ctas = t.num_items;
  dim3 grid(ctas, ntile);
  conv_strip_is_kernel<BN, CIN, STATS><<<grid, NTHREADS, C::SMEM, st>>>(mx, mw, t);
  RCFD_CHECK_LAUNCH("conv_strip_is");
  return RCFD_OK;
}

template <int BN, int CIN>
int launch_strip_is(const ConvKP& k, StripP& t, cudaStream_t st) {
  return t.ssum != nullptr ? launch_strip_is_s<BN, CIN, true>(k, t, st) : launch_strip_is_s<BN, CIN, false>(k, t, st);
}


template <int BN, int CIN, bool STATS>
int launch_strip_up_is_s(const ConvKP& k, StripP& t, cudaStream_t st) {
  note_kernel("conv_strip_up_is_kernel<%d,%d,%d>", BN, CIN, (int)STATS);
  typedef StripUpIsCfg<BN, CIN> C;
  static_assert(C::SMEM <= 227 * 1024, "row-streaming configuration exceeds shared memory");
  static int per_sm_dev[16] = {}; int& per_sm = per_sm_dev[cur_dev()];
  if (per_sm == 0) {
    cudaError_t e = cudaFuncSetAttribute(conv_strip_up_is_kernel<BN, CIN, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) { set_error("conv_strip_up_is: smem attribute: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, conv_strip_up_is_kernel<BN, CIN, STATS>);
    if (e != cudaSuccess) { set_error("conv_strip_up_is: attributes: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
    const int regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * NTHREADS;
    per_sm = (C::SMEM <= 112 * 1024 && 2 * regs_per_cta <= 65536 && 2 * C::TMEM_COLS <= 512) ? 2 : 1;
  }
  alignas(64) CUtensorMap mx, mw;
  if (!make_row_map(&mx, k.src0, k.n, k.h0, k.w0, k.c0) || !make_w_map(&mw, k.weight_up2x, 4 * k.cout, 4 * k.c0, CIN, BN)) {
    set_error("conv_strip_up_is: cuTensorMapEncodeTiled failed");
    return RCFD_ECUDA;
  }
  const int ntile = ceil_div(k.cout, BN);
  int ctas = num_sms() * per_sm / ntile;
  if (ctas < 1) ctas = 1;
  plan_row_chunks(t.h, t.n * t.strips, ctas, 6, 4, &t.rows_per_chunk, &t.chunks_per_col);
  t.num_items = t.n * t.strips * t.chunks_per_col;
  if (ctas > t.num_items) ctas = t.num_items;
  dim3 grid(ctas, ntile);
  conv_strip_up_is_kernel<BN, CIN, STATS><<<grid, NTHREADS, C::SMEM, st>>>(mx, mw, t);
  RCFD_CHECK_LAUNCH("conv_strip_up_is");
  return RCFD_OK;
}

template <int BN, int CIN>
int launch_strip_up_is(const ConvKP& k, StripP& t, cudaStream_t st) {
  return t.ssum != nullptr ? launch_strip_up_is_s<BN, CIN, true>(k, t, st) : launch_strip_up_is_s<BN, CIN, false>(k, t, st);
}

template <int BN, int CIN, bool STATS>
int launch_strip_up_s(const ConvKP& k, StripP& t, cudaStream_t st) {
  note_kernel("conv_strip_up_kernel<%d,%d,%d>", BN, CIN, (int)STATS);
  typedef StripUpCfg<BN, CIN> C;
  static bool attr_set_dev[16] = {}; bool& attr_set = attr_set_dev[cur_dev()];   // per device: the attribute belongs to the device's copy of the kernel
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_strip_up_kernel<BN, CIN, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) { set_error("conv_strip_up: smem attribute: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
    attr_set = true;
  }
  alignas(64) CUtensorMap mx, mw;
  if (!make_row_map(&mx, k.src0, k.n, k.h0, k.w0, k.c0) || !make_w_map(&mw, k.weight_up2x, 4 * k.cout, 4 * k.c0, CIN, BN)) {
    set_error("conv_strip_up: cuTensorMapEncodeTiled failed");
    return RCFD_ECUDA;
  }
  const int ntile = ceil_div(k.cout, BN);
  int ctas = num_sms() / ntile;
  if (ctas < 1) ctas = 1;
  if (ctas > t.num_items) ctas = t.num_items;
  dim3 grid(ctas, ntile);
  conv_strip_up_kernel<BN, CIN, STATS><<<grid, NTHREADS, C::SMEM, st>>>(mx, mw, t);
  RCFD_CHECK_LAUNCH("conv_strip_up");
  return RCFD_OK;
}

template <int BN, int CIN>
int launch_strip_up(const ConvKP& k, StripP& t, cudaStream_t st) {
  return t.ssum != nullptr ? launch_strip_up_s<BN, CIN, true>(k, t, st) : launch_strip_up_s<BN, CIN, false>(k, t, st);
}

}  // namespace

int g_strip_desc_mode = 0;
int g_strip_input_stationary = 1;  // rcfd_set_option("strip_input_stationary"): 0 = output-stationary kernel for every 3x3 conv
int g_strip_max_waste = 50;       // rcfd_set_option("strip_max_waste"): padded strip columns tolerated, percent of the map width
int g_strip_up_max_waste = 30;    // same for the up-sampling variant (measured on the low-res width)

bool conv_strip_up_supported(const ConvKP& p, int dtype) {
  if (dtype != RCFD_BF16 || !p.up || p.weight_up2x == nullptr || p.dil != 1 || p.c1 != 0) return false;
  if (p.kh != 3 || p.kw != 3 || p.stride != 1 || p.pad != 1) return false;
  if (p.hin != 2 * p.h0 || p.win != 2 * p.w0 || p.ho != p.hin || p.wo != p.win) return false;
  if (p.c0 != 32 && p.c0 != 64) return false;
  if (p.cout % 16 != 0 || p.act == RCFD_ACT_DEPTH_HEAD) return false;      // dst_f32: scalar-store epilogue (parity mode)
  if ((reinterpret_cast<uintptr_t>(p.src0) & 15) || (reinterpret_cast<uintptr_t>(p.weight_up2x) & 15)) return false;
  return get_encode() != nullptr;
}

bool conv_strip_up_preferred(const ConvKP& p, int dtype) {
  if (!conv_strip_up_supported(p, dtype)) return false;
  const int strips = ceil_div(p.w0, SW);
  // many rows (the RadarNet decoder: 64 point crops per image) amortise the padded columns better than the per-tap
  // TMA engine does: measured 49.4 -> 45.2 ms per 16-image batch with up to 80 % padding tolerated there
  const int waste = (long)p.n * p.h0 >= 4096 ? (g_strip_up_max_waste > 80 ? g_strip_up_max_waste : 80) : g_strip_up_max_waste;
  return p.h0 >= 32 && (long)strips * SW * 100 <= (long)p.w0 * (100 + waste);
}

int conv_strip_up_launch(const ConvKP& p, cudaStream_t st) {
  StripP t;
  t.n = p.n; t.h = p.h0; t.w = p.w0; t.cin = p.c0; t.cout = p.cout;     // the kernel walks the LOW-RES grid
  t.strips = ceil_div(p.w0, SW);
  const int bn = p.cout % 64 == 0 ? 64 : (p.cout % 32 == 0 ? 32 : 16);
  const int ntile = ceil_div(p.cout, bn);
  const int ctas = num_sms() / ntile > 0 ? num_sms() / ntile : 1;
  const int cols = p.n * t.strips;
  plan_row_chunks(p.h0, cols, ctas, 6, 4, &t.rows_per_chunk, &t.chunks_per_col);
  t.num_items = cols * t.chunks_per_col;
  t.desc_mode = 0;
  t.dst = p.dst; t.scale = p.scale; t.shift = p.shift; t.act = p.act; t.p0 = p.p0; t.p1 = p.p1;
  t.residual = p.residual; t.ssum = p.ssum; t.ssq = p.ssq; t.accumulate = p.accumulate; t.dst_f32 = p.dst_f32;
  // input-stationary variant (4 stacked MMAs per input row and chunk): measured faster only for 16-wide cout tiles (RadarNet's
  // 352x288 up-conv); with 32 the window-merged kernel below wins (65.7 vs 85.6 us on deconv0.deconv).  2 = force, 1 = heuristic.
  if ((g_strip_input_stationary == 2 && bn <= 32) || (g_strip_input_stationary == 1 && bn == 16)) {
    if (p.c0 == 64) return bn == 32 ? launch_strip_up_is<32, 64>(p, t, st) : launch_strip_up_is<16, 64>(p, t, st);
    return bn == 32 ? launch_strip_up_is<32, 32>(p, t, st) : launch_strip_up_is<16, 32>(p, t, st);
  }
  if (p.c0 == 64) {
    switch (bn) {
      case 64: return launch_strip_up<64, 64>(p, t, st);
      case 32: return launch_strip_up<32, 64>(p, t, st);
      default: return launch_strip_up<16, 64>(p, t, st);
    }
  }
  switch (bn) {
    case 64: return launch_strip_up<64, 32>(p, t, st);
    case 32: return launch_strip_up<32, 32>(p, t, st);
    default: return launch_strip_up<16, 32>(p, t, st);
  }
}

// concat pairs (c0 | c1) the kernel is instantiated for, with the cout tile that fits shared memory
static int strip_dual_bn(const ConvKP& p) {
  if (p.kh != 3) return 0;
  if (p.c0 == 64 && p.c1 == 32) return p.cout % 64 == 0 ? 64 : (p.cout % 32 == 0 ? 32 : 0);
  if (p.c0 == 64 && p.c1 == 64) return p.cout % 32 == 0 ? 32 : 0;
  if (p.c0 == 32 && p.c1 == 32) return p.cout % 64 == 0 ? 64 : (p.cout % 32 == 0 ? 32 : 0);
  return 0;
}

bool conv_strip_supported(const ConvKP& p, int dtype) {
  if (dtype != RCFD_BF16 || p.up || p.dil != 1) return false;
  if (p.c1 != 0) {
    if (strip_dual_bn(p) == 0 || p.kw != 3 || p.stride != 1 || p.pad != 1 || p.ho != p.hin || p.wo != p.win) return false;
    if ((reinterpret_cast<uintptr_t>(p.src0) & 15) || (reinterpret_cast<uintptr_t>(p.src1) & 15) ||
        (reinterpret_cast<uintptr_t>(p.weight) & 15))
      return false;
    return get_encode() != nullptr;
  }
  const bool stem = p.kh == 4 && p.kw == 4 && p.stride == 1 && p.pad == 2 && p.c0 == 16 && (p.cout == 16 || p.cout == 32);
  if (!stem && (p.kh != 3 || p.kw != 3 || p.stride != 1 || p.pad != 1)) return false;
  if (p.c0 != 16 && p.c0 != 32 && p.c0 != 64) return false;
  if (p.ho != p.hin || p.wo != p.win) return false;
  if ((reinterpret_cast<uintptr_t>(p.src0) & 15) || (reinterpret_cast<uintptr_t>(p.weight) & 15)) return false;
  return get_encode() != nullptr;
}

// worth it only where the 128-wide strips fit the image width reasonably and there is enough height
bool conv_strip_preferred(const ConvKP& p, int dtype) {
  if (!conv_strip_supported(p, dtype)) return false;
  if (p.c0 == 64 && p.c1 == 64) return false;      // cout tile 32 (two passes over the input): measured slower than the per-tap engine
  const int strips = ceil_div(p.wo, SW);
  // default <= 50 % padded columns (176-wide maps: 2 strips); many rows (RadarNet: 64 crops per image) tolerate 80 %
  const int waste = (long)p.n * p.ho >= 4096 ? (g_strip_max_waste > 80 ? g_strip_max_waste : 80) : g_strip_max_waste;
  return p.ho >= 64 && (long)strips * SW * 100 <= (long)p.wo * (100 + waste);
}

int conv_strip_launch(const ConvKP& p, cudaStream_t st) {
  StripP t;
  t.n = p.n; t.h = p.ho; t.w = p.wo; t.cin = p.c0 + p.c1; t.cout = p.cout;
  t.strips = ceil_div(p.wo, SW);
  const int bn = p.c1 != 0 ? strip_dual_bn(p) : (p.cout % 64 == 0 ? 64 : (p.cout % 32 == 0 ? 32 : 16));
  const int ntile = ceil_div(p.cout, bn);
  (void)ntile;                      // the row chunks are planned in launch_strip_s, once the resident CTA count is known
  t.desc_mode = g_strip_desc_mode;
  t.dst = p.dst; t.scale = p.scale; t.shift = p.shift; t.act = p.act; t.p0 = p.p0; t.p1 = p.p1;
  t.residual = p.residual; t.ssum = p.ssum; t.ssq = p.ssq; t.accumulate = p.accumulate; t.dst_f32 = p.dst_f32;
  if (p.c1 != 0) {                  // conv over torch.cat([x, skip], 1): both sources streamed
    if (p.c0 == 64 && p.c1 == 32) return bn == 64 ? launch_strip<64, 64, 3, 32>(p, t, st) : launch_strip<32, 64, 3, 32>(p, t, st);
    if (p.c0 == 64 && p.c1 == 64) return launch_strip<32, 64, 3, 64>(p, t, st);
    return bn == 64 ? launch_strip<64, 32, 3, 32>(p, t, st) : launch_strip<32, 32, 3, 32>(p, t, st);
  }
  if (g_strip_input_stationary && p.kh == 3 && p.c1 == 0) {       // one A fetch per horizontal tap, vertical taps stacked along N
    if (p.c0 == 64) return bn == 64 ? launch_strip_is<64, 64>(p, t, st) : (bn == 32 ? launch_strip_is<32, 64>(p, t, st) : launch_strip_is<16, 64>(p, t, st));
    if (p.c0 == 32) return bn == 64 ? launch_strip_is<64, 32>(p, t, st) : (bn == 32 ? launch_strip_is<32, 32>(p, t, st) : launch_strip_is<16, 32>(p, t, st));
    if (p.c0 == 16) return bn == 64 ? launch_strip_is<64, 16>(p, t, st) : (bn == 32 ? launch_strip_is<32, 16>(p, t, st) : launch_strip_is<16, 16>(p, t, st));
  }
  if (p.c0 == 64) {
    switch (bn) {
      case 64: return launch_strip<64, 64>(p, t, st);
      case 32: return launch_strip<32, 64>(p, t, st);
      default: return launch_strip<16, 64>(p, t, st);
    }
  }
  if (p.kh == 4) {                  // 7x7 / stride-2 stems as 4x4 windows over the space-to-depth input
    return bn == 32 ? launch_strip<32, 16, 4>(p, t, st) : launch_strip<16, 16, 4>(p, t, st);
  }
  if (p.c0 == 16) {                 // d(logit) of the 1-channel head, stored with 16 channels
    switch (bn) {
      case 64: return launch_strip<64, 16>(p, t, st);
      case 32: return launch_strip<32, 16>(p, t, st);
      default: return launch_strip<16, 16>(p, t, st);
    }
  }
  switch (bn) {
    case 64: return launch_strip<64, 32>(p, t, st);
    case 32: return launch_strip<32, 32>(p, t, st);
    default: return launch_strip<16, 32>(p, t, st);
  }
}

}  // namespace rcfd

// Host-only helper exported for tests and tools: how the row-streaming kernels cut `h` rows of `cols` column strips into
// chunks for `ctas` persistent CTAs (tma_common.cuh plan_row_chunks).  No device work.
extern "C" int rcfd_plan_row_chunks(int32_t h, int32_t cols, int32_t ctas, int32_t overhead_rows, int32_t min_rows,
                                    int32_t* rows_per_chunk, int32_t* chunks_per_col) {
  RCFD_CHECK_ARG(h > 0 && cols > 0 && ctas > 0 && overhead_rows >= 0 && min_rows > 0 && rows_per_chunk && chunks_per_col,
                 "plan_row_chunks: bad args");
  int rpc = 0, cpc = 0;
  rcfd::tma::plan_row_chunks(h, cols, ctas, overhead_rows, min_rows, &rpc, &cpc);
  *rows_per_chunk = rpc;
  *chunks_per_col = cpc;
  return RCFD_OK;
}

