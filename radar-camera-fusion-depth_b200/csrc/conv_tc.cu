// tcgen05 engine of the implicit-GEMM convolution (bf16 operands, fp32 accumulation in TMEM).
//
// One CTA computes a 128-pixel x BN-channel output tile:
//   * warps 0-3 (128 threads) are PRODUCERS: thread r owns output pixel m0+r and gathers its
//     K = (tap, channel) row in 16-byte cp.async chunks straight into the canonical
//     K-major SWIZZLE_128B shared-memory layout the tensor core reads -- the filter taps,
//     zero padding, stride, nearest up-sampling, channel concat and zero insertion (dgrad of
//     stride-2 convs) are all address arithmetic here; nothing is materialised in HBM.
//     The same threads also stream the weight tile (B operand).
//   * warp 4 is the MMA ISSUER: one elected lane issues tcgen05.mma (M=128, N=BN, K=16) on
//     shared-memory descriptors, accumulating in tensor memory; tcgen05.commit releases the
//     smem stage back to the producers through an mbarrier.
//   * after the K loop warps 0-3 become the EPILOGUE: tcgen05.ld their 32 TMEM lanes, then
//     BatchNorm statistics (butterfly transpose-reduce) / folded BN + activation (+ residual)
//     and 16-byte bf16 stores.
// smem stages form an mbarrier ring (full: 128 producer arrivals after
// cp.async.wait_group + fence.proxy.async; empty: tcgen05.commit).
#include "tc_common.cuh"

namespace rcfd {
namespace {

using namespace tc;
constexpr int BKE = 64;          // K elements per stage (128 bytes of bf16 = one swizzle row)
constexpr int NPROD = 128;
constexpr int NTHREADS = 160;

template <int BN>
struct TcCfg {
  static constexpr int STAGES = BN >= 128 ? 3 : 4;
  static constexpr int A_BYTES = TM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS) conv_tc_kernel(const ConvKP p) {
  typedef TcCfg<BN> C;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + C::STAGES * C::A_BYTES;
  const uint32_t sBar = sB + C::STAGES * C::B_BYTES;          // full[STAGES], empty[STAGES], accum
  const uint32_t sTmem = sBar + 8 * (2 * C::STAGES + 1);
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (sTmem - base));
  float* red = reinterpret_cast<float*>(gen_base);             // epilogue scratch aliases stage 0 (drained by then)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * BN;
  const int num_kb = (p.K + BKE - 1) / BKE;

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(sBar + 8 * s, NPROD);
      mbar_init(sBar + 8 * (C::STAGES + s), 1);
    }
    mbar_init(sBar + 8 * (2 * C::STAGES), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmem), "n"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // =========================================================== PRODUCER
    // Thread t serves rows (t >> 1) and (t >> 1) + 64 and the 16-byte chunks j = (t & 1) + 2 i of
    // each 128-byte K row: an even/odd lane pair fetches whole 32-byte sectors, and the
    // expensive part -- (pixel, tap) -> source address with padding / stride / zero insertion /
    // nearest up-sampling / concat -- is computed once per (row, tap, source) and reused for
    // all chunks that fall into it (one recompute per k block when C >= 64).
    const int half = tid & 1;
    const int row0 = tid >> 1;                  // rows row0 and row0 + 64; (row0 + 64) & 7 == row0 & 7
    const uint32_t swz = (uint32_t)(row0 & 7);
    int pn_[2], iy0_[2], ix0_[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int g = m0 + row0 + 64 * h;
      if (g < p.M) {
        const int pn = g / (p.ho * p.wo);
        const int rem = g - pn * p.ho * p.wo;
        const int oy = rem / p.wo;
        pn_[h] = pn;
        iy0_[h] = oy * p.stride - p.pad;
        ix0_[h] = (rem - oy * p.wo) * p.stride - p.pad;
      } else {
        pn_[h] = 0;
        iy0_[h] = -(1 << 20);                   // fails every bounds test -> zero fill
        ix0_[h] = -(1 << 20);
      }
    }
    const bf16* S0 = reinterpret_cast<const bf16*>(p.src0);
    const bf16* S1 = reinterpret_cast<const bf16*>(p.src1);
    const bf16* Wt = reinterpret_cast<const bf16*>(p.weight);
    // (tap row, tap col, channel) of this thread's FIRST chunk of the current k block
    int tc = half * 8, ts = 0, tr = 0;
    while (tc >= p.ctot) {
      tc -= p.ctot;
      if (++ts == p.kw) { ts = 0; ++tr; }
    }
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % C::STAGES;
      if (kb >= C::STAGES) mbar_wait(sBar + 8 * (C::STAGES + s), ((kb / C::STAGES) & 1) ^ 1);
      const uint32_t a_st = sA + s * C::A_BYTES + (uint32_t)row0 * 128u;
      int cc = tc, cs = ts, cr = tr;            // walks this thread's 4 chunks (16 channels apart)
      int prev_r = -1, prev_s = -1, prev_src = -1;
      const bf16* base_[2] = {S0, S0};
      bool ok_[2] = {false, false};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int from1 = cc >= p.c0 ? 1 : 0;
        if (cr != prev_r || cs != prev_s || from1 != prev_src) {
          prev_r = cr; prev_s = cs; prev_src = from1;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            int iy = iy0_[h] + cr, ix = ix0_[h] + cs;
            bool ok = cr < p.kh;
            if (p.dil == 2) {
              ok = ok && (((iy | ix) & 1) == 0);
              iy >>= 1;
              ix >>= 1;
            }
            ok = ok && iy >= 0 && iy < p.hin && ix >= 0 && ix < p.win;
            const bf16* b = S0;
            if (ok) {
              if (!from1) {
                int sy = iy, sx = ix;
                if (p.up) {
                  sy = nearest_src(iy, p.sch, p.h0);
                  sx = nearest_src(ix, p.scw, p.w0);
                }
                b = S0 + ((size_t)(pn_[h] * p.h0 + sy) * p.w0 + sx) * p.c0;
              } else {
                b = S1 + ((size_t)(pn_[h] * p.hin + iy) * p.win + ix) * p.c1 - p.c0;
              }
            }
            base_[h] = b;
            ok_[h] = ok;
          }
        }
        const uint32_t joff = (((uint32_t)(half + 2 * i)) ^ swz) << 4;
        cp_async16(a_st + joff, ok_[0] ? base_[0] + cc : S0, ok_[0] ? 16u : 0u);
        cp_async16(a_st + 64u * 128u + joff, ok_[1] ? base_[1] + cc : S0, ok_[1] ? 16u : 0u);
        cc += 16;
        while (cc >= p.ctot) {
          cc -= p.ctot;
          if (++cs == p.kw) { cs = 0; ++cr; }
        }
      }
      tc += BKE;                                // next k block
      while (tc >= p.ctot) {
        tc -= p.ctot;
        if (++ts == p.kw) { ts = 0; ++tr; }
      }
      // weight tile: BN rows x 8 chunks (consecutive lanes -> consecutive chunks of a row)
      const uint32_t b_st = sB + s * C::B_BYTES;
      const int kbase = kb * BKE;
      for (int i = tid; i < BN * 8; i += NPROD) {
        const int n = i >> 3, jj = i & 7;
        const int k = kbase + jj * 8;
        const bool ok = (n0 + n < p.cout) && (k < p.K);
        const bf16* src = ok ? Wt + (size_t)(n0 + n) * p.K + k : Wt;
        cp_async16(b_st + (uint32_t)n * 128u + (((uint32_t)jj ^ (uint32_t)(n & 7)) << 4), src, ok ? 16u : 0u);
      }
      // publish the stage: the barrier fires when this thread's copies have landed (no thread-side
      // wait, so up to STAGES stages of loads stay in flight); the MMA warp does the proxy fence.
      cp_async_mbar_arrive_noinc(sBar + 8 * s);
    }

    // =========================================================== EPILOGUE (same warps)
    const int gm = m0 + tid;                      // TMEM lane == tile row == output pixel
    const bool mvalid = gm < p.M;
    mbar_wait(sBar + 8 * (2 * C::STAGES), 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const bool want_stats = p.ssum != nullptr;
    const bool vector_epilogue = (p.cout % 16 == 0) && !p.dst_f32 && p.act != RCFD_ACT_DEPTH_HEAD;
    const bf16* R = reinterpret_cast<const bf16*>(p.residual);
    bf16* D = reinterpret_cast<bf16*>(p.dst);
#pragma unroll 1
    for (int cb = 0; cb < BN; cb += 16) {
      float v[16];
      tmem_ld16(trow + cb, v);
      if (want_stats) {
        // butterfly transpose-reduce: 16 columns x 32 lanes -> lanes 0..15 hold one column sum each
        float s16[16], q16[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          s16[i] = v[i];            // rows >= M gathered zeros -> contribute nothing
          q16[i] = v[i] * v[i];
        }
        // step over lane bit 4 keeps all 16 columns (just halves the lanes), then halve columns
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          s16[i] += __shfl_xor_sync(0xffffffffu, s16[i], 16);
          q16[i] += __shfl_xor_sync(0xffffffffu, q16[i], 16);
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
        // lane l (< 16) now holds column bitrev-free index: column = (lane & 15) by construction
        if (lane < 16) {
          red[(warp * BN + cb + lane) * 2 + 0] = s16[0];
          red[(warp * BN + cb + lane) * 2 + 1] = q16[0];
        }
      }
      if (mvalid) {
        const int nb = n0 + cb;
        const size_t o = (size_t)gm * p.cout + nb;
        if (!vector_epilogue) {
          // generic path: cout not a multiple of 16 (e.g. the 1-channel output head), float
          // destination, depth-head activation
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
    }
    if (want_stats) {
      // combine the 4 warps' partials, one double atomic per channel and CTA
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int c = tid; c < BN; c += NPROD) {
        if (n0 + c < p.cout) {
          double s = 0.0, q = 0.0;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            s += (double)red[(w * BN + c) * 2 + 0];
            q += (double)red[(w * BN + c) * 2 + 1];
          }
          atomicAdd(p.ssum + n0 + c, s);
          atomicAdd(p.ssq + n0 + c, q);
        }
      }
    }
    tc_fence_before();
  } else {
    // =========================================================== MMA ISSUER (warp 4)
    const uint32_t idesc = umma_idesc(BN);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % C::STAGES;
      mbar_wait(sBar + 8 * s, (kb / C::STAGES) & 1);
      fence_proxy_async();      // generic-proxy (cp.async) writes -> async-proxy (UMMA) reads
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_st = sA + s * C::A_BYTES, b_st = sB + s * C::B_BYTES;
#pragma unroll
        for (int k = 0; k < BKE / 16; ++k) {
          umma_f16(tmem_base, umma_desc_sw128(a_st + k * 32), umma_desc_sw128(b_st + k * 32), idesc,
                   (uint32_t)((kb | k) != 0));
        }
        umma_commit(sBar + 8 * (C::STAGES + s));                  // frees the smem stage when the MMAs retire
        if (kb == num_kb - 1) umma_commit(sBar + 8 * (2 * C::STAGES));   // accumulator complete
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

template <int BN>
int launch_tc(const ConvKP& p, cudaStream_t st) {
  note_kernel("conv_tc_kernel<%d>", BN);
  typedef TcCfg<BN> C;
  static bool attr_set_dev[16] = {}; bool& attr_set = attr_set_dev[cur_dev()];   // per device: the attribute belongs to the device's copy of the kernel
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) { set_error("conv_tc: smem attribute: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
    attr_set = true;
  }
  dim3 grid(ceil_div(p.M, TM), ceil_div(p.cout, BN));
  conv_tc_kernel<BN><<<grid, NTHREADS, C::SMEM, st>>>(p);
  RCFD_CHECK_LAUNCH("conv_tc");
  return RCFD_OK;
}

}  // namespace

bool conv_tc_supported(const ConvKP& p, int dtype) {
  if (dtype != RCFD_BF16) return false;
  if (p.c0 % 8 != 0 || p.c1 % 8 != 0) return false;
  return true;
}

int conv_tc_launch(const ConvKP& p, cudaStream_t st) {
  if (p.cout % 128 == 0) return launch_tc<128>(p, st);
  if (p.cout % 64 == 0) return launch_tc<64>(p, st);
  if (p.cout % 32 == 0) return launch_tc<32>(p, st);
  if (p.cout <= 16) return launch_tc<16>(p, st);
  if (p.cout % 16 == 0 || p.cout < 32) return launch_tc<16>(p, st);
  return launch_tc<32>(p, st);
}

}  // namespace rcfd
