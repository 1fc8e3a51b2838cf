// TMA + tcgen05 persistent implicit-GEMM convolution (bf16, fp32 accumulation in TMEM).
//
// The fast engine for every convolution whose im2col rows are boxes of the NHWC tensor:
// stride 1 or 2, no up-sampling, no zero insertion (all encoder / decoder convs, the 1x1
// fusion and projection convs, and every stride-1 dgrad).  Per filter tap and 64-channel
// chunk ONE elected thread issues one 4-D TMA tile load
//       box = (BKC channels, TW, TH, 1)  at  (c, x0*s + tap_x - pad, y0*s + tap_y - pad, n)
// (traversal stride = conv stride; out-of-range coordinates are zero-filled by the TMA unit =
// the conv padding) which lands in shared memory already in the K-major swizzled layout that
// tcgen05.mma consumes, plus one 2-D load of the weight slice.  No per-element address
// arithmetic, no im2col buffer, concat = a second tensor map in the K loop.
//
// Persistent, warp-specialised CTAs (one per SM): warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2-5 = epilogue.  smem stages form an mbarrier ring (full: expect_tx bytes, empty:
// tcgen05.commit); the accumulator is double-buffered in TMEM so the epilogue of tile i
// (TMEM -> registers -> BN statistics / folded BN + activation + residual -> global) overlaps
// the main loop of tile i+1.  Output tile = TH x TW = 128 pixels of one image.
#include "tma_common.cuh"

namespace rcfd {
namespace {

using namespace tc;
using namespace tma;
constexpr int EPI_GROUPS = 2;     // epilogue warps per TMEM lane quarter (each takes every 4th column chunk)
constexpr int NEPI = 128 * EPI_GROUPS;
constexpr int NTHREADS = 64 + NEPI;     // warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue

// Build with -DRCFD_TRACE (tools/trace_tma.py) to stamp clock64() at the phases of a CTA's life; compiled out of the product.
#ifdef RCFD_TRACE
__device__ long long g_trace[148 * 32];
#define TRACE(slot) do { if (blockIdx.x < 148) g_trace[blockIdx.x * 32 + (slot)] = clock64(); } while (0)
#define TRACE_VAL(slot, v) do { if (blockIdx.x < 148) g_trace[blockIdx.x * 32 + (slot)] = (long long)(v); } while (0)
#else
#define TRACE(slot) do { } while (0)
#define TRACE_VAL(slot, v) do { } while (0)
#endif

struct TmaConvP {
  int n, ho, wo, cout;
  int kh, kw, stride, pad;
  int c0, c1;               // channels of source 0 / source 1
  int bkc;                  // channels per K step (16 / 32 / 64)
  int tw, th;               // spatial tile
  int tiles_x, tiles_y, tiles_n, num_tiles;
  int ksteps;               // kh*kw*((c0+c1)/bkc)
  void* dst;
  const float* scale; const float* shift;
  int act; float p0, p1;
  const void* residual;
  double* ssum; double* ssq;
  int accumulate, dst_f32;
  int kpair;                // 1: two 64-channel chunks per k-step (one full barrier / one commit per PAIR of stages)
  int phases;               // 1, or 4 = sub-pixel phases of a 2x nearest up-sampled 3x3 conv (2x2 taps each)
  int out_h, out_w, out_s;  // destination extent and pixel stride (out_s = 2 with phases)
  int splitk;               // 1, or S = 2 | 4: the k-loop of ONE tile is split over the S CTAs of a cluster (grid = tiles x S), partial
                            // accumulators merged through distributed shared memory; each CTA runs the epilogue of BN / S columns
};

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_addr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
  return v;
}

template <int BN>
struct TmaCfg {
  static constexpr int A_BYTES = TM * 128;                  // sized for bkc = 64
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN <= 32 ? 10 : (BN <= 64 ? 8 : (BN <= 128 ? 6 : 4));     // even: stages can be consumed in pairs
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int RED_BYTES = 2 * 4 * BN * 2 * 4;      // [acc][warp][BN][sum, sq]
  static constexpr int SMEM = STAGES * STAGE + RED_BYTES + 1024 + 256;
  static_assert(SMEM <= 227 * 1024 && STAGES % 2 == 0, "per-tap engine: stage ring");
};

template <int BN, bool SPLIT>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tma_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                const __grid_constant__ CUtensorMap map_w, const TmaConvP p) {
  typedef TmaCfg<BN> C;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sStage = base;
  const uint32_t sRed = base + C::STAGES * C::STAGE;
  const uint32_t sBar = sRed + C::RED_BYTES;        // full[S], empty[S], tfull[2], tempty[2]
  const uint32_t sTmem = sBar + 8 * (2 * C::STAGES + 4);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (sTmem - base));
  float* red = reinterpret_cast<float*>(gen_base + (sRed - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) TRACE(0);

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(sBar + 8 * s, 1);
      mbar_init(sBar + 8 * (C::STAGES + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(sBar + 8 * (2 * C::STAGES + a), 1);
      mbar_init(sBar + 8 * (2 * C::STAGES + 2 + a), NEPI);
    }
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

  const int a_bytes = TM * p.bkc * 2, b_bytes = BN * p.bkc * 2;
  const int chunks0 = p.c0 / p.bkc, chunks = (p.c0 + p.c1) / p.bkc;
  const int ctot = p.c0 + p.c1;
  // k-loop units (a 64-channel chunk of one tap, or a pair of them) and this CTA's share of them (split-K: rank = blockIdx.y)
  const int upt = p.kpair ? chunks / 2 : chunks;              // units per tap
  const int units = p.kh * p.kw * upt;
  const int S = SPLIT ? p.splitk : 1, srank = SPLIT ? (int)blockIdx.y : 0;       // SPLIT = false compiles the cluster paths away
  const int ubeg = (int)((long long)units * srank / S), uend = (int)((long long)units * (srank + 1) / S);

  if (warp == 0) {
    // =========================================================== TMA PRODUCER (one lane)
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a0) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
      if (p.c1 > 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a1) : "memory");
      uint32_t it = 0;
      TRACE(2);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int ph = tile % p.phases;
        int sp = tile / p.phases;
        const int nt = sp % p.tiles_n; sp /= p.tiles_n;
        const int tx = sp % p.tiles_x; sp /= p.tiles_x;
        const int ty = sp % p.tiles_y;
        const int img = sp / p.tiles_y;
        // phase (a, b) of the up-sampled conv reads rows i-1+a .. i+a: its "padding" is 1 - a
        const int x0 = tx * p.tw * p.stride - (p.pad - (ph & 1)), y0 = ty * p.th * p.stride - (p.pad - (ph >> 1));
        const int n0 = ph * p.cout + nt * BN;
        if (p.kpair) {
          // stages are filled and released in PAIRS (s, s+1): 8 MMAs per barrier round trip / commit instead of 4
          constexpr uint32_t NP = C::STAGES / 2;
          for (int u = ubeg; u < uend; ++u, ++it) {
            const int tap = u / upt, ch = 2 * (u - tap * upt);
            const int tr = tap / p.kw, ts = tap - tr * p.kw;
            const uint32_t pr = it % NP, s = 2 * pr;
            if (it >= NP) mbar_wait(sBar + 8 * (C::STAGES + s), ((it / NP) & 1) ^ 1);
            const uint32_t full = sBar + 8 * s;
            mbar_expect_tx(full, (uint32_t)(2 * (a_bytes + b_bytes)));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int c = ch + h;
              const uint32_t a_dst = sStage + (s + h) * C::STAGE, b_dst = a_dst + C::A_BYTES;
              if (c < chunks0) tma_load_4d(a_dst, &map_a0, full, c * p.bkc, x0 + ts, y0 + tr, img);
              else tma_load_4d(a_dst, &map_a1, full, (c - chunks0) * p.bkc, x0 + ts, y0 + tr, img);
              tma_load_2d(b_dst, &map_w, full, tap * ctot + c * p.bkc, n0);
            }
          }
          continue;
        }
        for (int u = ubeg; u < uend; ++u, ++it) {
          const int tap = u / upt, ch = u - tap * upt;
          const int tr = tap / p.kw, ts = tap - tr * p.kw;
          const int s = it % C::STAGES;
          if (it >= (uint32_t)C::STAGES) mbar_wait(sBar + 8 * (C::STAGES + s), ((it / C::STAGES) & 1) ^ 1);
          const uint32_t full = sBar + 8 * s;
          const uint32_t a_dst = sStage + s * C::STAGE, b_dst = a_dst + C::A_BYTES;
          mbar_expect_tx(full, (uint32_t)(a_bytes + b_bytes));
          if (ch < chunks0) tma_load_4d(a_dst, &map_a0, full, ch * p.bkc, x0 + ts, y0 + tr, img);
          else tma_load_4d(a_dst, &map_a1, full, (ch - chunks0) * p.bkc, x0 + ts, y0 + tr, img);
          tma_load_2d(b_dst, &map_w, full, tap * ctot + ch * p.bkc, n0);
        }
      }
      TRACE(3);
    }
    if (SPLIT) { __syncwarp(); cluster_sync_all(); cluster_sync_all(); }      // the epilogue's two cluster barriers count every thread
  } else if (warp == 1) {
    // =========================================================== MMA ISSUER
    const uint32_t idesc = umma_idesc(BN);
    // swizzle span = bkc*2 bytes: 128 -> layout 2 (SBO 1024), 64 -> layout 4 (SBO 512), 32 -> layout 6 (SBO 256)
    const uint32_t layout = p.bkc == 64 ? 2u : (p.bkc == 32 ? 4u : 6u);
    const uint32_t sbo = (uint32_t)(8 * p.bkc * 2);
    uint32_t it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t acc = tcount & 1;
      if (tcount >= 2) mbar_wait(sBar + 8 * (2 * C::STAGES + 2 + acc), ((tcount >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      if (p.kpair) {
        constexpr uint32_t NP = C::STAGES / 2;
        for (int kp = ubeg; kp < uend; ++kp, ++it) {
          const uint32_t pr = it % NP, s = 2 * pr;
          mbar_wait(sBar + 8 * s, (it / NP) & 1);
          tc_fence_after();
          if (lane == 0 && it == 0) TRACE(4);
          if (lane == 0 && it == 1) TRACE(14);
          if (lane == 0 && it == 2) TRACE(15);
          if (lane == 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t a_st = sStage + (s + h) * C::STAGE, b_st = a_st + C::A_BYTES;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(d_tmem, umma_desc(a_st + k * 32, 16, sbo, layout), umma_desc(b_st + k * 32, 16, sbo, layout), idesc,
                         (uint32_t)(((kp - ubeg) | h | k) != 0));
            }
            umma_commit(sBar + 8 * (C::STAGES + s));
            if (kp == uend - 1) umma_commit(sBar + 8 * (2 * C::STAGES + acc));
            if (kp == uend - 1) { if (tcount == 0) TRACE(5); TRACE(6); }
          }
          __syncwarp();
        }
        continue;
      }
      for (int ks = ubeg; ks < uend; ++ks, ++it) {
        const int s = it % C::STAGES;
        mbar_wait(sBar + 8 * s, (it / C::STAGES) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_st = sStage + s * C::STAGE, b_st = a_st + C::A_BYTES;
          for (int k = 0; k < p.bkc / 16; ++k) {
            umma_f16(d_tmem, umma_desc(a_st + k * 32, 16, sbo, layout), umma_desc(b_st + k * 32, 16, sbo, layout), idesc,
                     (uint32_t)(((ks - ubeg) | k) != 0));
          }
          umma_commit(sBar + 8 * (C::STAGES + s));
          if (ks == uend - 1) umma_commit(sBar + 8 * (2 * C::STAGES + acc));
          if (it == 0) TRACE(4);
          if (ks == uend - 1) { if (tcount == 0) TRACE(5); TRACE(6); }
        }
        __syncwarp();
      }
    }
    tc_fence_before();
    if (SPLIT) { __syncwarp(); cluster_sync_all(); cluster_sync_all(); }
  } else {
    // =========================================================== EPILOGUE (warps 2..9)
    // EPI_GROUPS warps per TMEM lane quarter, each taking every EPI_GROUPS-th CW-column chunk of its 32 tile rows; BatchNorm statistics
    // by a butterfly transpose-reduce over CW = 32 columns at a time (31 exchange steps leave lane l with the total of
    // column l: half the shuffles of the first version's 16-column form).  In-kernel clock stamps
    // (profiles/r2_conv_tma_trace.txt) had shown 3.9 us per 128 x 128 tile for the 4-warp epilogue: as long as half the
    // k-loop of a 256-channel 3x3 layer, and not overlapped when a CTA owns one tile (every level <= 22x44 at batch 8).
    constexpr int CW = BN >= 64 ? 32 : 16;     // columns per chunk
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;           // which chunks of the row this warp takes
    const int r = q * 32 + lane;               // tile row == TMEM lane
    const int th = r / p.tw, tw_ = r - th * p.tw;
    const bool want_stats = p.ssum != nullptr;
    const bool vector_epilogue = (p.cout % 16 == 0) && !p.dst_f32 && p.act != RCFD_ACT_DEPTH_HEAD;
    const bf16* R = reinterpret_cast<const bf16*>(p.residual);
    bf16* D = reinterpret_cast<bf16*>(p.dst);
    const int etid = tid - 64;                 // index within the epilogue group
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t acc = tcount & 1;
      const int ph = tile % p.phases;
      int sp = tile / p.phases;
      const int nt = sp % p.tiles_n; sp /= p.tiles_n;
      const int tx = sp % p.tiles_x; sp /= p.tiles_x;
      const int ty = sp % p.tiles_y;
      const int img = sp / p.tiles_y;
      const int oy = ty * p.th + th, ox = tx * p.tw + tw_;
      // phases: destination pixel (2 oy + a, 2 ox + b); an odd destination extent (stride-2 dgrad of 11 -> 6 rows) drops the last one
      const bool mvalid = oy < p.ho && ox < p.wo && (oy * p.out_s + (ph >> 1)) < p.out_h && (ox * p.out_s + (ph & 1)) < p.out_w;
      const size_t gm = ((size_t)img * p.out_h + (oy * p.out_s + (ph >> 1))) * p.out_w + (ox * p.out_s + (ph & 1));
      const int n0 = nt * BN;
      float* redt = red + acc * (4 * BN * 2);
      mbar_wait(sBar + 8 * (2 * C::STAGES + acc), (tcount >> 1) & 1);
      tc_fence_after();
      if (tid == 64 && tcount == 0) TRACE(7);
      const uint32_t trow = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      // split-K: every CTA of the cluster stages its partial accumulator ([column][row] floats in the stage ring, which is
      // free once the accumulator is ready), then sums ITS column slice over all peers and runs the epilogue on that slice
      int c_lo = 0, c_hi = BN;
      if (SPLIT) {
        float* stg = reinterpret_cast<float*>(gen_base);
#pragma unroll 1
        for (int cb = grp * CW; cb < BN; cb += EPI_GROUPS * CW) {
          uint32_t raw[CW];
#pragma unroll
          for (int j = 0; j < CW; j += 16) tmem_ld16_nowait(trow + cb + j, raw + j);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < CW; ++i) stg[(cb + i) * TM + r] = __uint_as_float(raw[i]);      // lanes = consecutive rows
        }
        cluster_sync_all();
        c_lo = srank * (BN / S);
        c_hi = c_lo + BN / S;
      }
#pragma unroll 1
      for (int cb = c_lo + grp * CW; cb < c_hi; cb += EPI_GROUPS * CW) {
        float v[CW];
        if (SPLIT) {
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = 0.f;
          for (int pr = 0; pr < S; ++pr) {
#pragma unroll
            for (int i = 0; i < CW; ++i) v[i] += ld_dsmem_f32(sStage + (uint32_t)(((cb + i) * TM + r) * 4), (uint32_t)pr);
          }
        } else {
          uint32_t raw[CW];
#pragma unroll
          for (int j = 0; j < CW; j += 16) tmem_ld16_nowait(trow + cb + j, raw + j);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(raw[i]);
        }
        if (want_stats) {
          float s_[CW], q_[CW];
#pragma unroll
          for (int i = 0; i < CW; ++i) {
            const float x = mvalid ? v[i] : 0.f;        // rows outside the image carry partial sums of real taps
            s_[i] = x;
            q_[i] = x * x;
          }
          if (CW == 16) {                               // 16 columns on 32 lanes: fold the two half-warps first
#pragma unroll
            for (int i = 0; i < CW; ++i) {
              s_[i] += __shfl_xor_sync(0xffffffffu, s_[i], 16);
              q_[i] += __shfl_xor_sync(0xffffffffu, q_[i], 16);
            }
          }
#pragma unroll
          for (int w = CW / 2; w >= 1; w >>= 1) {
            const bool hi = (lane & w) != 0;
#pragma unroll
            for (int i = 0; i < w; ++i) {
              const float send_s = hi ? s_[i] : s_[i + w];
              const float keep_s = hi ? s_[i + w] : s_[i];
              s_[i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, w);
              const float send_q = hi ? q_[i] : q_[i + w];
              const float keep_q = hi ? q_[i + w] : q_[i];
              q_[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, w);
            }
          }
          if (lane < CW) {
            redt[(q * BN + cb + lane) * 2 + 0] = s_[0];
            redt[(q * BN + cb + lane) * 2 + 1] = q_[0];
          }
        }
        if (mvalid) {
#pragma unroll
          for (int c16 = 0; c16 < CW; c16 += 16) {
            const int nb = n0 + cb + c16;
            const size_t o = gm * p.cout + nb;
            float* v16 = v + c16;
            if (!vector_epilogue) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int n = nb + i;
                if (n < p.cout) {
                  float x = v16[i];
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
                for (int i = 0; i < 16; ++i) v16[i] = fmaf(v16[i], __ldg(p.scale + nb + i), __ldg(p.shift + nb + i));
              }
              if (p.act == RCFD_ACT_LEAKY) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v16[i] = leaky(v16[i]);
              } else if (p.act == RCFD_ACT_SIGMOID) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v16[i] = sigmoid_precise(v16[i]);
              }
              if (R) {
                const uint4 r0 = *reinterpret_cast<const uint4*>(R + o);
                const uint4 r1 = *reinterpret_cast<const uint4*>(R + o + 8);
                const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr[i]));
                  v16[2 * i] = leaky(v16[2 * i] + f.x);
                  v16[2 * i + 1] = leaky(v16[2 * i + 1] + f.y);
                }
              }
              if (p.accumulate) {
                const uint4 r0 = *reinterpret_cast<const uint4*>(D + o);
                const uint4 r1 = *reinterpret_cast<const uint4*>(D + o + 8);
                const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr[i]));
                  v16[2 * i] += f.x;
                  v16[2 * i + 1] += f.y;
                }
              }
              uint32_t pk[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                __nv_bfloat162 h = __floats2bfloat162_rn(v16[2 * i], v16[2 * i + 1]);
                pk[i] = *reinterpret_cast<uint32_t*>(&h);
              }
              *reinterpret_cast<uint4*>(D + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(D + o + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      mbar_arrive(sBar + 8 * (2 * C::STAGES + 2 + acc));
      if (tid == 64 && tcount == 0) TRACE(8);
      if (want_stats) {
        asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory");
        for (int c = c_lo + etid; c < c_hi; c += NEPI) {
          if (n0 + c < p.cout) {
            double s = 0.0, qq = 0.0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              s += (double)redt[(w * BN + c) * 2 + 0];
              qq += (double)redt[(w * BN + c) * 2 + 1];
            }
            atomicAdd(p.ssum + n0 + c, s);
            atomicAdd(p.ssq + n0 + c, qq);
          }
        }
      }
      if (tid == 64) { if (tcount == 0) TRACE(9); TRACE(10); TRACE_VAL(12, tcount + 1); TRACE_VAL(13, p.ksteps); }
      if (SPLIT) cluster_sync_all();            // peers have read this CTA's staging: it may exit
    }
  }
  __syncthreads();
  if (tid == 0) TRACE(11);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

}  // namespace
extern int g_tma_bn_cap;
extern int g_tma_pair;
extern int g_tma_split_k;
namespace {

template <int BN>
int launch_tma(const ConvKP& k, const TmaConvP& tp, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w,
               cudaStream_t st) {
  note_kernel("conv_tma_kernel<%d>", BN);
  typedef TmaCfg<BN> C;
  static bool attr_set_dev[16] = {}; bool& attr_set = attr_set_dev[cur_dev()];   // per device: the attribute belongs to the device's copy of the kernel
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tma_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tma_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) { set_error("conv_tma: smem attribute: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
    attr_set = true;
  }
  (void)k;
  // split-K over a thread-block cluster for layers with fewer tiles than SMs (every level <= 11x22 at batch 8, everything at
  // batch 1): there the latency of ONE CTA's k-loop is the kernel's duration, so S CTAs share it and merge their partial
  // accumulators through distributed shared memory.  S = the largest of 4, 2 that keeps the whole grid resident at once.
  constexpr int CWh = BN >= 64 ? 32 : 16;
  TmaConvP t2 = tp;
  t2.splitk = 1;
  // Measured: inference batch 1 0.653 -> 0.592 ms, batch 8 1.416 -> 1.392 ms; the training step got slower with it (the extra
  // CTAs and the all-or-nothing start of a cluster compete with the weight-gradient kernels), so only the eval-mode
  // epilogue (folded BatchNorm, no statistics) splits; rcfd_set_option("tma_split_k", 2) forces it everywhere (tests).
  if (g_tma_split_k == 2 || (g_tma_split_k == 1 && tp.ssum == nullptr && tp.scale != nullptr)) {
    const int units = tp.kh * tp.kw * ((tp.c0 + tp.c1) / tp.bkc) / (tp.kpair ? 2 : 1);
    static int max_clusters_dev[16][3] = {};
    for (int sidx = 2; sidx >= 1; --sidx) {
      const int S = 1 << sidx;
      if ((BN / S) % CWh != 0 || BN / S < CWh || units < 2 * S || tp.num_tiles * S > num_sms()) continue;
      int& mc = max_clusters_dev[cur_dev()][sidx];
      if (mc == 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(1, S, 1); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = C::SMEM;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = S; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, conv_tma_kernel<BN, true>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = -1; }
        mc = n;
      }
      if (mc >= tp.num_tiles) { t2.splitk = S; break; }
    }
  }
  if (t2.splitk > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tp.num_tiles, t2.splitk, 1);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = C::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = t2.splitk; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tma_kernel<BN, true>, a0, a1, w, t2);
    if (e != cudaSuccess) { set_error("conv_tma (split-K) launch: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
    RCFD_CHECK_LAUNCH("conv_tma");
    return RCFD_OK;
  }
  int grid = tp.num_tiles < num_sms() ? tp.num_tiles : num_sms();
  conv_tma_kernel<BN, false><<<grid, NTHREADS, C::SMEM, st>>>(a0, a1, w, t2);
  RCFD_CHECK_LAUNCH("conv_tma");
  return RCFD_OK;
}

}  // namespace

#ifdef RCFD_TRACE
extern "C" int rcfd_debug_read_trace(long long* host_out, int n) {
  return cudaMemcpyFromSymbol(host_out, g_trace, sizeof(long long) * n) == cudaSuccess ? 0 : -2;
}
extern "C" int rcfd_debug_clear_trace() {
  static long long zeros[148 * 32] = {};
  return cudaMemcpyToSymbol(g_trace, zeros, sizeof(zeros)) == cudaSuccess ? 0 : -2;
}
#endif

int g_tma_split_k = 1;     // rcfd_set_option("tma_split_k"): 0 = never split the k-loop of a tile over a cluster
int g_tma_pair = 1;        // rcfd_set_option("tma_pair"): 0 = one 64-channel chunk per k-step (4 MMAs per commit)
int g_tma_bn_cap = -1;    // rcfd_set_option("tma_bn_cap"): -1 = widest tile (default: narrower tiles measured slower, k-steps are latency bound), 0 = occupancy heuristic, n = cap the cout tile at n

// 2x nearest up-sampling folded into a 3x3 / stride-1 / pad-1 conv == four 2x2 convs (one per
// output sub-pixel phase) on the LOW-RES source with summed weights: 4/9 of the MACs, every
// load a TMA box (needs the phase weights from rcfd_pack_upconv2x_weight).
bool conv_tma_up2x_supported(const ConvKP& p, int dtype) {
  if (dtype != RCFD_BF16 || !p.up || p.weight_up2x == nullptr) return false;
  if (p.kh != 3 || p.kw != 3 || p.stride != 1 || p.pad != 1 || p.dil != 1 || p.c1 != 0) return false;
  if (p.hin != 2 * p.h0 || p.win != 2 * p.w0 || p.ho != p.hin || p.wo != p.win) return false;
  if (p.c0 % 16 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(p.src0) & 15) || (reinterpret_cast<uintptr_t>(p.weight_up2x) & 15)) return false;
  return get_encode() != nullptr;
}

// Data gradient of a 3x3 / stride-2 / pad-1 convolution WITHOUT zero insertion: the destination pixel (2i + a, 2j + b)
// only sees the taps of matching parity, i.e. phase (a, b) is a 2x2 convolution over dy rows i-1+a .. i+a (the same
// geometry as the sub-pixel up-conv above) with the phase weights of rcfd_pack_dgrad_s2_weight: 16 tap GEMMs on the
// dy grid instead of 9 on the 4x larger zero-inserted grid through the gather engine.
bool conv_tma_dgrad_s2_supported(const ConvKP& p, int dtype) {
  if (dtype != RCFD_BF16 || p.dil != 2 || p.weight_up2x == nullptr || p.up) return false;
  if (p.kh != 3 || p.kw != 3 || p.stride != 1 || p.pad != 1 || p.c1 != 0) return false;
  if (p.c0 % 16 != 0 || p.cout % 16 != 0) return false;
  if (!(p.ho == 2 * p.h0 || p.ho == 2 * p.h0 - 1) || !(p.wo == 2 * p.w0 || p.wo == 2 * p.w0 - 1)) return false;
  if ((reinterpret_cast<uintptr_t>(p.src0) & 15) || (reinterpret_cast<uintptr_t>(p.weight_up2x) & 15)) return false;
  return get_encode() != nullptr;
}

bool conv_tma_supported(const ConvKP& p, int dtype) {
  if (conv_tma_up2x_supported(p, dtype) || conv_tma_dgrad_s2_supported(p, dtype)) return true;
  if (dtype != RCFD_BF16) return false;
  if (p.up || p.dil != 1) return false;
  if (p.stride != 1 && p.stride != 2) return false;
  if (p.c0 % 16 != 0 || p.c1 % 16 != 0) return false;
  if (p.K % 8 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(p.src0) & 15) || (p.c1 > 0 && (reinterpret_cast<uintptr_t>(p.src1) & 15)) ||
      (reinterpret_cast<uintptr_t>(p.weight) & 15))
    return false;
  return get_encode() != nullptr;
}

int conv_tma_launch(const ConvKP& pin, cudaStream_t st) {
  ConvKP p = pin;
  TmaConvP t;
  t.phases = 1; t.out_h = p.ho; t.out_w = p.wo; t.out_s = 1;
  t.kpair = 0;
  if (conv_tma_up2x_supported(p, RCFD_BF16) || conv_tma_dgrad_s2_supported(p, RCFD_BF16)) {
    // run on the low-res grid: 2x2 taps, per-phase weights, destination pixels (2i+a, 2j+b)
    t.phases = 4; t.out_s = 2; p.dil = 1;
    p.ho = p.h0; p.wo = p.w0; p.hin = p.h0; p.win = p.w0;
    p.kh = 2; p.kw = 2; p.K = 4 * p.c0;
    p.weight = p.weight_up2x;
  }
  t.n = p.n; t.ho = p.ho; t.wo = p.wo; t.cout = p.cout;
  t.kh = p.kh; t.kw = p.kw; t.stride = p.stride; t.pad = p.pad;
  t.c0 = p.c0; t.c1 = p.c1;
  auto gcd_ok = [&](int b) { return p.c0 % b == 0 && p.c1 % b == 0; };
  t.bkc = gcd_ok(64) ? 64 : (gcd_ok(32) ? 32 : 16);
  // spatial tile: widest TW in {32,16,8} that wastes the fewest output pixels
  int best_tw = 8;
  long best_waste = -1;
  for (int tw = 32; tw >= 8; tw >>= 1) {
    const int th = TM / tw;
    const long covered = (long)ceil_div(p.wo, tw) * tw * ceil_div(p.ho, th) * th;
    if (best_waste < 0 || covered < best_waste) { best_waste = covered; best_tw = tw; }
  }
  t.tw = best_tw; t.th = TM / best_tw;
  t.tiles_x = ceil_div(p.wo, t.tw); t.tiles_y = ceil_div(p.ho, t.th);
  int bn = p.cout % 128 == 0 ? 128 : (p.cout % 64 == 0 ? 64 : (p.cout % 32 == 0 ? 32 : (p.cout <= 16 || p.cout % 16 == 0 || p.cout < 32 ? 16 : 32)));
  // small spatial extents (the 6x11 ... 22x44 encoder levels): a CTA's k-loop is bound by its own
  // L2 -> shared-memory ingest (A tile + B tile per k-step), so narrower cout tiles on more SMs win
  // as long as the grid still fits one wave.
  {
    const int mtiles = p.n * t.tiles_y * t.tiles_x * t.phases;
    if (g_tma_bn_cap > 0) {
      while (bn > g_tma_bn_cap && bn > 16) bn >>= 1;
    } else if (g_tma_bn_cap == 0) {
      while (bn > 32 && mtiles * ceil_div(p.cout, bn) * 2 <= num_sms()) bn >>= 1;
    }
  }
  t.tiles_n = ceil_div(p.cout, bn);
  t.num_tiles = p.n * t.tiles_y * t.tiles_x * t.tiles_n * t.phases;
  t.ksteps = p.kh * p.kw * ((p.c0 + p.c1) / t.bkc);
  // 128 / 256-channel layers: consume the 64-channel chunks two at a time (needs an even chunk count per tap, and the
  // split between the two sources on a pair boundary)
  t.kpair = (g_tma_pair && t.bkc == 64 && ((p.c0 + p.c1) / 64) % 2 == 0 && (p.c0 / 64) % 2 == 0) ? 1 : 0;
  t.dst = p.dst; t.scale = p.scale; t.shift = p.shift; t.act = p.act; t.p0 = p.p0; t.p1 = p.p1;
  t.residual = p.residual; t.ssum = p.ssum; t.ssq = p.ssq; t.accumulate = p.accumulate; t.dst_f32 = p.dst_f32;
  alignas(64) CUtensorMap a0, a1, w;
  if (!make_act_map(&a0, p.src0, p.n, p.hin, p.win, p.c0, t.bkc, t.tw, t.th, p.stride) ||
      !make_act_map(&a1, p.c1 > 0 ? p.src1 : p.src0, p.n, p.hin, p.win, p.c1 > 0 ? p.c1 : p.c0, t.bkc, t.tw, t.th, p.stride) ||
      !make_w_map(&w, p.weight, p.cout * t.phases, p.K, t.bkc, bn)) {
    set_error("conv_tma: cuTensorMapEncodeTiled failed");
    return RCFD_ECUDA;
  }
  switch (bn) {
    case 128: return launch_tma<128>(p, t, a0, a1, w, st);
    case 64: return launch_tma<64>(p, t, a0, a1, w, st);
    case 32: return launch_tma<32>(p, t, a0, a1, w, st);
    default: return launch_tma<16>(p, t, a0, a1, w, st);
  }
}

}  // namespace rcfd
