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


// ===================================================================================================================
// Second generation (round 2), used for every layer with >= 64 channels on both sides.  What the in-kernel clock stamps
// of the kernel above showed (profiles/r2_conv_tma_trace.txt): every 64-pixel step pulls a fresh A tile and a fresh dY
// tile with no reuse -- 144 CTAs x 18 steps x 48 KB = 124 MB of L2 -> shared-memory traffic for a 256 -> 256 layer whose
// operands are 8 MB, i.e. the k-loop runs at the L2 bandwidth, dY being re-read by all 18 k-tiles -- and a third of the
// kernel is the fp32 atomic merge of the 8 pixel splits (plus a memset kernel in front).  Here:
//   * a CTA owns KT k-tiles (KT x BN <= 512 TMEM columns, the whole tensor memory) that SHARE the dY tile of a step:
//     bytes per MMA drop by a third to a half, and half as many CTAs do the same work (weight gradients sit beside the
//     step's dependency chains: what they cost is SM time);
//   * the pixel splits of a unit form a thread-block CLUSTER; their partial accumulators are summed through distributed
//     shared memory (every CTA stages its accumulator, then sums its own share of the columns from all peers) and
//     written once with plain coalesced stores: no atomics, no memset, deterministic.  A cluster lives inside one GPC and
//     every CTA takes a whole SM, so at most 16 clusters of 8 (32 of 4, 72 of 2) are resident at once; and one CTA ingests
//     only ~50 GB/s through TMA boxes with 128-byte rows (measured: 64 KB steps take 1.7 us whatever the L2 load), so the
//     k-loop needs MANY CTAs: when one cluster per unit leaves SMs idle, several clusters ("groups") share a unit and
//     merge with float4 atomics (<= 6 contributions per element instead of 8-18 scalar ones).
struct WgTma2P {
  WgTmaP b;
  int kt;              // k-tiles (128 rows of dW^T) per CTA
  int pb;              // pixels per reduction step (64 | 32)
  int stages;
  int ktiles;          // ceil(K / 128)
  int cs;              // cluster size: pixel splits merged through distributed shared memory (gridDim.z = cs x groups)
};

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_addr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote) : "memory");
  return v;
}

constexpr int WG2_SMEM = 200 * 1024 + 1024 + 256;      // stage ring (<= 200 KB, reused as the reduction staging) + barriers

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
wgrad_tma2_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                  const __grid_constant__ CUtensorMap map_dy, const WgTma2P q, float* __restrict__ dw) {
  const WgTmaP& p = q.b;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sStage = base;
  const uint32_t sBar = base + 200 * 1024;                    // full[S], empty[S], accum
  const uint32_t sTmem = sBar + 8 * (2 * 8 + 1);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + (sTmem - base));
  float* stg = reinterpret_cast<float*>(gen);                  // [BN columns][128 rows] fp32, after the k-loop

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = q.stages, KT = q.kt, PBX = q.pb;
  const int kt0 = blockIdx.x * KT;                             // first k-tile of this unit
  const int kt_live = min(KT, q.ktiles - kt0);
  const int n0 = blockIdx.y * BN;
  const int split = blockIdx.z, nsplit = gridDim.z;            // pixel split; clusters are runs of q.cs consecutive z
  const int CS = q.cs, crank = split % CS;
  const bool atomic_merge = nsplit > CS;                       // several clusters per unit (dw was zeroed by the launcher)
  const int my_tiles = (p.num_ptiles - split + nsplit - 1) / nsplit;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(sBar + 8 * s, 1);
      mbar_init(sBar + 8 * (8 + s), 1);
    }
    mbar_init(sBar + 8 * 16, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sTmem) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int ctot = p.c0 + p.c1;
  const int a_blk_bytes = PBX * 64 * 2;               // one box: PB pixels x 64 channels
  const int a_tile_bytes = 2 * a_blk_bytes;           // one k-tile = 128 k = two boxes
  const int nblocks = BN / 64;
  const int b_blk_bytes = PBX * 64 * 2;
  const int stage_bytes = KT * a_tile_bytes + nblocks * b_blk_bytes;

  if (warp == 0) {
    if (lane == 0 && my_tiles > 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a0) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
      if (p.c1 > 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a1) : "memory");
      int a_live = 0;
      for (int b = 0; b < 2 * kt_live; ++b) a_live += (kt0 * TM + b * 64) < p.K ? 1 : 0;
      int b_live = 0;
      for (int b = 0; b < nblocks; ++b) b_live += (n0 + b * 64) < p.cout ? 1 : 0;
      const uint32_t tx_bytes = (uint32_t)(a_live * a_blk_bytes + b_live * b_blk_bytes);
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % S;
        if (i >= S) mbar_wait(sBar + 8 * (8 + s), ((i / S) & 1) ^ 1);
        int sp = split + i * nsplit;
        const int tx = sp % p.tiles_x; sp /= p.tiles_x;
        const int ty = sp % p.tiles_y;
        const int img = sp / p.tiles_y;
        const int ox0 = tx * p.tw, oy0 = ty * p.th;
        const uint32_t full = sBar + 8 * s;
        const uint32_t a_dst = sStage + s * stage_bytes, b_dst = a_dst + KT * a_tile_bytes;
        mbar_expect_tx(full, tx_bytes);
        for (int b = 0; b < 2 * kt_live; ++b) {
          const int k = kt0 * TM + b * 64;
          if (k >= p.K) break;
          const int tap = k / ctot, c = k - tap * ctot;
          const int tr = tap / p.kw, ts = tap - tr * p.kw;
          const int ix = ox0 * p.stride - p.pad + ts, iy = oy0 * p.stride - p.pad + tr;
          if (c < p.c0) tma_load_4d(a_dst + b * a_blk_bytes, &map_a0, full, c, ix, iy, img);
          else tma_load_4d(a_dst + b * a_blk_bytes, &map_a1, full, c - p.c0, ix, iy, img);
        }
        for (int b = 0; b < nblocks; ++b) {
          const int co = n0 + b * 64;
          if (co >= p.cout) break;
          tma_load_4d(b_dst + b * b_blk_bytes, &map_dy, full, co, ox0, oy0, img);
        }
      }
    }
  } else if (warp == 1) {
    // both operands MN-major, 64-channel boxes: 128-byte swizzle, 8 pixel rows = 1024 bytes
    const uint32_t idesc = umma_idesc_ex(TM, BN, 1, 1);
    for (int i = 0; i < my_tiles; ++i) {
      const int s = i % S;
      mbar_wait(sBar + 8 * s, (i / S) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_st = sStage + s * stage_bytes, b_st = a_st + KT * a_tile_bytes;
        for (int kt = 0; kt < kt_live; ++kt) {
          for (int kk = 0; kk < PBX / 16; ++kk) {
            umma_f16(tmem_base + kt * BN, umma_desc(a_st + kt * a_tile_bytes + kk * 2048, (uint32_t)a_blk_bytes, 1024, 2),
                     umma_desc(b_st + kk * 2048, (uint32_t)b_blk_bytes, 1024, 2), idesc, (uint32_t)((i | kk) != 0));
          }
        }
        umma_commit(sBar + 8 * (8 + s));
        if (i == my_tiles - 1) umma_commit(sBar + 8 * 16);
      }
      __syncwarp();
    }
    tc_fence_before();
  }

  // ===================================================== merge of the pixel splits through distributed shared memory
  // (all threads: the cluster barriers count every thread of every CTA)
  __syncwarp();
  if (warp >= 2 && my_tiles > 0) {
    mbar_wait(sBar + 8 * 16, 0);                      // every MMA has completed: the stage ring is free, the accumulators final
    tc_fence_after();
  }
  const int etid = tid - 64;                          // 0..127 for the epilogue warps
  const int cols_per = BN / CS;                       // CS in {1, 2, 4, 8}: this CTA sums columns [crank * cols_per, +cols_per)
  for (int kt = 0; kt < kt_live; ++kt) {
    if (warp >= 2) {
      const int qd = warp & 3;
      const int row = qd * 32 + lane;
      const uint32_t trow = tmem_base + kt * BN + ((uint32_t)(qd * 32) << 16);
#pragma unroll 1
      for (int cb = 0; cb < BN; cb += 16) {
        float v[16];
        if (my_tiles > 0) {
          tmem_ld16(trow + cb, v);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) stg[(cb + i) * TM + row] = v[i];      // lanes = consecutive rows: conflict free
      }
    }
    cluster_sync_all();
    if (warp >= 2) {
      // this CTA's columns, all 128 rows, summed over the cluster: float4 along the rows (= consecutive k of one cout row of dW)
      const int k_base = (kt0 + kt) * TM;
      for (int e = etid; e < cols_per * (TM / 4); e += 128) {
        const int c = crank * cols_per + e / (TM / 4);
        const int r4 = (e % (TM / 4)) * 4;
        const uint32_t laddr = base + (uint32_t)((c * TM + r4) * 4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int pr = 0; pr < CS; ++pr) {
          const float4 x = ld_dsmem_f4(laddr, (uint32_t)pr);
          acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
        }
        const int co = n0 + c, k = k_base + r4;
        if (co < p.cout && k < p.K) {                                                                  // K % 64 == 0
          float4* dst = reinterpret_cast<float4*>(dw + (size_t)co * p.K + k);
          if (atomic_merge) atomicAdd(dst, acc); else *dst = acc;
        }
      }
    }
    cluster_sync_all();                               // peers have read this CTA's staging: it may be rewritten / the CTA may exit
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

}  // namespace
// rcfd_set_option("wgrad_tma_cs_max" / "wgrad_tma_groups"): weight gradients run beside the step's dependency chains, so what
// they cost is SM time, not latency: ONE cluster per unit (fewest CTAs, no atomics) measured 5.71 ms per training step,
// filling the GPU with several clusters per unit 6.05 ms (first-generation kernel: 5.82 ms).
int g_wgrad_tma_cs_max = 8;
int g_wgrad_tma_groups = 0;
namespace {

template <int BN>
int launch_wg_tma2(const WgTma2P& t, int units_k, int units_n, const CUtensorMap& a0, const CUtensorMap& a1,
                   const CUtensorMap& dy, float* dw, cudaStream_t st) {
  note_kernel("wgrad_tma2_kernel<%d>", BN);
  static bool attr_set_dev[16] = {}; bool& attr_set = attr_set_dev[cur_dev()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tma2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG2_SMEM);
    if (e != cudaSuccess) { set_error("wgrad_tma2: smem attribute: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
    attr_set = true;
  }
  // cluster size cs (pixel splits merged in shared memory) and groups (clusters per unit, merged with atomics): the
  // combination that puts the most CTAs on the GPU at once (ties: the larger cluster), every CTA keeping >= 2 pixel tiles
  static int max_clusters_dev[16][4] = {};
  const int units = units_k * units_n;
  int cs = 1, groups = 1, best_ctas = 0;
  for (int i = 3; i >= 0; --i) {
    const int c = 1 << i;
    int& mc = max_clusters_dev[cur_dev()][i];
    if (mc == 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(1, 1, c); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = WG2_SMEM;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = c;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, wgrad_tma2_kernel<BN>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = c == 1 ? num_sms() : -1; }
      mc = n;
    }
    if (mc < units || c > g_wgrad_tma_cs_max) continue;
    int g = g_wgrad_tma_groups ? mc / units : 1;
    while (g > 1 && t.b.num_ptiles < 2 * g * c) --g;
    if (t.b.num_ptiles < 2 * c && c > 1) continue;
    if (units * g * c > best_ctas) { best_ctas = units * g * c; cs = c; groups = g; }
  }
  if (best_ctas == 0) { set_error("wgrad_tma2: %d units do not fit the GPU", units); return RCFD_EUNSUPPORTED; }
  WgTma2P tq = t;
  tq.cs = cs;
  if (groups > 1) {
    cudaError_t em = cudaMemsetAsync(dw, 0, (size_t)t.b.cout * t.b.K * sizeof(float), st);
    if (em != cudaSuccess) { set_error("wgrad_tma2 memset: %s", cudaGetErrorString(em)); return RCFD_ECUDA; }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(units_k, units_n, cs * groups);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = WG2_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = cs;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, wgrad_tma2_kernel<BN>, a0, a1, dy, tq, dw);
  if (e != cudaSuccess) { set_error("wgrad_tma2 launch: %s", cudaGetErrorString(e)); return RCFD_ECUDA; }
  RCFD_CHECK_LAUNCH("wgrad_tma2");
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

int g_wgrad_tma_v2 = 1;       // rcfd_set_option("wgrad_tma_v2"): 0 = the first-generation kernel everywhere (A/B measurements)

static int wgrad_tma2_launch(const ConvKP& p, float* dw, cudaStream_t st) {
  WgTma2P q;
  WgTmaP& t = q.b;
  t.n = p.n; t.ho = p.ho; t.wo = p.wo; t.cout = p.cout;
  t.kh = p.kh; t.kw = p.kw; t.stride = p.stride; t.pad = p.pad;
  t.c0 = p.c0; t.c1 = p.c1; t.K = p.K; t.bkc = 64; t.bnb = 64;
  const int bn = p.cout > 128 ? 256 : (p.cout > 64 ? 128 : 64);
  q.ktiles = ceil_div(p.K, TM);
  // k-tiles per CTA: fill the 512 TMEM columns, with the least padding of the last unit (ties: the larger)
  const int kt_max = 512 / bn < 4 ? 512 / bn : 4;
  int best_kt = 1, best_pad = 1 << 30;
  for (int kt = 1; kt <= kt_max; ++kt) {
    const int padded = ceil_div(q.ktiles, kt) * kt;
    if (padded <= best_pad) { best_pad = padded; best_kt = kt; }
  }
  q.kt = best_kt;
  const int stage64 = 64 * (q.kt * TM + bn) * 2;
  q.pb = 3 * stage64 <= 200 * 1024 ? 64 : 32;
  const int stage = q.pb * (q.kt * TM + bn) * 2;
  q.stages = 200 * 1024 / stage;
  if (q.stages > 8) q.stages = 8;
  int best_tw = 8;
  long best_cov = -1;
  for (int tw = 32; tw >= 8; tw >>= 1) {
    const int th = q.pb / tw;
    if (th < 1) continue;
    const long cov = (long)ceil_div(p.wo, tw) * tw * ceil_div(p.ho, th) * th;
    if (best_cov < 0 || cov < best_cov) { best_cov = cov; best_tw = tw; }
  }
  t.tw = best_tw; t.th = q.pb / best_tw;
  t.tiles_x = ceil_div(p.wo, t.tw); t.tiles_y = ceil_div(p.ho, t.th);
  t.num_ptiles = p.n * t.tiles_x * t.tiles_y;
  alignas(64) CUtensorMap a0, a1, dy;
  if (!make_act_map(&a0, p.src0, p.n, p.hin, p.win, p.c0, 64, t.tw, t.th, p.stride) ||
      !make_act_map(&a1, p.c1 > 0 ? p.src1 : p.src0, p.n, p.hin, p.win, p.c1 > 0 ? p.c1 : p.c0, 64, t.tw, t.th, p.stride) ||
      !make_act_map(&dy, p.dst, p.n, p.ho, p.wo, p.cout, 64, t.tw, t.th, 1)) {
    set_error("wgrad_tma2: cuTensorMapEncodeTiled failed");
    return RCFD_ECUDA;
  }
  const int units_k = ceil_div(q.ktiles, q.kt), units_n = ceil_div(p.cout, bn);
  switch (bn) {
    case 256: return launch_wg_tma2<256>(q, units_k, units_n, a0, a1, dy, dw, st);
    case 128: return launch_wg_tma2<128>(q, units_k, units_n, a0, a1, dy, dw, st);
    default: return launch_wg_tma2<64>(q, units_k, units_n, a0, a1, dy, dw, st);
  }
}

int wgrad_tma_launch(const ConvKP& p, float* dw, cudaStream_t st) {
  if (g_wgrad_tma_v2 && p.c0 % 64 == 0 && p.c1 % 64 == 0 && p.cout % 64 == 0 && p.K % 64 == 0 &&
      (reinterpret_cast<uintptr_t>(dw) & 15) == 0)
    return wgrad_tma2_launch(p, dw, st);
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
