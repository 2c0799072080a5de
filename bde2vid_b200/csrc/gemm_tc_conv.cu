// Persistent, fully TMA-fed tcgen05 convolution kernel (bf16 operands, fp32 accumulation in TMEM) for sm_100a.
//
//   C[m, n] = sum_k A[m, k] * W[n, k]     m = output pixel, k = (64-channel chunk, tap, channel in chunk)
//
// Used for the convolutions whose channel counts are multiples of 64 (ConvLSTM gates, decoders, deep encoders):
// both operands reach shared memory through the TMA, so no thread ever computes an im2col address.
//
//   * A operand.  The activation tensor [N, H, W, C] is described to the TMA as a 4-D tensor (C, d1, d2, N) where d1
//     is the pixel axis along which the 8 rows of a UMMA swizzle atom run (y by default: the strides of a tensor map
//     are free, so the map simply lists H before W).  An output tile is 8 (d1) x 16 (d2) pixels of one image.
//       - halo mode (stride 1): ONE box {64 ch, 16, 16 + k - 1} per 64-channel chunk brings the tile plus its halo
//         (zero filled outside the image by the TMA) into shared memory as [d2 column][16 pixels along d1][128 B].
//         The A tile of tap (k1, k2) is then nothing but a shifted view of that buffer: UMMA descriptor start =
//         halo + ((k2 * 16 + k1) * 128) bytes, atom stride (SBO) = 2048 bytes.  The swizzle phase of a 128-byte row
//         is a function of its shared-memory address, which the shift preserves (descriptor base offset 0; verified
//         on B200: setting the base-offset field to (start >> 7) & 7 gives wrong results).  Every input pixel is
//         fetched from L2 once per tile instead of k*k times.
//       - tap mode (stride 2): one box {64, 8, 16} per (chunk, tap), shifted and strided by the TMA itself
//         (elementStrides = 2).
//   * B operand: [BN x 64] weight tiles, 2-D TMA, SWIZZLE_128B (as in gemm_tc.cu).
//   * One CTA per SM loops over tiles; the TMEM accumulator is double buffered (2 x BN columns) so that the epilogue
//     of tile i (8 warps: tcgen05.ld -> smem transpose -> fused bias/activation/ConvLSTM gate math -> coalesced
//     stores) overlaps the main loop of tile i + 1.
//   * Pair mode (kPair): two CTAs of a cluster (one TPC) compute a 256 x BN tile with tcgen05.mma.cta_group::2.
//     Each CTA stages the A rows of its own 128 pixels and HALF of the weight tile; the leader CTA issues the MMAs
//     for both, every TMA load signals the leader's "full" barrier, and the leader's commits are multicast to the
//     "empty" / "accumulator ready" barriers of both CTAs.  Halves the shared-memory operand traffic per SM, which
//     is what bounds the single-CTA form (measured: 174 cycles per 128x256x16 MMA instead of 128).
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (owns TMEM), warps 2-9 = epilogue.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <tuple>

#include "tc_common.cuh"

namespace bde {
namespace tc {

constexpr int kCvThreads = 320;
constexpr int kCvEpiWarp0 = 2;
constexpr int kCvEpiWarps = 8;
constexpr int kCvD1 = 8, kCvD2 = 16;   // output tile: 8 pixels along the atom axis x 16 along the other one
constexpr int kCvHaloPitch = 16;       // pixels per halo column (8 + k - 1 <= 16)

// epilogue specialisations (compile-time: the generic one interprets the descriptor at run time)
constexpr int kEpiGeneric = 0;  // anything bde_gemm supports (fp32 output, residuals, every activation)
constexpr int kEpiLstm = 1;     // ConvLSTM gate math, c fp32 + h bf16
constexpr int kEpiStore = 2;    // bf16 store with none / ReLU / ReLU6, no residual

struct CvParams {
  TcParams p;
  int t1_tiles, t2_tiles;  // tiles along d1 / d2 per image
  int swap;                // 1: d1 = y, d2 = x;  0: d1 = x, d2 = y
  int num_m_tiles;
  int nchunks, ntaps;
  int halo_bytes;          // (16 + k - 1) * 2048
  int groups_per_n;        // ceil(num_m_tiles / CTAs per group): the CTAs of a pair take consecutive M tiles
  int num_groups;          // groups_per_n * (N / BN)
  float act_lo, act_hi;    // kEpiStore: clamp bounds (-inf / 0, 6 / +inf)
};

// NSUB = 2 ("dual tile", halo mode, narrow N): one CTA works on two neighbouring 8 x 16 tiles at once (one 8 x 32 halo
// load, two TMEM accumulators) and alternates its MMAs between them.  With N <= 128 a tcgen05.mma takes ~105 cycles
// however small N is -- consecutive MMAs into ONE accumulator are a dependent chain -- so two independent chains
// nearly double the issue rate, and every weight tile is fetched once for 256 pixels instead of 128.
template <int BN, bool kHalo, bool kPair, int NSUB = 1>
struct CvCfg {
  static constexpr int kABytes = kHalo ? (NSUB * kCvD2 + 4) * kCvHaloPitch * 128 : BM * 128;
  static constexpr int kBBytes = (kPair ? BN / 2 : BN) * 128;   // per CTA
  static constexpr int kEpiBytes = kCvEpiWarps * 4096;
  static constexpr int kBudget = 232448 - 1024 - 512 - kEpiBytes;
  static constexpr int kPairStages = (kBudget / (kABytes + kBBytes)) < 8 ? (kBudget / (kABytes + kBBytes)) : 8;
  static constexpr int kSA = kHalo ? 2 : kPairStages;
  static constexpr int kSBraw = kHalo ? (kBudget - 2 * kABytes) / kBBytes : kPairStages;
  static constexpr int kSB = kSBraw < 8 ? kSBraw : 8;
  static constexpr int kSmemBytes = kSA * kABytes + kSB * kBBytes + kEpiBytes + 1024 + 512;
  static constexpr int kAccCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int kTmemCols = 2 * NSUB * kAccCols;
  static_assert(kTmemCols <= 512, "TMEM budget");
  static_assert(kSA >= 2 && kSB >= 2, "pipeline too shallow");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// ---- PTX wrappers specific to this kernel ---------------------------------------------------------------------
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// pair mode: the destination is this CTA's shared memory, the mbarrier may live in the peer (leader) CTA
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2,
                                                 int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at the same offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// K-major SWIZZLE_128B operand whose 8-row atoms are `sbo` bytes apart
__device__ __forceinline__ uint64_t make_smem_desc_sbo(uint32_t smem_addr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor with an explicit M (256 for cta_group::2)
__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct CvTile {
  int img, u0, v0, n0;
  // group g: N tile g / groups_per_n, M tile (g % groups_per_n) * cs + rank (past the end: an all-zero dummy tile --
  // image index n_img is outside the tensor map, so the TMA fills zeros -- whose rows are never stored)
  __device__ __forceinline__ void init(const CvParams& cp, int group, int rank, int cs, int bn, int nsub = 1) {
    const int gn = group / cp.groups_per_n;
    const int m_tile = (group - gn * cp.groups_per_n) * cs + rank;
    n0 = gn * bn;
    const int per_img = cp.t1_tiles * cp.t2_tiles;
    img = m_tile / per_img;
    const int t = m_tile - img * per_img;
    const int t2 = t / cp.t1_tiles;
    u0 = (t - t2 * cp.t1_tiles) * kCvD1;
    v0 = t2 * kCvD2 * nsub;
  }
  // tile row r (TMEM lane) -> linear output pixel index, -1 outside the image
  __device__ __forceinline__ int row_m(const CvParams& cp, int r, int sub = 0) const {
    const int u = u0 + (r & 7), v = v0 + sub * kCvD2 + (r >> 3);
    const int y = cp.swap ? u : v, x = cp.swap ? v : u;
    return (y < cp.p.h_out && x < cp.p.w_out && img < cp.p.n_img) ? (img * cp.p.h_out + y) * cp.p.w_out + x : -1;
  }
};

template <int BN, bool kHalo, int EPI, bool kPair, int NSUB>
__global__ void __launch_bounds__(kCvThreads, 1)
conv_tma_kernel(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1,
                const __grid_constant__ CUtensorMap tmap_b, const CvParams cp) {
  using Cfg = CvCfg<BN, kHalo, kPair, NSUB>;
  constexpr int SA = Cfg::kSA, SB = Cfg::kSB;
  constexpr int CS = kPair ? 2 : 1;
  static_assert(NSUB == 1 || (kHalo && !kPair), "dual tiles need the halo form");
  const TcParams& p = cp.p;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_a + SA * Cfg::kABytes;
  const uint32_t epi_stage = smem_b + SB * Cfg::kBBytes;
  const uint32_t bar_base = epi_stage + Cfg::kEpiBytes;
  const uint32_t bar_afull = bar_base;                   // SA x 8
  const uint32_t bar_aempty = bar_afull + 8 * SA;        // SA x 8
  const uint32_t bar_bfull = bar_aempty + 8 * SA;        // SB x 8
  const uint32_t bar_bempty = bar_bfull + 8 * SB;        // SB x 8
  const uint32_t bar_tfull = bar_bempty + 8 * SB;        // 2 x 8
  const uint32_t bar_tempty = bar_tfull + 16;            // 2 x 8
  const uint32_t tmem_slot = bar_tempty + 16;            // 4
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kArriveWarps = BN >= 64 ? kCvEpiWarps : kCvEpiWarps / 2;
  const int rank = kPair ? (int)cluster_ctarank() : 0;
  const int group0 = kPair ? (int)cluster_id_x() : (int)blockIdx.x;
  const int group_step = kPair ? (int)num_clusters_x() : (int)gridDim.x;
  long long dbg_wait_acc = 0, dbg_wait_ops = 0, dbg_epi_wait = 0, dbg_epi_busy = 0, dbg_prod_wait = 0;
  const bool dbg = p.dbg != nullptr;
  const long long dbg_t0 = dbg ? clock64() : 0;

  pdl_trigger();
  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, kArriveWarps * CS);   // pair: the leader collects both CTAs' epilogue warps
    }
    fence_barrier_init();
    prefetch_tmap(&tmap_a0);
    prefetch_tmap(&tmap_a1);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();  // both CTAs' barriers exist before the peer signals them
  tcgen05_fence_after();
  const uint32_t tmem_acc = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  // programmatic dependent launch (common.cuh): the set-up above overlapped the previous kernel's tail; every thread
  // waits here, before the first activation load / residual read / output store
  pdl_wait();

  const int ks = p.ksize;
  if (warp == 0) {
    // =============================== TMA producer (all lanes wait, one elected lane issues) ===
    {
      uint32_t ia = 0, ib = 0;  // A / B stages issued so far
      // pair mode: "full" barriers live in the leader; its expect_tx covers the bytes of both CTAs
      const uint32_t afull0 = kPair ? mapa_cluster(bar_afull, 0) : bar_afull;
      const uint32_t bfull0 = kPair ? mapa_cluster(bar_bfull, 0) : bar_bfull;
      const uint32_t a_tx = (uint32_t)(kHalo ? cp.halo_bytes : BM * 128) * CS;
      for (int grp = group0; grp < cp.num_groups; grp += group_step) {
        CvTile tl;
        tl.init(cp, grp, rank, CS, BN, NSUB);
        const int o1 = tl.u0 * p.stride - p.pad, o2 = tl.v0 * p.stride - p.pad;  // input coordinate of tap (0, 0)
        const int nrow0 = tl.n0 + (kPair ? rank * (BN / 2) : 0);
        int kb = 0;
        for (int ch = 0; ch < cp.nchunks; ++ch) {
          const bool from0 = ch * BK < p.c0;
          const CUtensorMap* am = from0 ? &tmap_a0 : &tmap_a1;
          const int coff = from0 ? ch * BK : ch * BK - p.c0;
          if (kHalo) {
            const uint32_t s = ia % SA;
            mbar_wait(bar_aempty + 8 * s, ((ia / SA) & 1u) ^ 1u);
            if (elect_one_sync()) {
              if (!kPair || rank == 0) mbar_arrive_expect_tx(bar_afull + 8 * s, a_tx);
              if (kPair)
                tma_load_4d_pair(smem_a + s * Cfg::kABytes, am, afull0 + 8 * s, coff, o1, o2, tl.img);
              else
                tma_load_4d(smem_a + s * Cfg::kABytes, am, bar_afull + 8 * s, coff, o1, o2, tl.img);
            }
            __syncwarp();
            ++ia;
          }
          for (int ky = 0; ky < ks; ++ky)
            for (int kx = 0; kx < ks; ++kx, ++kb) {
              if (!kHalo) {
                const int k1 = cp.swap ? ky : kx, k2 = cp.swap ? kx : ky;
                const uint32_t s = ia % SA;
                mbar_wait(bar_aempty + 8 * s, ((ia / SA) & 1u) ^ 1u);
                if (elect_one_sync()) {
                  if (!kPair || rank == 0) mbar_arrive_expect_tx(bar_afull + 8 * s, a_tx);
                  if (kPair)
                    tma_load_4d_pair(smem_a + s * Cfg::kABytes, am, afull0 + 8 * s, coff, o1 + k1, o2 + k2, tl.img);
                  else
                    tma_load_4d(smem_a + s * Cfg::kABytes, am, bar_afull + 8 * s, coff, o1 + k1, o2 + k2, tl.img);
                }
                __syncwarp();
                ++ia;
              }
              const uint32_t s = ib % SB;
              const long long w0 = dbg ? clock64() : 0;
              mbar_wait(bar_bempty + 8 * s, ((ib / SB) & 1u) ^ 1u);
              if (dbg) dbg_prod_wait += clock64() - w0;
              if (elect_one_sync()) {
                if (!kPair || rank == 0) mbar_arrive_expect_tx(bar_bfull + 8 * s, (uint32_t)Cfg::kBBytes * CS);
                if (kPair)
                  tma_load_2d_pair(smem_b + s * Cfg::kBBytes, &tmap_b, bfull0 + 8 * s, kb * BK, nrow0);
                else
                  tma_load_2d(smem_b + s * Cfg::kBBytes, &tmap_b, bar_bfull + 8 * s, kb * BK, nrow0);
              }
              __syncwarp();
              ++ib;
            }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =============================== MMA issuer (pair mode: leader only) ================
    // The whole warp walks the loops (all lanes poll the barriers) and ONE elected lane issues the tcgen05
    // instructions: with a guard that comes from elect.sync ptxas emits plain UTCHMMA / UTCBAR, whereas under
    // `if (lane == 0)` it wraps every one of them in an ELECT ... BRA.U.ANY waterfall (~40 issue cycles per MMA, which
    // is what capped narrow-N tiles at ~105 cycles per instruction).
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_mn(kPair ? 256 : 128, BN);
      uint32_t ia = 0, ib = 0, lt = 0;
      for (int grp = group0; grp < cp.num_groups; grp += group_step, ++lt) {
        const uint32_t a = lt & 1u, aph = (lt >> 1) & 1u;
        const long long w0 = dbg ? clock64() : 0;
        mbar_wait(bar_tempty + 8 * a, aph ^ 1u);  // the epilogue(s) drained this accumulator slot
        if (dbg) dbg_wait_acc += clock64() - w0;
        tcgen05_fence_after();
        const uint32_t acc = tmem_acc + a * (NSUB * Cfg::kAccCols);
        uint32_t first = 1u;
        for (int ch = 0; ch < cp.nchunks; ++ch) {
          uint32_t sa = 0;
          if (kHalo) {
            sa = ia % SA;
            const long long w1 = dbg ? clock64() : 0;
            mbar_wait(bar_afull + 8 * sa, (ia / SA) & 1u);
            if (dbg) dbg_wait_ops += clock64() - w1;
            ++ia;
          }
          for (int ky = 0; ky < ks; ++ky)
            for (int kx = 0; kx < ks; ++kx) {
              uint32_t a_start;
              if (kHalo) {
                const int k1 = cp.swap ? ky : kx, k2 = cp.swap ? kx : ky;
                a_start = smem_a + sa * Cfg::kABytes + (uint32_t)(k2 * kCvHaloPitch + k1) * 128u;
              } else {
                sa = ia % SA;
                mbar_wait(bar_afull + 8 * sa, (ia / SA) & 1u);
                ++ia;
                a_start = smem_a + sa * Cfg::kABytes;
              }
              const uint32_t sb = ib % SB;
              const long long w2 = dbg ? clock64() : 0;
              mbar_wait(bar_bfull + 8 * sb, (ib / SB) & 1u);
              if (dbg) dbg_wait_ops += clock64() - w2;
              ++ib;
              tcgen05_fence_after();
              const bool last_tap = ky == ks - 1 && kx == ks - 1;
              const bool last_kb = last_tap && ch == cp.nchunks - 1;
              if (elect_one_sync()) {
                const uint64_t adesc = make_smem_desc_sbo(a_start, kHalo ? kCvHaloPitch * 128u : 1024u);
                const uint64_t bdesc = make_smem_desc(smem_b + sb * Cfg::kBBytes);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  if (kPair) {
                    umma_bf16_pair(acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (first && k == 0) ? 0u : 1u);
                  } else {
#pragma unroll
                    for (int sub = 0; sub < NSUB; ++sub)   // sub-tile `sub`: 16 halo columns further, its own accumulator
                      umma_bf16(acc + sub * Cfg::kAccCols, adesc + (uint64_t)(2 * k) + (uint64_t)(sub * ((kCvD2 * kCvHaloPitch * 128) >> 4)),
                                bdesc + (uint64_t)(2 * k), idesc, (first && k == 0) ? 0u : 1u);
                  }
                }
                if (kPair) {
                  umma_commit_pair(bar_bempty + 8 * sb);
                  if (!kHalo || last_tap) umma_commit_pair(bar_aempty + 8 * sa);
                  if (last_kb) umma_commit_pair(bar_tfull + 8 * a);
                } else {
                  umma_commit(bar_bempty + 8 * sb);
                  if (!kHalo || last_tap) umma_commit(bar_aempty + 8 * sa);
                  if (last_kb) umma_commit(bar_tfull + 8 * a);
                }
              }
              __syncwarp();
              first = 0u;
            }
        }
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue warps ======================================
    const int ew = warp - kCvEpiWarp0;
    const int q = warp & 3;      // TMEM lane quarter this warp may access
    const int half = ew >> 2;    // which of the two interleaved 32-column chunk sets
    if (half * 32 < BN) {
      float* stg = reinterpret_cast<float*>(smem_gen + (epi_stage - smem_base)) + ew * 1024;
      const uint32_t tempty0 = kPair ? mapa_cluster(bar_tempty, 0) : bar_tempty;
      constexpr int NCH = (BN + 63) / 64;   // chunks per warp
      uint32_t lt = 0;
      for (int grp = group0; grp < cp.num_groups; grp += group_step, ++lt) {
        const uint32_t a = lt & 1u, aph = (lt >> 1) & 1u;
        CvTile tl;
        tl.init(cp, grp, rank, CS, BN, NSUB);
        long long w1 = 0;
#pragma unroll 1
        for (int sub = 0; sub < NSUB; ++sub) {
        const uint32_t lane_taddr = tmem_acc + (a * NSUB + sub) * Cfg::kAccCols + ((uint32_t)(q * 32) << 16);
        if (EPI == kEpiGeneric) {
          const int rq = lane >> 3, cq4 = lane & 7, cq = cq4 * 4;
          const bool need_aux = p.epi != BDE_EPI_STORE || p.residual != nullptr;
          int m_it[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) m_it[it] = tl.row_m(cp, q * 32 + it * 4 + rq, sub);
          float4 aux[8];
#pragma unroll
          for (int it = 0; it < 8; ++it)
            aux[it] = (need_aux && m_it[it] >= 0) ? aux_load(p, m_it[it], tl.n0 + half * 32 + cq, -1) : make_float4(0.f, 0.f, 0.f, 0.f);
          const long long w0 = dbg ? clock64() : 0;
          mbar_wait(bar_tfull + 8 * a, aph);
          if (sub == 0) {
            w1 = dbg ? clock64() : 0;
            dbg_epi_wait += w1 - w0;
          }
          tcgen05_fence_after();
#pragma unroll 1
          for (int cb = half * 32; cb < BN; cb += 64) {
            uint32_t raw[32];
            tmem_ld_32x32b_x32(lane_taddr + (uint32_t)cb, raw);
            tmem_ld_wait();
            if (cb + 64 >= BN && sub == NSUB - 1) {
              // last TMEM read of this tile by this warp: hand the accumulator slot back to the MMA warp
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (kPair) mbar_arrive_cluster(tempty0 + 8 * a); else mbar_arrive(bar_tempty + 8 * a);
              }
            }
            __syncwarp();  // previous chunk fully read back before it is overwritten
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
              *reinterpret_cast<float4*>(stg + lane * 32 + ((j4 ^ (lane & 7)) << 2)) =
                  make_float4(__uint_as_float(raw[4 * j4]), __uint_as_float(raw[4 * j4 + 1]), __uint_as_float(raw[4 * j4 + 2]),
                              __uint_as_float(raw[4 * j4 + 3]));
            __syncwarp();
            float4 accv[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int r = it * 4 + rq;
              accv[it] = *reinterpret_cast<const float4*>(stg + r * 32 + ((cq4 ^ (r & 7)) << 2));
            }
            float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias != nullptr) bq = __ldg(reinterpret_cast<const float4*>(p.bias + tl.n0 + cb + cq));
#pragma unroll
            for (int it = 0; it < 8; ++it)
              if (m_it[it] >= 0) epilogue_quad_aux(p, m_it[it], tl.n0 + cb + cq, accv[it], bq, -1, aux[it]);
            if (need_aux && cb + 64 < BN) {
#pragma unroll
              for (int it = 0; it < 8; ++it)
                if (m_it[it] >= 0) aux[it] = aux_load(p, m_it[it], tl.n0 + cb + 64 + cq, -1);
            }
          }
        } else {
          // ---- specialised epilogues: two lanes per tile row, 16 consecutive columns each -----------------------
          // pass ps covers rows ps*16 + lane/2 of this warp's 32; hf = lane & 1 selects columns [16 hf, 16 hf + 16)
          const int hf = lane & 1;
          int m_ps[2];
#pragma unroll
          for (int ps = 0; ps < 2; ++ps) m_ps[ps] = tl.row_m(cp, q * 32 + ps * 16 + (lane >> 1), sub);
          const int hid = p.N >> 2;
          // LSTM: c_prev of every chunk of this warp, in flight while the accumulator is still being computed
          float4 cpv[NCH][2];
          if (EPI == kEpiLstm) {
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
              for (int ps = 0; ps < 2; ++ps) {
                const int cb = half * 32 + c * 64;
                cpv[c][ps] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.c_prev != nullptr && m_ps[ps] >= 0 && cb < BN)
                  cpv[c][ps] = __ldg(reinterpret_cast<const float4*>(p.c_prev + (size_t)m_ps[ps] * hid + ((tl.n0 + cb) >> 2) + hf * 4));
              }
          }
          const long long w0 = dbg ? clock64() : 0;
          mbar_wait(bar_tfull + 8 * a, aph);
          if (sub == 0) {
            w1 = dbg ? clock64() : 0;
            dbg_epi_wait += w1 - w0;
          }
          tcgen05_fence_after();
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            const int cb = half * 32 + c * 64;
            if (cb < BN) {
              uint32_t raw[32];
              tmem_ld_32x32b_x32(lane_taddr + (uint32_t)cb, raw);
              // bias of this lane's 16 columns (L1-resident) while the TMEM load is in flight
              float4 bv[4];
#pragma unroll
              for (int k = 0; k < 4; ++k)
                bv[k] = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + tl.n0 + cb + hf * 16 + 4 * k))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
              tmem_ld_wait();
              if (cb + 64 >= BN && sub == NSUB - 1) {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) {
                  if (kPair) mbar_arrive_cluster(tempty0 + 8 * a); else mbar_arrive(bar_tempty + 8 * a);
                }
              }
              __syncwarp();
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4)
                *reinterpret_cast<float4*>(stg + lane * 32 + ((j4 ^ (lane & 7)) << 2)) =
                    make_float4(__uint_as_float(raw[4 * j4]), __uint_as_float(raw[4 * j4 + 1]), __uint_as_float(raw[4 * j4 + 2]),
                                __uint_as_float(raw[4 * j4 + 3]));
              __syncwarp();
#pragma unroll
              for (int ps = 0; ps < 2; ++ps) {
                const int r = ps * 16 + (lane >> 1);
                float4 v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  v[k] = *reinterpret_cast<const float4*>(stg + r * 32 + (((hf * 4 + k) ^ (r & 7)) << 2));
                  v[k].x += bv[k].x; v[k].y += bv[k].y; v[k].z += bv[k].z; v[k].w += bv[k].w;
                }
                const int m = m_ps[ps];
                if (EPI == kEpiLstm) {
                  // quad k = (in, remember, out, cell) of hidden channel (n0 + cb) / 4 + 4 hf + k  (submodules.py:320-332)
                  const float cp4[4] = {cpv[c][ps].x, cpv[c][ps].y, cpv[c][ps].z, cpv[c][ps].w};
                  float cc[4], hh[4];
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    cc[k] = fmaf(mufu_sigmoid(v[k].y), cp4[k], mufu_sigmoid(v[k].x) * mufu_tanh(v[k].w));
                    hh[k] = mufu_sigmoid(v[k].z) * mufu_tanh(cc[k]);
                  }
                  if (m >= 0) {
                    const size_t o = (size_t)m * hid + ((tl.n0 + cb) >> 2) + hf * 4;
                    *reinterpret_cast<float4*>(p.c_out + o) = make_float4(cc[0], cc[1], cc[2], cc[3]);
                    const __nv_bfloat162 h0 = __floats2bfloat162_rn(hh[0], hh[1]), h1 = __floats2bfloat162_rn(hh[2], hh[3]);
                    uint2 pk;
                    pk.x = *reinterpret_cast<const uint32_t*>(&h0);
                    pk.y = *reinterpret_cast<const uint32_t*>(&h1);
                    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o) = pk;
                  }
                } else {
                  float o16[16];
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    o16[4 * k + 0] = fminf(fmaxf(v[k].x, cp.act_lo), cp.act_hi);
                    o16[4 * k + 1] = fminf(fmaxf(v[k].y, cp.act_lo), cp.act_hi);
                    o16[4 * k + 2] = fminf(fmaxf(v[k].z, cp.act_lo), cp.act_hi);
                    o16[4 * k + 3] = fminf(fmaxf(v[k].w, cp.act_lo), cp.act_hi);
                  }
                  if (m >= 0) {
                    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)m * p.N + tl.n0 + cb + hf * 16);
                    dst[0] = pack8_bf16(o16);
                    dst[1] = pack8_bf16(o16 + 8);
                  }
                }
              }
            }
          }
        }
        }  // sub
        if (dbg) dbg_epi_busy += clock64() - w1;
      }
    }
  }
  if (dbg) {
    long long* o = p.dbg + (size_t)blockIdx.x * 8;
    if (threadIdx.x == 0) o[5] = dbg_prod_wait;
    if (threadIdx.x == 32) { o[1] = dbg_wait_acc; o[2] = dbg_wait_ops; }
    if (threadIdx.x == 64) { o[3] = dbg_epi_wait; o[4] = dbg_epi_busy; o[0] = clock64() - dbg_t0; }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();  // no CTA leaves while the peer may still signal its barriers or read its operands
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_acc, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn();

// 4-D activation map (C, d1, d2, N) over an NHWC tensor; box {64, box1, box2, 1}; element stride es along d1 / d2
// `pitch` = elements between consecutive pixels (>= C: the source may be a channel slice of a wider map)
static int get_act_tmap(const void* base, int C, int pitch, int H, int W, int N, int swap, int box1, int box2, int es, CUtensorMap* out) {
  static std::mutex mu;
  static std::map<std::tuple<const void*, int, int, int, int, int, int, int, int, int>, CUtensorMap> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_tuple(base, C, pitch, H, W, N, swap, box1, box2, es);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return 0;
  }
  auto encode = get_encode_fn();
  BDE_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t row = (cuuint64_t)pitch * 2, line = (cuuint64_t)W * pitch * 2, image = (cuuint64_t)H * W * pitch * 2;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)(swap ? H : W), (cuuint64_t)(swap ? W : H), (cuuint64_t)N};
  cuuint64_t gstride[3] = {swap ? line : row, swap ? row : line, image};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)box1, (cuuint32_t)box2, 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
  CUtensorMap m;
  CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BDE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (activation, swap=%d, box %dx%d, stride %d) failed (%d)", swap, box1, box2,
              es, (int)r);
  if (cache.size() > 16384) cache.clear();
  cache[key] = m;
  *out = m;
  return 0;
}

template <int BN, bool kHalo, int EPI, bool kPair, int NSUB = 1>
static int launch_conv(const CUtensorMap& ta0, const CUtensorMap& ta1, const CUtensorMap& tb, const CvParams& cp, cudaStream_t s) {
  using Cfg = CvCfg<BN, kHalo, kPair, NSUB>;
  constexpr int CS = kPair ? 2 : 1;
  auto kern = conv_tma_kernel<BN, kHalo, EPI, kPair, NSUB>;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[2];
  unsigned n_attr = 0;
  if (kPair) {
    attr[n_attr].id = cudaLaunchAttributeClusterDimension;
    attr[n_attr].val.clusterDim.x = CS;
    attr[n_attr].val.clusterDim.y = 1;
    attr[n_attr].val.clusterDim.z = 1;
    ++n_attr;
  }
  if (pdl_enabled()) {
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  cfg.blockDim = dim3(kCvThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = s;
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  if (first_use_on_device((const void*)kern)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    BDE_REQUIRE(e == cudaSuccess, "bde_gemm(tcgen05 conv): smem attribute: %s", cudaGetErrorString(e));
  }
  const int n_sm = device_sm_count();
  int max_clusters = n_sm / CS;     // persistent: one CTA (or CTA pair) per SM
  if (kPair) {
    cfg.gridDim = dim3(n_sm / CS * CS);
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    BDE_REQUIRE(e == cudaSuccess && n > 0, "bde_gemm(tcgen05 conv): cluster occupancy query: %s", cudaGetErrorString(e));
    if (n < max_clusters) max_clusters = n;
  }
  const int clusters = cp.num_groups < max_clusters ? cp.num_groups : max_clusters;
  cfg.gridDim = dim3((unsigned)(clusters * CS));
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta0, ta1, tb, cp);
  BDE_REQUIRE(e == cudaSuccess, "bde_gemm(tcgen05 conv): launch: %s", cudaGetErrorString(e));
  return check_launch("conv_tma_kernel");
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e != nullptr && e[0] != 0) ? atoi(e) : dflt;
}

// 1 if this problem is served by the TMA convolution kernel
bool conv_tma_eligible(const TcParams& p, bool ln) {
  if (ln || !env_flag("BDE2VID_CONV_TMA", true)) return false;
  // 3x3 .. 5x5 with chunk-major K, or a 1x1 convolution (any K order: one tap) -- a per-pixel linear layer on an NHWC map
  if (p.epi == BDE_EPI_SCATTER) return false;
  if (p.ksize == 1) {
    if (p.stride != 1 || p.pad != 0 || p.out_f32 || p.residual != nullptr || !env_flag("BDE2VID_CONV_1X1", true)) return false;
  } else if (p.ksize < 3 || p.ksize > 5 || p.k_order != 1) {
    return false;
  }
  if (p.c0 % BK != 0 || p.c1 % BK != 0) return false;
  if (p.stride == 2) {
    if (!env_flag("BDE2VID_CONV_S2", true)) return false;
  } else if (p.stride != 1) {
    return false;
  }
  if (p.h_out < 8 || p.w_out < 8) return false;
  return p.N % 32 == 0;
}

template <int BN, bool kHalo, bool kPair, int NSUB>
static int launch_conv_epi(int epi, const CUtensorMap& ta0, const CUtensorMap& ta1, const CUtensorMap& tb, const CvParams& cp,
                           cudaStream_t s) {
  if (epi == kEpiLstm) return launch_conv<BN, kHalo, kEpiLstm, kPair, NSUB>(ta0, ta1, tb, cp, s);
  if (epi == kEpiStore) return launch_conv<BN, kHalo, kEpiStore, kPair, NSUB>(ta0, ta1, tb, cp, s);
  return launch_conv<BN, kHalo, kEpiGeneric, kPair, NSUB>(ta0, ta1, tb, cp, s);
}

template <int BN>
static int launch_conv_bn(bool halo, bool pair, int nsub, int epi, const CUtensorMap& ta0, const CUtensorMap& ta1, const CUtensorMap& tb,
                          const CvParams& cp, cudaStream_t s) {
  if (halo) {
    // the pair form (cta_group::2) is kept for the widest tile only, the dual-tile form for N tiles up to 128
    if constexpr (BN == 256) {
      if (pair) return launch_conv_epi<BN, true, true, 1>(epi, ta0, ta1, tb, cp, s);
    } else {
      if (nsub == 2) return launch_conv_epi<BN, true, false, 2>(epi, ta0, ta1, tb, cp, s);
    }
    return launch_conv_epi<BN, true, false, 1>(epi, ta0, ta1, tb, cp, s);
  }
  return launch_conv_epi<BN, false, false, 1>(epi, ta0, ta1, tb, cp, s);
}

int conv_tma_launch(const bde_gemm_desc* d, const TcParams& p0, cudaStream_t s) {
  CvParams cp;
  memset(&cp, 0, sizeof(cp));
  cp.p = p0;
  const TcParams& p = cp.p;
  const bool halo = p.stride == 1 && env_flag("BDE2VID_CONV_HALO", true);
  cp.swap = env_int("BDE2VID_CONV_SWAP", 1) ? 1 : 0;
  const int d1_out = cp.swap ? p.h_out : p.w_out, d2_out = cp.swap ? p.w_out : p.h_out;
  // widest N tile that divides N
  const int bn_max = (p.N % 256 == 0) ? 256 : (p.N % 128 == 0) ? 128 : (p.N % 64 == 0) ? 64 : 32;
  // dual tiles (two 8 x 16 tiles per CTA) for narrow N, where single MMAs are latency-bound; needs enough tiles to
  // keep every SM busy and a map at least two tiles wide
  int nsub = 1;
  if (halo && bn_max <= 128 && env_flag("BDE2VID_CONV_DUAL", true) && d2_out > kCvD2 &&
      (size_t)p.n_img * ceil_div(d1_out, kCvD1) * ceil_div(d2_out, 2 * kCvD2) * (p.N / bn_max) >= (size_t)device_sm_count())
    nsub = 2;
  cp.t1_tiles = (int)ceil_div(d1_out, kCvD1);
  cp.t2_tiles = (int)ceil_div(d2_out, kCvD2 * nsub);
  cp.num_m_tiles = p.n_img * cp.t1_tiles * cp.t2_tiles;
  cp.ntaps = p.ksize * p.ksize;
  cp.nchunks = p.ctot / BK;
  cp.halo_bytes = (kCvD2 * nsub + p.ksize - 1) * kCvHaloPitch * 128;
  // CTA pairs (cta_group::2): off by default -- measured on B200 the pair form is 3-8 % SLOWER than one CTA per SM on
  // every shape of the path (the main loop is bound by the tensor pipe at ~700 cycles per 128x256x64 block either way)
  const bool pair = env_flag("BDE2VID_CONV_PAIR", false) && cp.num_m_tiles >= 2 && halo && bn_max == 256 && nsub == 1;
  const int cs = pair ? 2 : 1;
  const int units = device_sm_count() / cs;   // CTAs or CTA pairs that run concurrently
  // tile width: the widest N tile that divides N, narrowed only while even the narrower tiles leave half of the SMs idle.
  // A narrow tile re-streams the activations per N tile and issues MMAs of half the width: measured on the ConvLSTM gate
  // convolutions of one sequence (tools/conv_probe.py, us per launch for BN = 64 / 128 / 256): level 2 (110 tiles of 256)
  // 36.9 / 29.2 / 20.8, level 3 (60 tiles of 256) 46.2 / 29.1 / 32.0; two sequences, level 3: 85.8 / 52.9 / 33.9.
  int bn = bn_max;
  while (nsub == 1 && !pair && bn > 64 && 2 * ceil_div(cp.num_m_tiles, cs) * (size_t)(p.N / bn) <= (size_t)units) bn /= 2;
  {
    const int f = env_int("BDE2VID_CONV_BN", 0);
    if ((f == 32 || f == 64 || f == 128 || f == 256) && p.N % f == 0 && nsub == 1 && !pair) bn = f;
  }
  cp.groups_per_n = (int)ceil_div(cp.num_m_tiles, cs);
  const size_t groups = (size_t)cp.groups_per_n * (p.N / bn);
  BDE_REQUIRE(groups < ((size_t)1 << 30), "bde_gemm(tcgen05 conv): too many tiles");
  cp.num_groups = (int)groups;
  if (g_dbg != nullptr && (size_t)device_sm_count() <= g_dbg_ctas) cp.p.dbg = g_dbg;
  // epilogue specialisation
  int epi = kEpiGeneric;
  if (env_flag("BDE2VID_CONV_EPI_SPEC", true)) {
    if (p.epi == BDE_EPI_LSTM) {
      epi = kEpiLstm;
    } else if (p.epi == BDE_EPI_STORE && !p.out_f32 && p.residual == nullptr &&
               (p.act == BDE_ACT_NONE || p.act == BDE_ACT_RELU || p.act == BDE_ACT_RELU6)) {
      epi = kEpiStore;
      cp.act_lo = p.act == BDE_ACT_NONE ? -INFINITY : 0.0f;
      cp.act_hi = p.act == BDE_ACT_RELU6 ? 6.0f : INFINITY;
    }
  }
  CUtensorMap ta0, ta1, tb;
  const int box1 = halo ? kCvHaloPitch : (p.stride == 2 ? 2 * kCvD1 : kCvD1);
  const int box2 = halo ? kCvD2 * nsub + p.ksize - 1 : (p.stride == 2 ? 2 * kCvD2 : kCvD2);
  const int ld0 = d->a0_ld > 0 ? d->a0_ld : p.c0, ld1 = d->a1_ld > 0 ? d->a1_ld : p.c1;
  BDE_REQUIRE(ld0 % 8 == 0 && ld1 % 8 == 0, "bde_gemm(tcgen05 conv): pixel pitches must be multiples of 8 elements");
  int rc = get_act_tmap(d->a0, p.c0, ld0, p.h_in, p.w_in, p.n_img, cp.swap, box1, box2, p.stride, &ta0);
  if (rc != 0) return rc;
  if (p.c1 > 0) {
    rc = get_act_tmap(d->a1, p.c1, ld1, p.h_in, p.w_in, p.n_img, cp.swap, box1, box2, p.stride, &ta1);
    if (rc != 0) return rc;
  } else {
    ta1 = ta0;
  }
  rc = get_weight_tmap(d->w, p.N, p.w_ld, bn / cs, &tb);
  if (rc != 0) return rc;
  switch (bn) {
    case 32: return launch_conv_bn<32>(halo, pair, nsub, epi, ta0, ta1, tb, cp, s);
    case 64: return launch_conv_bn<64>(halo, pair, nsub, epi, ta0, ta1, tb, cp, s);
    case 128: return launch_conv_bn<128>(halo, pair, nsub, epi, ta0, ta1, tb, cp, s);
    default: return launch_conv_bn<256>(halo, pair, nsub, epi, ta0, ta1, tb, cp, s);
  }
}

}  // namespace tc
}  // namespace bde
