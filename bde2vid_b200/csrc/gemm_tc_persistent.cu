// Persistent tcgen05 implicit-GEMM kernel (bf16 operands, fp32 accumulation in TMEM) for sm_100a.
//
// One CTA per SM loops over output tiles (128 x BN).  Compared with the one-tile-per-CTA kernel in
// gemm_tc.cu the per-CTA setup (barrier init, TMEM allocation, tensor-map fetch) is paid once, the
// operand pipeline runs ahead across tile boundaries, and the epilogue of tile i overlaps the main
// loop of tile i+1 through a double-buffered TMEM accumulator.  Warp roles (448 threads):
//   warps 0-7    A producers (cp.async im2col gather, or the LayerNorm-gather producer for kLn > 0)
//   warp  8      B producer: TMA 2-D tiled loads of the [BN x 64] weight tile (SWIZZLE_128B)
//   warp  9      MMA issuer: tcgen05.mma M128 x BN x K16, commits to the stage / accumulator barriers
//   warps 10-13  epilogue: TMEM -> registers -> shared-memory transpose -> fused quad epilogue ->
//                coalesced global stores; releases the accumulator slot as soon as it is drained
// Barriers: full/empty per smem stage, tmem_full/tmem_empty per accumulator slot.
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace bde {
namespace tc {

constexpr int kEpiWarp0 = kNumProducerWarps + 2;  // 10
constexpr int kNumEpiWarps = 4;
constexpr int kPThreads = (kEpiWarp0 + kNumEpiWarps) * 32;  // 448
constexpr int kStagePitch = 36;                              // floats per staged row (see gemm_tc.cu)

template <int BN, int kLn>
struct PCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BN <= 64 ? 8 : (BN <= 128 ? 6 : 4);
  static constexpr int kEpiStageBytes = kNumEpiWarps * 32 * kStagePitch * 4;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiStageBytes + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int kAccCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int kTmemCols = 2 * kAccCols;  // two accumulator slots
  static_assert(kLn <= kStages, "LayerNorm producer needs all K blocks of a tile resident");
};

template <int BN, bool kBTma, int kLn>
__global__ void __launch_bounds__(kPThreads, 1)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmap_b, const TcParams p, int num_m_tiles, int num_tiles) {
  using Cfg = PCfg<BN, kLn>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + S * Cfg::kABytes;
  const uint32_t epi_stage = smem_base + S * Cfg::kStageBytes;
  const uint32_t bar_base = epi_stage + Cfg::kEpiStageBytes;
  const uint32_t bar_full = bar_base;                  // S x 8B
  const uint32_t bar_empty = bar_base + 8 * S;         // S x 8B
  const uint32_t bar_tfull = bar_base + 16 * S;        // 2 x 8B
  const uint32_t bar_tempty = bar_base + 16 * S + 16;  // 2 x 8B
  const uint32_t tmem_slot = bar_base + 16 * S + 32;   // 4B
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = p.num_kb;

  if (threadIdx.x == 0) {
    const uint32_t full_count = kNumProducerThreads + (kBTma ? 1 : 32);
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, full_count);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, kNumEpiWarps * 32);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_acc = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp < kNumProducerWarps) {
    // =============================== A producers ========================================
    const int j = lane & 7;
    const int rsub = lane >> 3;
    const int ntaps = p.ksize * p.ksize;
    uint32_t dsto[kRowsPerThread];
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
      const int row = warp * (4 * kRowsPerThread) + i * 4 + rsub;
      dsto[i] = (uint32_t)row * 128u + (((uint32_t)j ^ (uint32_t)(row & 7)) << 4);
    }
    uint32_t it = 0;  // K blocks issued so far (stage = it % S, phase = (it / S) & 1)
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      TileGeom geom;
      geom.init(p, tile % num_m_tiles);
      if (kLn > 0) {
        // ---- LayerNorm-gather producer: fp32 rows -> normalised bf16 tile, all K blocks of the tile ----
        constexpr int nchunk = kLn > 0 ? kLn : 1;
        constexpr int C = 64 * nchunk;
        for (int kb = 0; kb < nchunk; ++kb) {
          const uint32_t itk = it + kb;
          mbar_wait(bar_empty + 8 * (itk % S), ((itk / S) & 1u) ^ 1u);
        }
#pragma unroll(kLn <= 2 ? 4 : 2)
        for (int i = 0; i < kRowsPerThread; ++i) {
          const int row = warp * (4 * kRowsPerThread) + i * 4 + rsub;
          const int mm = geom.m0 + row;
          const float* src = nullptr;
          if (mm < p.M) {
            int pix = mm, d = 0;
            if (p.ln_map != nullptr) {
              const int tok = mm % p.ln_ntok;
              const int r2 = mm / p.ln_ntok;
              d = r2 % p.ln_D;
              pix = __ldg(p.ln_map + (r2 / p.ln_D) * p.ln_ntok + tok);
            }
            const float* fr = p.ln_f[0];
#pragma unroll
            for (int t = 1; t < 8; ++t) fr = (d == t) ? p.ln_f[t] : fr;
            if (fr != nullptr && pix >= 0) src = fr + (size_t)pix * C + j * 8;
          }
          float v[nchunk][8];
          float sum = 0.f;
#pragma unroll
          for (int kb = 0; kb < nchunk; ++kb) {
            if (src != nullptr) {
              const float4 t0 = *reinterpret_cast<const float4*>(src + kb * 64), t1 = *reinterpret_cast<const float4*>(src + kb * 64 + 4);
              v[kb][0] = t0.x; v[kb][1] = t0.y; v[kb][2] = t0.z; v[kb][3] = t0.w;
              v[kb][4] = t1.x; v[kb][5] = t1.y; v[kb][6] = t1.z; v[kb][7] = t1.w;
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[kb][e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) sum += v[kb][e];
          }
          sum += __shfl_xor_sync(0xffffffffu, sum, 1);
          sum += __shfl_xor_sync(0xffffffffu, sum, 2);
          sum += __shfl_xor_sync(0xffffffffu, sum, 4);
          const float mean = sum / (float)C;
          float sq = 0.f;
#pragma unroll
          for (int kb = 0; kb < nchunk; ++kb)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float dlt = v[kb][e] - mean;
              v[kb][e] = dlt;
              sq += dlt * dlt;
            }
          sq += __shfl_xor_sync(0xffffffffu, sq, 1);
          sq += __shfl_xor_sync(0xffffffffu, sq, 2);
          sq += __shfl_xor_sync(0xffffffffu, sq, 4);
          const float rstd = 1.0f / sqrtf(sq / (float)C + 1e-5f);
#pragma unroll
          for (int kb = 0; kb < nchunk; ++kb) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = v[kb][e] * rstd;
            const uint4 pk = pack8_bf16(o);
            const uint32_t dst = smem_a + ((it + kb) % S) * Cfg::kABytes + dsto[i];
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
          }
        }
        fence_proxy_async_smem();
        for (int kb = 0; kb < nchunk; ++kb) mbar_arrive(bar_full + 8 * ((it + kb) % S));
        it += nchunk;
      } else {
        int pix0[kRowsPerThread];
        uint32_t vmask[kRowsPerThread];
#pragma unroll
        for (int i = 0; i < kRowsPerThread; ++i) {
          const int row = warp * (4 * kRowsPerThread) + i * 4 + rsub;
          pix0[i] = 0;
          vmask[i] = 0u;
          int img, oy, ox, mm;
          if (!geom.row_pixel(p, row, img, oy, ox, mm)) continue;
          if (p.dense) {
            pix0[i] = mm;
            vmask[i] = 1u;
          } else {
            const int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
            pix0[i] = (img * p.h_in + iy0) * p.w_in + ix0;
            const int kx_lo = max(0, -ix0), kx_hi = min(p.ksize, p.w_in - ix0);
            const int ky_lo = max(0, -iy0), ky_hi = min(p.ksize, p.h_in - iy0);
            uint32_t msk = 0u;
            if (kx_hi > kx_lo && ky_hi > ky_lo) {
              const uint32_t xm = ((1u << (kx_hi - kx_lo)) - 1u) << kx_lo;
              const uint32_t rows = ((1u << (ky_hi * p.ksize)) - 1u) & ~((1u << (ky_lo * p.ksize)) - 1u);
              msk = xm * (p.ypat_all & rows);
            }
            vmask[i] = msk;
          }
        }
        int tap, c;
        if (p.k_order == 1) {
          tap = 0;
          c = j * 8;
        } else {
          tap = (j * 8) / p.ctot;
          c = (j * 8) - tap * p.ctot;
        }
        int ky = tap / p.ksize, kx = tap - ky * p.ksize;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % S;
          const uint32_t ph = (it / S) & 1u;
          const bool from0 = c < p.c0;
          const __nv_bfloat16* sbase = from0 ? p.a0 + c : p.a1 + (c - p.c0);
          const int cs = from0 ? p.c0 : p.c1;
          const int tapoff = ky * p.w_in + kx;
          const uint32_t tapbit = (tap < ntaps && c < p.ctot) ? (1u << tap) : 0u;
          mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          const uint32_t stage_a = smem_a + s * Cfg::kABytes;
#pragma unroll
          for (int i = 0; i < kRowsPerThread; ++i) {
            const bool ok = (vmask[i] & tapbit) != 0u;
            const __nv_bfloat16* src = ok ? sbase + (size_t)(pix0[i] + tapoff) * cs : p.a0;
            cp_async_16(stage_a + dsto[i], src, ok ? 16u : 0u);
          }
          cp_async_mbar_arrive_noinc(bar_full + 8 * s);
          if (p.k_order == 1) {
            ++tap;
            if (++kx == p.ksize) {
              kx = 0;
              if (++ky == p.ksize) {
                ky = 0;
                tap = 0;
                c += BK;
              }
            }
          } else {
            c += BK;
            while (c >= p.ctot) {
              c -= p.ctot;
              ++tap;
              if (++kx == p.ksize) {
                kx = 0;
                ++ky;
              }
            }
          }
        }
      }
    }
  } else if (warp == kTmaWarp) {
    // =============================== B producer =========================================
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n0 = (tile / num_m_tiles) * BN;
      if (kBTma) {
        if (lane == 0) {
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const uint32_t s = it % S;
            mbar_wait(bar_empty + 8 * s, ((it / S) & 1u) ^ 1u);
            mbar_arrive_expect_tx(bar_full + 8 * s, Cfg::kBBytes);
            tma_load_2d(smem_b + s * Cfg::kBBytes, &tmap_b, bar_full + 8 * s, kb * BK, n0);
          }
        }
      } else {
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % S;
          mbar_wait(bar_empty + 8 * s, ((it / S) & 1u) ^ 1u);
          for (int row = lane >> 3; row < BN; row += 4) {
            const int jj = lane & 7;
            const __nv_bfloat16* src = p.w + (size_t)(n0 + row) * p.w_ld + kb * BK + jj * 8;
            cp_async_16(smem_b + s * Cfg::kBBytes + (uint32_t)row * 128u + (((uint32_t)jj ^ (uint32_t)(row & 7)) << 4), src, 16u);
          }
          cp_async_mbar_arrive_noinc(bar_full + 8 * s);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =============================== MMA issuer =========================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      uint32_t it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const uint32_t a = lt & 1u, aph = (lt >> 1) & 1u;
        mbar_wait(bar_tempty + 8 * a, aph ^ 1u);  // epilogue has drained this accumulator slot
        tcgen05_fence_after();
        const uint32_t acc = tmem_acc + a * Cfg::kAccCols;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % S;
          mbar_wait(bar_full + 8 * s, (it / S) & 1u);
          fence_proxy_async_smem();  // cp.async / st.shared operands -> async proxy
          tcgen05_fence_after();
          const uint64_t adesc = make_smem_desc(smem_a + s * Cfg::kABytes);
          const uint64_t bdesc = make_smem_desc(smem_b + s * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(bar_empty + 8 * s);
        }
        umma_commit(bar_tfull + 8 * a);
      }
    }
  } else {
    // =============================== epilogue warps ======================================
    const int ew = warp - kEpiWarp0;   // staging slot
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    float* stg = reinterpret_cast<float*>(smem_gen + (epi_stage - smem_base)) + ew * (32 * kStagePitch);
    const int rq = lane >> 3, cq = (lane & 7) * 4;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const uint32_t a = lt & 1u, aph = (lt >> 1) & 1u;
      const int n0 = (tile / num_m_tiles) * BN;
      TileGeom geom;
      geom.init(p, tile % num_m_tiles);
      int m_it[8], dst_it[8];
#pragma unroll
      for (int i8 = 0; i8 < 8; ++i8) {
        int im, oy, ox, mm;
        const bool ok = geom.row_pixel(p, q * 32 + i8 * 4 + rq, im, oy, ox, mm);
        m_it[i8] = ok ? mm : -1;
        dst_it[i8] = (ok && p.epi == BDE_EPI_SCATTER) ? __ldg(p.row_map + mm) : -1;
      }
      mbar_wait(bar_tfull + 8 * a, aph);
      tcgen05_fence_after();
      const uint32_t lane_taddr = tmem_acc + a * Cfg::kAccCols + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int cb = 0; cb < BN; cb += 32) {
        float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias != nullptr) bq = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + cb + cq));
        uint32_t raw[32];
        tmem_ld_32x32b_x32(lane_taddr + (uint32_t)cb, raw);
        tmem_ld_wait();
        if (cb + 32 >= BN) {
          // last TMEM read of this tile: hand the accumulator slot back to the MMA warp
          tcgen05_fence_before();
          mbar_arrive(bar_tempty + 8 * a);
        }
        __syncwarp();
#pragma unroll
        for (int jq = 0; jq < 32; jq += 4)
          *reinterpret_cast<float4*>(stg + lane * kStagePitch + jq) =
              make_float4(__uint_as_float(raw[jq]), __uint_as_float(raw[jq + 1]), __uint_as_float(raw[jq + 2]),
                          __uint_as_float(raw[jq + 3]));
        __syncwarp();
#pragma unroll
        for (int i8 = 0; i8 < 8; ++i8) {
          const float4 acc = *reinterpret_cast<const float4*>(stg + (i8 * 4 + rq) * kStagePitch + cq);
          if (m_it[i8] >= 0) epilogue_quad(p, m_it[i8], n0 + cb + cq, acc, bq, dst_it[i8]);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_acc, Cfg::kTmemCols);
  }
}

template <int BN, bool kBTma, int kLn>
int launch_persistent(const CUtensorMap& tmap, const TcParams& p, cudaStream_t s) {
  using Cfg = PCfg<BN, kLn>;
  auto kern = gemm_tc_persistent_kernel<BN, kBTma, kLn>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    BDE_REQUIRE(e == cudaSuccess, "bde_gemm(tcgen05 persistent): smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const size_t m_tiles = p.tiles_x > 0 ? (size_t)p.n_img * p.tiles_x * p.tiles_y : ceil_div(p.M, BM);
  const size_t tiles = m_tiles * (p.N / BN);
  BDE_REQUIRE(tiles < ((size_t)1 << 31), "bde_gemm(tcgen05 persistent): too many tiles");
  const unsigned grid = (unsigned)(tiles < (size_t)kNumSMs ? tiles : (size_t)kNumSMs);
  kern<<<grid, kPThreads, Cfg::kSmemBytes, s>>>(tmap, p, (int)m_tiles, (int)tiles);
  return check_launch("gemm_tc_persistent_kernel");
}

}  // namespace tc

int gemm_tcgen05_persistent(const bde_gemm_desc* d, cudaStream_t s) {
  using namespace tc;
  TcParams p;
  bool ln = false;
  int rc = fill_params(d, p, ln);
  if (rc != 0) return rc;
  if (p.M == 0) return 0;
  // tile width: the widest N tile that divides N (wide tiles amortise the A operand; persistence takes care
  // of SM load balance), except that very small problems prefer more, narrower tiles
  const size_t m_tiles = p.tiles_x > 0 ? (size_t)p.n_img * p.tiles_x * p.tiles_y : ceil_div(p.M, BM);
  int bn;
  if (ln) {
    bn = (p.N % 256 == 0) ? 256 : (p.N % 192 == 0) ? 192 : (p.N % 128 == 0) ? 128 : (p.N % 64 == 0) ? 64 : 32;
  } else {
    bn = (p.N % 256 == 0) ? 256 : (p.N % 128 == 0) ? 128 : (p.N % 64 == 0) ? 64 : 32;
    while (bn > 64 && p.N % (bn / 2) == 0 && m_tiles * (p.N / bn) < (size_t)kNumSMs) bn /= 2;
  }
  const bool tma = b_via_tma();
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (tma) {
    rc = get_weight_tmap(d->w, p.N, p.w_ld, bn, &tmap);
    if (rc != 0) return rc;
  }
#define BDE_P(BN_, LN_) return tma ? launch_persistent<BN_, true, LN_>(tmap, p, s) : launch_persistent<BN_, false, LN_>(tmap, p, s)
#define BDE_P_BN(LN_)          \
  switch (bn) {                \
    case 32: BDE_P(32, LN_);   \
    case 64: BDE_P(64, LN_);   \
    case 128: BDE_P(128, LN_); \
    case 192: BDE_P(192, LN_); \
    default: BDE_P(256, LN_);  \
  }
  if (ln) {
    const int chunks = p.c0 / 64;
    if (chunks == 1) { BDE_P_BN(1) }
    if (chunks == 2) { BDE_P_BN(2) }
    BDE_P_BN(4)
  }
  BDE_P_BN(0)
#undef BDE_P_BN
#undef BDE_P
}

}  // namespace bde
