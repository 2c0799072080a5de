// Fused transformer MLP half of a SwinTransformerBlock3D (model/BDE2VID/DTransformer.py:279-283,302-304):
//
//   x[m, :] += fc2( GELU( fc1( LayerNorm(x[m, :]) ) ) )          x: float32 [P, C], updated in place
//
// for C = 64, hidden = 256 (attention level 1 of the assumed cfg; every pixel is a token).  One CTA owns
// 128 rows.  Both GEMMs run on tcgen05 with fp32 accumulators in TMEM and the [128 x 256] hidden
// activation never leaves the SM: the GELU epilogue of GEMM 1 writes it as bf16 straight into the
// 128B-swizzled K-major shared-memory layout that GEMM 2 consumes as its A operand.  (The unfused path
// wrote the hidden tensor, 4x the size of x, to HBM and read it back.)
//
//   warps 0-7  LayerNorm producer (fp32 rows -> normalised bf16 A tile; affine folded into W1 / b1 by the
//              caller), then epilogue 1 (TMEM -> +b1 -> GELU -> bf16 -> smem) and epilogue 2
//              (TMEM -> +b2 + residual -> coalesced fp32 store through a smem transpose)
//   warp 8     one thread: TMA loads of W1 [256 x 64] and W2 [64 x 256] (SWIZZLE_128B), then both MMA chains
// Shared memory ~100 KB and 256 TMEM columns (accumulator 2 reuses accumulator 1's columns) -> 2 CTAs / SM.
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace bde {
namespace tc {

constexpr int kMlpC = 64, kMlpH = 256;
constexpr int kMlpThreads = kNumProducerThreads + 32;
constexpr int kMlpOffW1 = BM * 128;                      // A tile: 128 rows x 128 B
constexpr int kMlpOffW2 = BM * kMlpH * 2;                // after the 64 KB region shared by (A, W1) and H
constexpr int kMlpOffBar = kMlpOffW2 + kMlpC * kMlpH * 2;
constexpr int kMlpOffBias = kMlpOffBar + 128;
constexpr int kMlpSmem = kMlpOffBias + (kMlpH + kMlpC) * 4 + 1024;

struct MlpParams {
  float* x;          // [P, 64] in / out
  const float* b1;   // [256]
  const float* b2;   // [64]
  int P;
  // optional fused "x + merged" of the generator (..._V5.py:166-169) after the last block of a frame:
  // sum_io[m] += x_new[m] (fp32, in place) and sum_t[m] = bf16(sum_io[m])
  float* sum_io;
  __nv_bfloat16* sum_t;
};

__device__ __forceinline__ void mlp_fused_sum(float* sum_io, __nv_bfloat16* sum_t, size_t off, float4 xnew) {
  float4 s4 = *reinterpret_cast<const float4*>(sum_io + off);
  s4.x += xnew.x; s4.y += xnew.y; s4.z += xnew.z; s4.w += xnew.w;
  *reinterpret_cast<float4*>(sum_io + off) = s4;
  if (sum_t != nullptr) {
    const __nv_bfloat162 t0 = __floats2bfloat162_rn(s4.x, s4.y), t1 = __floats2bfloat162_rn(s4.z, s4.w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&t0);
    pk.y = *reinterpret_cast<const uint32_t*>(&t1);
    *reinterpret_cast<uint2*>(sum_t + off) = pk;
  }
}

__global__ void __launch_bounds__(kMlpThreads, 2)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_w2, const MlpParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sb - smem_u32(smem_raw));
  const uint32_t s_a = sb, s_w1 = sb + kMlpOffW1, s_h = sb, s_w2 = sb + kMlpOffW2;
  const uint32_t bar_a = sb + kMlpOffBar, bar_w1 = bar_a + 8, bar_w2 = bar_a + 16, bar_acc1 = bar_a + 24, bar_h = bar_a + 32,
                 bar_acc2 = bar_a + 40, tmem_slot = bar_a + 48;
  float* b1_s = reinterpret_cast<float*>(sgen + kMlpOffBias);
  float* b2_s = b1_s + kMlpH;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;

  for (int i = threadIdx.x; i < kMlpH + kMlpC; i += kMlpThreads) b1_s[i] = i < kMlpH ? __ldg(p.b1 + i) : __ldg(p.b2 + i - kMlpH);
  if (threadIdx.x == 0) {
    mbar_init(bar_a, kNumProducerThreads);
    mbar_init(bar_w1, 1);
    mbar_init(bar_w2, 1);
    mbar_init(bar_acc1, 1);
    mbar_init(bar_h, kNumProducerThreads);
    mbar_init(bar_acc2, 1);
    fence_barrier_init();
  }
  if (warp == kNumProducerWarps) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_acc = *reinterpret_cast<volatile uint32_t*>(sgen + kMlpOffBar + 48);

  if (warp < kNumProducerWarps) {
    // ---------------- LayerNorm producer: lane j = lane & 7 owns 8 channels of rows warp*16 + 4i + (lane >> 3) -------
    const int j = lane & 7, rsub = lane >> 3;
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
      const int row = warp * (4 * kRowsPerThread) + i * 4 + rsub;
      const int mm = m0 + row;
      float v[8];
      if (mm < p.P) {
        const float4 t0 = *reinterpret_cast<const float4*>(p.x + (size_t)mm * kMlpC + j * 8);
        const float4 t1 = *reinterpret_cast<const float4*>(p.x + (size_t)mm * kMlpC + j * 8 + 4);
        v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
      }
      float sum = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += v[e];
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      sum += __shfl_xor_sync(0xffffffffu, sum, 4);
      const float mean = sum / (float)kMlpC;
      float sq = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[e] -= mean;
        sq += v[e] * v[e];
      }
      sq += __shfl_xor_sync(0xffffffffu, sq, 1);
      sq += __shfl_xor_sync(0xffffffffu, sq, 2);
      sq += __shfl_xor_sync(0xffffffffu, sq, 4);
      const float rstd = 1.0f / sqrtf(sq / (float)kMlpC + 1e-5f);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] *= rstd;
      const uint4 pk = pack8_bf16(v);
      const uint32_t dst = s_a + (uint32_t)row * 128u + (((uint32_t)j ^ (uint32_t)(row & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
    }
    fence_proxy_async_smem();
    mbar_arrive(bar_a);

    // ---------------- epilogue 1: hidden = GELU(acc1 + b1) -> bf16, swizzled K-major tile for GEMM 2 -----------------
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;  // TMEM lane == tile row
    const uint32_t lane_taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
    mbar_wait(bar_acc1, 0);
    tcgen05_fence_after();
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      const int c0 = half * 128 + ch * 32;
      uint32_t raw[32];
      tmem_ld_32x32b_x32(lane_taddr + (uint32_t)c0, raw);
      tmem_ld_wait();
      const uint32_t kb_base = s_h + (uint32_t)(c0 >> 6) * (BM * 128) + (uint32_t)row * 128u;
      const int j0 = (c0 & 63) >> 3;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fast_gelu(__uint_as_float(raw[jj * 8 + e]) + b1_s[c0 + jj * 8 + e]);
        const uint4 pk = pack8_bf16(o);
        const uint32_t dst = kb_base + ((((uint32_t)(j0 + jj)) ^ (uint32_t)(row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
      }
    }
    tcgen05_fence_before();   // our TMEM reads are done before GEMM 2 overwrites the columns
    fence_proxy_async_smem();
    mbar_arrive(bar_h);

    // ---------------- epilogue 2: x += acc2 + b2, coalesced through a per-warp smem transpose ---------------------------
    mbar_wait(bar_acc2, 0);
    tcgen05_fence_after();
    constexpr int kPitch = 36;
    float* stg = reinterpret_cast<float*>(sgen) + warp * (32 * kPitch);  // the H region is free once GEMM 2 has completed
    const int rq = lane >> 3, cq = (lane & 7) * 4;
    {
      uint32_t raw[32];
      tmem_ld_32x32b_x32(lane_taddr + (uint32_t)(half * 32), raw);
      tmem_ld_wait();
#pragma unroll
      for (int jq = 0; jq < 32; jq += 4)
        *reinterpret_cast<float4*>(stg + lane * kPitch + jq) =
            make_float4(__uint_as_float(raw[jq]), __uint_as_float(raw[jq + 1]), __uint_as_float(raw[jq + 2]), __uint_as_float(raw[jq + 3]));
      __syncwarp();
      const float4 bb = *reinterpret_cast<const float4*>(b2_s + half * 32 + cq);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int mm = m0 + q * 32 + it * 4 + rq;
        if (mm < p.P) {
          const float4 acc = *reinterpret_cast<const float4*>(stg + (it * 4 + rq) * kPitch + cq);
          float4* dst = reinterpret_cast<float4*>(p.x + (size_t)mm * kMlpC + half * 32 + cq);
          float4 cur = *dst;
          cur.x += acc.x + bb.x; cur.y += acc.y + bb.y; cur.z += acc.z + bb.z; cur.w += acc.w + bb.w;
          *dst = cur;
          if (p.sum_io != nullptr) mlp_fused_sum(p.sum_io, p.sum_t, (size_t)mm * kMlpC + half * 32 + cq, cur);
        }
      }
    }
    tcgen05_fence_before();
  } else {
    // ---------------- TMA + MMA warp: all lanes wait, one elected lane issues (elect_one_sync, tc_common.cuh) --------------
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(bar_w1, kMlpH * 128);
      tma_load_2d(s_w1, &tmap_w1, bar_w1, 0, 0);
      mbar_arrive_expect_tx(bar_w2, kMlpC * kMlpH * 2);
#pragma unroll
      for (int kb = 0; kb < kMlpH / BK; ++kb) tma_load_2d(s_w2 + kb * (kMlpC * 128), &tmap_w2, bar_w2, kb * BK, 0);
    }
    __syncwarp();
    // GEMM 1: [128 x 64] x [64 x 256]
    mbar_wait(bar_a, 0);
    mbar_wait(bar_w1, 0);
    tcgen05_fence_after();
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc(kMlpH);
      const uint64_t adesc = make_smem_desc(s_a), bdesc = make_smem_desc(s_w1);
#pragma unroll
      for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
      umma_commit(bar_acc1);
    }
    __syncwarp();
    // GEMM 2: [128 x 256] x [256 x 64], A = the hidden tile written by epilogue 1
    mbar_wait(bar_h, 0);
    mbar_wait(bar_w2, 0);
    tcgen05_fence_after();
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc(kMlpC);
#pragma unroll
      for (int kb = 0; kb < kMlpH / BK; ++kb) {
        const uint64_t adesc = make_smem_desc(s_h + kb * (BM * 128)), bdesc = make_smem_desc(s_w2 + kb * (kMlpC * 128));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
      }
      umma_commit(bar_acc2);
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == kNumProducerWarps) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_acc, 256);
  }
}


// =====================================================================================================================
// C = 256, hidden = 1024 (attention level 3 of the assumed cfg).  Same contract, different schedule: the hidden
// activation is produced and consumed in 8 chunks of 128 columns so that it never exceeds 2 x 32 KB of shared memory:
//
//   acc1[a] (128 TMEM cols, double buffered) = LN(x) [128 x 256] . W1[chunk j]^T          (fc1, N = 128, K = 256)
//   H[a]    (bf16, swizzled K-major, double buffered) = GELU(acc1[a] + b1[chunk j])        (epilogue warps)
//   acc2    (256 TMEM cols) += H[a] [128 x 128] . W2[:, chunk j]^T                         (fc2, N = 256, K = 128)
//
// so the tensor core works on chunk j + 1 / j - 1 while the eight epilogue warps run the GELU of chunk j.
//   warps 0-7  LayerNorm producer, GELU epilogues, final residual epilogue      warp 8  TMA (weights, 32 KB ring stages)
//   warp 9     MMA issuer (owns the 512 TMEM columns)
// One CTA per 128 rows (46 CTAs for four 33 x 44 maps); replaces two GEMM launches whose 276 CTAs each re-derived the
// LayerNorm or round-tripped the [rows x 1024] hidden tensor through HBM.
constexpr int kM3C = 256, kM3H = 1024, kM3Chunk = 128, kM3NChunks = kM3H / kM3Chunk;
constexpr int kM3Threads = kNumProducerThreads + 64;
constexpr int kM3Stage = 32768, kM3Stages = 3;
constexpr int kM3OffXn = 0;                                  // 4 slabs x 16 KB
constexpr int kM3OffH = 65536;                               // 2 buffers x (2 slabs x 16 KB)
constexpr int kM3OffW = 131072;                              // ring
constexpr int kM3OffBar = kM3OffW + kM3Stages * kM3Stage;
constexpr int kM3Smem = kM3OffBar + 256 + 1024;

struct Mlp3Params {
  float* x;          // [P, 256] in / out
  const float* b1;   // [1024]
  const float* b2;   // [256]
  int P;
  float* sum_io;             // optional fused "x + merged" (see MlpParams)
  __nv_bfloat16* sum_t;
};

// CL > 1: a thread-block CLUSTER of CL CTAs shares one 128-row tile and splits the HIDDEN dimension: CTA `rank` runs the
// chunks rank, rank + CL, ... (fc1 + GELU + its K-slice of fc2), so every CTA streams only 1 / CL of the 1 MB of weights
// and does 1 / CL of the GELU work, and CL x as many SMs are busy (46 tiles -> 92 CTAs for four 33 x 44 maps; the single-CTA
// form ran on 46 of the 148 SMs).  The partial fc2 accumulators are then exchanged through distributed shared memory:
// CTA r finalises output columns [256 r / CL, 256 (r + 1) / CL) = its own TMEM partial + the peers' partials in a FIXED
// order (deterministic, no atomics), + bias + residual.
__device__ __forceinline__ uint32_t m3_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void m3_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t m3_mapa(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void m3_st_cluster_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int CL>
__global__ void __launch_bounds__(kM3Threads, 1)
mlp_fused256_kernel(const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_w2, const Mlp3Params p) {
  constexpr int NL = kM3NChunks / CL;          // hidden chunks of this CTA: rank, rank + CL, ...
  constexpr int NC = kM3C / CL;                // output columns this CTA finalises
  static_assert(CL == 1 || CL == 2 || CL == 4, "cluster sizes 1, 2, 4");
  const int rank = CL > 1 ? (int)m3_cluster_rank() : 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sb - smem_u32(smem_raw));
  const uint32_t s_xn = sb + kM3OffXn, s_h = sb + kM3OffH, s_w = sb + kM3OffW;
  const uint32_t bar0 = sb + kM3OffBar;
  const uint32_t bar_xn = bar0;                    // 8
  const uint32_t bar_wfull = bar0 + 8;             // 3 x 8
  const uint32_t bar_wempty = bar0 + 32;           // 3 x 8
  const uint32_t bar_a1full = bar0 + 56;           // 2 x 8
  const uint32_t bar_a1empty = bar0 + 72;          // 2 x 8
  const uint32_t bar_hfull = bar0 + 88;            // 2 x 8
  const uint32_t bar_hempty = bar0 + 104;          // 2 x 8
  const uint32_t bar_a2full = bar0 + 120;          // 8
  const uint32_t tmem_slot = bar0 + 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (blockIdx.x / CL) * BM;

  if (threadIdx.x == 0) {
    mbar_init(bar_xn, kNumProducerThreads);
    for (int s = 0; s < kM3Stages; ++s) {
      mbar_init(bar_wfull + 8 * s, 1);
      mbar_init(bar_wempty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_a1full + 8 * a, 1);
      mbar_init(bar_a1empty + 8 * a, kNumProducerWarps);
      mbar_init(bar_hfull + 8 * a, kNumProducerThreads);
      mbar_init(bar_hempty + 8 * a, 1);
    }
    mbar_init(bar_a2full, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sgen + kM3OffBar + 128);
  const uint32_t acc2 = tmem_base + 256;   // acc1[a] = tmem_base + 128 a

  if (warp < kNumProducerWarps) {
    // ---------------- LayerNorm producer (affine folded into W1 / b1 by the caller) ------------------------------------
    {
      const int j = lane & 7, rsub = lane >> 3;
#pragma unroll 2
      for (int i = 0; i < kRowsPerThread; ++i) {
        const int row = warp * (4 * kRowsPerThread) + i * 4 + rsub;
        const int mm = m0 + row;
        float v[4][8];
        float sum = 0.f;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          if (mm < p.P) {
            const float* src = p.x + (size_t)mm * kM3C + kb * 64 + j * 8;
            const float4 t0 = *reinterpret_cast<const float4*>(src), t1 = *reinterpret_cast<const float4*>(src + 4);
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
        const float mean = sum / (float)kM3C;
        float sq = 0.f;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            v[kb][e] -= mean;
            sq += v[kb][e] * v[kb][e];
          }
        sq += __shfl_xor_sync(0xffffffffu, sq, 1);
        sq += __shfl_xor_sync(0xffffffffu, sq, 2);
        sq += __shfl_xor_sync(0xffffffffu, sq, 4);
        const float rstd = 1.0f / sqrtf(sq / (float)kM3C + 1e-5f);
        const uint32_t dst = s_xn + (uint32_t)row * 128u + (((uint32_t)j ^ (uint32_t)(row & 7)) << 4);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = v[kb][e] * rstd;
          const uint4 pk = pack8_bf16(o);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + kb * (BM * 128)), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w)
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_xn);
    }
    // ---------------- GELU epilogues: thread = tile row, 64 hidden columns of the chunk per warp half ------------------
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
#pragma unroll 1
    for (int jl = 0; jl < NL; ++jl) {
      const int jc = rank + CL * jl;            // global hidden chunk
      const uint32_t a = jl & 1, ph = (jl >> 1) & 1;
      mbar_wait(bar_a1full + 8 * a, ph);
      tcgen05_fence_after();
      uint32_t raw0[32], raw1[32];
      tmem_ld_32x32b_x32(tmem_base + a * 128 + lane_off + (uint32_t)(half * 64), raw0);
      tmem_ld_32x32b_x32(tmem_base + a * 128 + lane_off + (uint32_t)(half * 64 + 32), raw1);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a1empty + 8 * a);   // fc1 of chunk jc + 2 may overwrite the accumulator
      mbar_wait(bar_hempty + 8 * a, ph ^ 1u);            // fc2 of chunk jc - 2 has finished reading H[a]
      const float* b1p = p.b1 + jc * kM3Chunk + half * 64;
      const uint32_t hrow = s_h + a * 32768 + (uint32_t)half * (BM * 128) + (uint32_t)row * 128u;   // slab `half` of H[a]
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float4 bA = __ldg(reinterpret_cast<const float4*>(b1p + jj * 8)), bB = __ldg(reinterpret_cast<const float4*>(b1p + jj * 8 + 4));
        const uint32_t* r = jj < 4 ? raw0 + jj * 8 : raw1 + (jj - 4) * 8;
        float o[8];
        o[0] = fast_gelu(__uint_as_float(r[0]) + bA.x); o[1] = fast_gelu(__uint_as_float(r[1]) + bA.y);
        o[2] = fast_gelu(__uint_as_float(r[2]) + bA.z); o[3] = fast_gelu(__uint_as_float(r[3]) + bA.w);
        o[4] = fast_gelu(__uint_as_float(r[4]) + bB.x); o[5] = fast_gelu(__uint_as_float(r[5]) + bB.y);
        o[6] = fast_gelu(__uint_as_float(r[6]) + bB.z); o[7] = fast_gelu(__uint_as_float(r[7]) + bB.w);
        const uint4 pk = pack8_bf16(o);
        const uint32_t dst = hrow + ((((uint32_t)jj) ^ (uint32_t)(row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_hfull + 8 * a);
    }
    mbar_wait(bar_a2full, 0);     // this CTA's (partial) fc2 accumulator is complete: all its MMAs have retired
    tcgen05_fence_after();
  } else if (warp == kTmaWarp) {
    // ---------------- weight loads, in the order the MMA warp consumes them ----------------------------------------------
    {
      uint32_t it = 0;
      auto load_w1 = [&](int chunk, int i) {   // item i of fc1(chunk): k slabs 2i, 2i + 1 of W1 rows [128 chunk, +128)
        const uint32_t s = it % kM3Stages;
        mbar_wait(bar_wempty + 8 * s, ((it / kM3Stages) & 1u) ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(bar_wfull + 8 * s, kM3Stage);
          tma_load_2d(s_w + s * kM3Stage, &tmap_w1, bar_wfull + 8 * s, (2 * i) * BK, chunk * kM3Chunk);
          tma_load_2d(s_w + s * kM3Stage + 16384, &tmap_w1, bar_wfull + 8 * s, (2 * i + 1) * BK, chunk * kM3Chunk);
        }
        __syncwarp();
        ++it;
      };
      auto load_w2 = [&](int chunk, int i) {   // item i of fc2(chunk): k slab 2 chunk + i of W2 (all 256 rows)
        const uint32_t s = it % kM3Stages;
        mbar_wait(bar_wempty + 8 * s, ((it / kM3Stages) & 1u) ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(bar_wfull + 8 * s, kM3Stage);
          tma_load_2d(s_w + s * kM3Stage, &tmap_w2, bar_wfull + 8 * s, (2 * chunk + i) * BK, 0);
        }
        __syncwarp();
        ++it;
      };
      load_w1(rank, 0); load_w1(rank, 1);
      if (NL > 1) { load_w1(rank + CL, 0); load_w1(rank + CL, 1); }
      for (int jl = 0; jl < NL; ++jl) {
        const int jc = rank + CL * jl;
        load_w2(jc, 0); load_w2(jc, 1);
        if (jl + 2 < NL) { load_w1(jc + 2 * CL, 0); load_w1(jc + 2 * CL, 1); }
      }
    }
    __syncwarp();
  } else {
    // ---------------- MMA issuer: all lanes walk the schedule, one elected lane issues -------------------------------------
    {
      constexpr uint32_t idesc1 = make_idesc(kM3Chunk), idesc2 = make_idesc(kM3C);
      uint32_t it = 0;
      auto fc1 = [&](int chunk) {   // chunk = LOCAL chunk index (the weights of the global chunk arrive through the ring)
        const uint32_t a = chunk & 1;
        mbar_wait(bar_a1empty + 8 * a, ((chunk >> 1) & 1u) ^ 1u);   // the epilogue drained this accumulator (chunk - 2)
        tcgen05_fence_after();
        for (int i = 0; i < 2; ++i, ++it) {
          const uint32_t s = it % kM3Stages;
          mbar_wait(bar_wfull + 8 * s, (it / kM3Stages) & 1u);
          tcgen05_fence_after();
          if (elect_one_sync()) {
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
              const uint64_t adesc = make_smem_desc(s_xn + (2 * i + sl) * (BM * 128)), bdesc = make_smem_desc(s_w + s * kM3Stage + sl * 16384);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                umma_bf16(tmem_base + a * 128, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc1, (i | sl | k) != 0 ? 1u : 0u);
            }
            umma_commit(bar_wempty + 8 * s);
            if (i == 1) umma_commit(bar_a1full + 8 * a);
          }
          __syncwarp();
        }
      };
      mbar_wait(bar_xn, 0);
      fc1(0);
      if (NL > 1) fc1(1);
      for (int jc = 0; jc < NL; ++jc) {
        const uint32_t a = jc & 1;
        mbar_wait(bar_hfull + 8 * a, (jc >> 1) & 1u);
        tcgen05_fence_after();
        for (int i = 0; i < 2; ++i, ++it) {
          const uint32_t s = it % kM3Stages;
          mbar_wait(bar_wfull + 8 * s, (it / kM3Stages) & 1u);
          tcgen05_fence_after();
          if (elect_one_sync()) {
            const uint64_t adesc = make_smem_desc(s_h + a * 32768 + i * (BM * 128)), bdesc = make_smem_desc(s_w + s * kM3Stage);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(acc2, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc2, (jc | i | k) != 0 ? 1u : 0u);
            umma_commit(bar_wempty + 8 * s);
            if (i == 1) {
              umma_commit(bar_hempty + 8 * a);   // H[a] may be rewritten once these MMAs have read it
              if (jc == NL - 1) umma_commit(bar_a2full);
            }
          }
          __syncwarp();
        }
        if (jc + 2 < NL) fc1(jc + 2);
      }
    }
  }

  // ---------------- exchange of the partial accumulators (CL > 1) and final epilogue -------------------------------------
  // staging slots (fp32 [128 rows][NC cols], 16-byte chunks XOR-swizzled by row) live in the XN / H regions, which are free
  // in EVERY CTA of the cluster once all of them have passed the first cluster barrier (their MMAs have retired)
  constexpr int kSlotBytes = BM * NC * 4;
  const int q = warp & 3, half = warp >> 2;
  const uint32_t lane_off = (uint32_t)(q * 32) << 16;
  if (CL > 1) {
    m3_cluster_sync();
    if (warp < kNumProducerWarps) {
      const int row = q * 32 + lane;
      for (int pr = 0; pr < CL; ++pr) {
        if (pr == rank) continue;
        const int slot = rank < pr ? rank : rank - 1;            // my slot in peer pr (its peers in rank order)
        const uint32_t dst_row = m3_mapa(sb + slot * kSlotBytes + (uint32_t)row * (NC * 4), (uint32_t)pr);
#pragma unroll 1
        for (int c0 = half * (NC / 2); c0 < (half + 1) * (NC / 2); c0 += 32) {      // my partial of the peer's columns
          uint32_t raw[32];
          tmem_ld_32x32b_x32(acc2 + lane_off + (uint32_t)(pr * NC + c0), raw);
          tmem_ld_wait();
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const int chunk = (c0 >> 2) + j4;
            m3_st_cluster_v4(dst_row + (uint32_t)(((chunk & ~7) | ((chunk ^ row) & 7)) << 4),
                             make_float4(__uint_as_float(raw[4 * j4]), __uint_as_float(raw[4 * j4 + 1]),
                                         __uint_as_float(raw[4 * j4 + 2]), __uint_as_float(raw[4 * j4 + 3])));
          }
        }
      }
    }
    m3_cluster_sync();   // every peer's partial of my columns has landed in my staging slots
  }
  if (warp < kNumProducerWarps) {
    // x += (sum of the partial accumulators) + b2 for my columns, coalesced through a per-warp smem transpose
    float* stg = reinterpret_cast<float*>(sgen + 98304) + warp * 1024;      // above the (CL - 1) staging slots (<= 96 KB)
    const int rq = lane >> 3, cq4 = lane & 7;
#pragma unroll 1
    for (int cl = half * 32; cl < NC; cl += 64) {      // local column offset inside my NC columns
      const int cb = rank * NC + cl;
      uint32_t raw[32];
      tmem_ld_32x32b_x32(acc2 + lane_off + (uint32_t)cb, raw);
      tmem_ld_wait();
      __syncwarp();
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        *reinterpret_cast<float4*>(stg + lane * 32 + ((j4 ^ (lane & 7)) << 2)) =
            make_float4(__uint_as_float(raw[4 * j4]), __uint_as_float(raw[4 * j4 + 1]), __uint_as_float(raw[4 * j4 + 2]),
                        __uint_as_float(raw[4 * j4 + 3]));
      __syncwarp();
      const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b2 + cb + cq4 * 4));
      float4 cur[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int mm = m0 + q * 32 + it * 4 + rq;
        cur[it] = mm < p.P ? *reinterpret_cast<const float4*>(p.x + (size_t)mm * kM3C + cb + cq4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + rq;
        const int trow = q * 32 + r;
        const int mm = m0 + trow;
        float4 acc = *reinterpret_cast<const float4*>(stg + r * 32 + ((cq4 ^ (r & 7)) << 2));
        if (CL > 1) {
          const int chunk = (cl >> 2) + cq4;
#pragma unroll
          for (int sl = 0; sl < CL - 1; ++sl) {
            const float4 pv = *reinterpret_cast<const float4*>(sgen + sl * kSlotBytes + trow * (NC * 4) + (((chunk & ~7) | ((chunk ^ trow) & 7)) << 4));
            acc.x += pv.x; acc.y += pv.y; acc.z += pv.z; acc.w += pv.w;
          }
        }
        if (mm < p.P) {
          const float4 xnew = make_float4(cur[it].x + acc.x + bb.x, cur[it].y + acc.y + bb.y, cur[it].z + acc.z + bb.z, cur[it].w + acc.w + bb.w);
          *reinterpret_cast<float4*>(p.x + (size_t)mm * kM3C + cb + cq4 * 4) = xnew;
          if (p.sum_io != nullptr) mlp_fused_sum(p.sum_io, p.sum_t, (size_t)mm * kM3C + cb + cq4 * 4, xnew);
        }
      }
    }
    tcgen05_fence_before();
  }

  __syncthreads();
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tc
}  // namespace bde

using namespace bde;

extern "C" int bde_mlp_fused_supported(int c, int hidden) {
  return ((c == tc::kMlpC && hidden == tc::kMlpH) || (c == tc::kM3C && hidden == tc::kM3H)) ? 1 : 0;
}

extern "C" int bde_mlp_fused_sum(float* x, size_t rows, int c, int hidden, const void* w1, const float* b1, const void* w2,
                                 const float* b2, float* sum_io, void* sum_t, void* stream);

extern "C" int bde_mlp_fused(float* x, size_t rows, int c, int hidden, const void* w1, const float* b1, const void* w2,
                             const float* b2, void* stream) {
  return bde_mlp_fused_sum(x, rows, c, hidden, w1, b1, w2, b2, nullptr, nullptr, stream);
}

extern "C" int bde_mlp_fused_sum(float* x, size_t rows, int c, int hidden, const void* w1, const float* b1, const void* w2,
                                 const float* b2, float* sum_io, void* sum_t, void* stream) {
  using namespace bde::tc;
  BDE_REQUIRE((((uintptr_t)sum_io) & 15) == 0 && (((uintptr_t)sum_t) & 7) == 0, "bde_mlp_fused: sum operands must be 16 / 8-byte aligned");
  if (rows == 0) return 0;
  BDE_REQUIRE(bde_mlp_fused_supported(c, hidden) == 1, "bde_mlp_fused: (c, hidden) must be (64, 256) or (256, 1024) (got %d, %d)", c, hidden);
  BDE_REQUIRE(x != nullptr && w1 != nullptr && w2 != nullptr && b1 != nullptr && b2 != nullptr, "bde_mlp_fused: null operand");
  BDE_REQUIRE((((uintptr_t)x) & 15) == 0 && (((uintptr_t)w1) & 127) == 0 && (((uintptr_t)w2) & 127) == 0,
              "bde_mlp_fused: operands must be 16-byte (weights 128-byte) aligned");
  BDE_REQUIRE(rows < ((size_t)1 << 31), "bde_mlp_fused: row count overflows int32");
  CUtensorMap t1, t2;
  memset(&t1, 0, sizeof(t1));
  memset(&t2, 0, sizeof(t2));
  if (c == kM3C) {
    int rc3 = get_weight_tmap(w1, kM3H, kM3C, kM3Chunk, &t1);   // boxes [64 k x 128 n]
    if (rc3 != 0) return rc3;
    rc3 = get_weight_tmap(w2, kM3C, kM3H, kM3C, &t2);           // boxes [64 k x 256 n]
    if (rc3 != 0) return rc3;
    Mlp3Params p3;
    p3.x = x; p3.b1 = b1; p3.b2 = b2; p3.P = (int)rows;
    p3.sum_io = sum_io; p3.sum_t = (__nv_bfloat16*)sum_t;
    // cluster size: split the hidden dimension over 2 (or 4) CTAs while the grid still fits one wave of SMs
    const int tiles = (int)ceil_div(rows, BM);
    const int n_sm = device_sm_count();
    // (measured on B200, tools/mlp_probe.py: 46 tiles 29.1 -> 24.8 us with 2 CTAs per tile; 4 per tile is not faster even for 12 tiles)
    int cl = tiles * 2 <= n_sm ? 2 : 1;
    if (const char* e = getenv("BDE2VID_MLP256_CLUSTER")) {
      const int f = atoi(e);
      if (f == 1 || f == 2 || f == 4) cl = f;
    }
    void (*kern)(const CUtensorMap, const CUtensorMap, const Mlp3Params) =
        cl == 4 ? mlp_fused256_kernel<4> : (cl == 2 ? mlp_fused256_kernel<2> : mlp_fused256_kernel<1>);
    cudaError_t e = cudaSuccess;
    if (first_use_on_device((const void*)kern)) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kM3Smem);
    BDE_REQUIRE(e == cudaSuccess, "bde_mlp_fused: smem attribute: %s", cudaGetErrorString(e));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)(tiles * cl));
    cfg.blockDim = dim3(kM3Threads);
    cfg.dynamicSmemBytes = kM3Smem;
    cfg.stream = (cudaStream_t)stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, t1, t2, p3);
    BDE_REQUIRE(e == cudaSuccess, "bde_mlp_fused: launch (cluster %d): %s", cl, cudaGetErrorString(e));
    return check_launch("mlp_fused256_kernel");
  }
  int rc = get_weight_tmap(w1, kMlpH, kMlpC, kMlpH, &t1);   // one box [64 k x 256 n]
  if (rc != 0) return rc;
  rc = get_weight_tmap(w2, kMlpC, kMlpH, kMlpC, &t2);       // boxes [64 k x 64 n]
  if (rc != 0) return rc;
  if (first_use_on_device((const void*)mlp_fused_kernel)) {
    cudaError_t e = cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMlpSmem);
    BDE_REQUIRE(e == cudaSuccess, "bde_mlp_fused: smem attribute: %s", cudaGetErrorString(e));
  }
  MlpParams p;
  p.x = x; p.b1 = b1; p.b2 = b2; p.P = (int)rows;
  p.sum_io = sum_io; p.sum_t = (__nv_bfloat16*)sum_t;
  mlp_fused_kernel<<<(unsigned)ceil_div(rows, BM), kMlpThreads, kMlpSmem, (cudaStream_t)stream>>>(t1, t2, p);
  return check_launch("mlp_fused_kernel");
}
