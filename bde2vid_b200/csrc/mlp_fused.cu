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
  long long* dbg;    // optional per-CTA phase timestamps (8 per CTA), bring-up profiling
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

// LayerNorm statistics of R rows held as v[R][NKB][8] (8 lanes per row; lane j owns channels 64 kb + 8 j .. + 7).  All R rows
// advance together and the sums are trees, so every dependent chain (adds, shuffles, rsqrt) has several independent copies
// in flight: with one row after the other and 32-long serial add chains the phase ran at ~19 cycles per instruction on the
// two warps a scheduler holds (2.3 K cycles per row, round-2 phase counters).  On return v holds x - mean.
template <int R, int NKB>
__device__ __forceinline__ void ln_rows(float (&v)[R][NKB][8], float (&rstd)[R]) {
  constexpr float kInvC = 1.0f / (float)(NKB * 64);
  float s[R];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb) {
      a0 += (v[i][kb][0] + v[i][kb][1]) + (v[i][kb][2] + v[i][kb][3]);
      a1 += (v[i][kb][4] + v[i][kb][5]) + (v[i][kb][6] + v[i][kb][7]);
    }
    s[i] = a0 + a1;
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1)
#pragma unroll
    for (int i = 0; i < R; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const float mean = s[i] * kInvC;
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        const float d0 = v[i][kb][e] - mean, d1 = v[i][kb][e + 1] - mean;
        v[i][kb][e] = d0; v[i][kb][e + 1] = d1;
        q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1);
      }
    s[i] = q0 + q1;
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1)
#pragma unroll
    for (int i = 0; i < R; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
#pragma unroll
  for (int i = 0; i < R; ++i) rstd[i] = rsqrtf(fmaf(s[i], kInvC, 1e-5f));
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
  const int n_tiles = (p.P + BM - 1) / BM;
  // bring-up timestamps (first tile of the CTA): the whole of warp 0 takes them (a warp-uniform branch) and re-converges at
  // once -- a single diverged lane made every later __shfl_sync of its warp take the divergent slow path and inflated the
  // LayerNorm phase by 6 K cycles
  const bool dbg = p.dbg != nullptr && threadIdx.x < 32;
  const long long t_begin = dbg ? clock64() : 0;
  auto mark = [&](int slot) {
    if (dbg) {
      const long long c = clock64() - t_begin;
      if (threadIdx.x == 0) p.dbg[(size_t)blockIdx.x * 8 + slot] = c;
      __syncwarp();
    }
  };

  pdl_trigger();   // the next kernel of the chain may start its prologue
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
  mark(0);

  // PERSISTENT: the CTA walks the row tiles blockIdx.x, + gridDim.x, ... (two CTAs per SM interleave their phases).  Set-up,
  // TMEM allocation and the W2 load happen once; W1 shares its shared-memory bytes with the hidden tile, so it is re-loaded
  // per tile, under the next tile's LayerNorm.  Every barrier completes one phase per tile: parity = tile counter & 1.
  if (warp < kNumProducerWarps) pdl_wait();   // x was written by the previous kernel of the chain (warp 8 only touches the static weights)
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int m0 = tile * BM;
    const uint32_t ph = it & 1u;
    const bool first = it == 0;
    if (warp < kNumProducerWarps) {
      // ---------------- LayerNorm producer: lane j = lane & 7 owns 8 channels of rows warp*16 + 4i + (lane >> 3) -------
      const int j = lane & 7, rsub = lane >> 3;
      {
        float v[kRowsPerThread][1][8];
#pragma unroll
        for (int i = 0; i < kRowsPerThread; ++i) {   // every load of the thread is in flight before the first reduction
          const int mm = m0 + warp * (4 * kRowsPerThread) + i * 4 + rsub;
          if (mm < p.P) {
            const float4 t0 = *reinterpret_cast<const float4*>(p.x + (size_t)mm * kMlpC + j * 8);
            const float4 t1 = *reinterpret_cast<const float4*>(p.x + (size_t)mm * kMlpC + j * 8 + 4);
            v[i][0][0] = t0.x; v[i][0][1] = t0.y; v[i][0][2] = t0.z; v[i][0][3] = t0.w;
            v[i][0][4] = t1.x; v[i][0][5] = t1.y; v[i][0][6] = t1.z; v[i][0][7] = t1.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[i][0][e] = 0.f;
          }
        }
        float rstd[kRowsPerThread];
        ln_rows<kRowsPerThread, 1>(v, rstd);
#pragma unroll
        for (int i = 0; i < kRowsPerThread; ++i) {
          const int row = warp * (4 * kRowsPerThread) + i * 4 + rsub;
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = v[i][0][e] * rstd[i];
          const uint4 pk = pack8_bf16(o);
          const uint32_t dst = s_a + (uint32_t)row * 128u + (((uint32_t)j ^ (uint32_t)(row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_a);
      if (first) mark(1);

      // ---------------- epilogue 1: hidden = GELU(acc1 + b1) -> bf16, swizzled K-major tile for GEMM 2 -----------------
      const int q = warp & 3, half = warp >> 2;
      const int row = q * 32 + lane;  // TMEM lane == tile row
      const uint32_t lane_taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
      mbar_wait(bar_acc1, ph);
      if (first) mark(2);
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
      if (first) mark(3);

      // ---------------- epilogue 2: x += acc2 + b2, coalesced through a per-warp smem transpose ---------------------------
      // The residual rows are fetched BEFORE waiting for GEMM 2 (they do not depend on it): inside the store loop every
      // load waited behind the previous store to the same array, eight global round trips in a row.
      const int rq = lane >> 3, cq = (lane & 7) * 4;
      float4 cur8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int mm = m0 + q * 32 + e * 4 + rq;
        cur8[e] = mm < p.P ? __ldcg(reinterpret_cast<const float4*>(p.x + (size_t)mm * kMlpC + half * 32 + cq)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      mbar_wait(bar_acc2, ph);
      if (first) mark(4);
      tcgen05_fence_after();
      constexpr int kPitch = 36;
      float* stg = reinterpret_cast<float*>(sgen) + warp * (32 * kPitch);  // the H region is free once GEMM 2 has completed
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
        for (int e = 0; e < 8; ++e) {
          const int mm = m0 + q * 32 + e * 4 + rq;
          if (mm < p.P) {
            const float4 acc = *reinterpret_cast<const float4*>(stg + (e * 4 + rq) * kPitch + cq);
            float4* dst = reinterpret_cast<float4*>(p.x + (size_t)mm * kMlpC + half * 32 + cq);
            float4 cur = cur8[e];
            cur.x += acc.x + bb.x; cur.y += acc.y + bb.y; cur.z += acc.z + bb.z; cur.w += acc.w + bb.w;
            *dst = cur;
            if (p.sum_io != nullptr) mlp_fused_sum(p.sum_io, p.sum_t, (size_t)mm * kMlpC + half * 32 + cq, cur);
          }
        }
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();   // the transpose tiles were read through the generic proxy; the next W1 load writes them through the async proxy
      if (first) mark(5);
    } else {
      // ---------------- TMA + MMA warp: all lanes wait, one elected lane issues (elect_one_sync, tc_common.cuh) --------------
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(bar_w1, kMlpH * 128);
        tma_load_2d(s_w1, &tmap_w1, bar_w1, 0, 0);
        if (first) {
          mbar_arrive_expect_tx(bar_w2, kMlpC * kMlpH * 2);
#pragma unroll
          for (int kb = 0; kb < kMlpH / BK; ++kb) tma_load_2d(s_w2 + kb * (kMlpC * 128), &tmap_w2, bar_w2, kb * BK, 0);
        }
      }
      __syncwarp();
      // GEMM 1: [128 x 64] x [64 x 256]
      mbar_wait(bar_a, ph);
      mbar_wait(bar_w1, ph);
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
      mbar_wait(bar_h, ph);
      mbar_wait(bar_w2, 0);     // completes once; later waits on the same parity return at once
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
    // end of tile: the transpose tiles of epilogue 2 (they overlay the A / W1 bytes) and the TMEM columns are free again
    // for the next tile's LayerNorm stores, W1 load and GEMM 1
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
  }

  if (warp == kNumProducerWarps) tmem_dealloc(tmem_acc, 256);
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
  long long* dbg;            // optional per-CTA phase timestamps (8 per CTA), bring-up profiling
};

// CL > 1: a thread-block CLUSTER of CL CTAs shares one 128-row tile and splits the HIDDEN dimension: CTA `rank` runs the
// chunks rank, rank + CL, ... (fc1 + GELU + its K-slice of fc2), so every CTA streams only 1 / CL of the 1 MB of weights
// and does 1 / CL of the GELU work, and CL x as many SMs are busy (46 tiles -> 92 CTAs for four 33 x 44 maps; the single-CTA
// form ran on 46 of the 148 SMs).  The partial fc2 accumulators are summed into x in CL phases separated by cluster
// barriers (see the final epilogue): a fixed order, deterministic, no atomics.
__device__ __forceinline__ uint32_t m3_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void m3_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
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
  // bring-up timestamps: the whole of warp 0 takes them (a warp-uniform branch) and re-converges at once -- a single diverged
  // lane made every later __shfl_sync of its warp take the divergent slow path and inflated the LayerNorm phase by 6 K cycles
  const bool dbg = p.dbg != nullptr && threadIdx.x < 32;
  const long long t_begin = dbg ? clock64() : 0;
  auto mark = [&](int slot) {
    if (dbg) {
      const long long c = clock64() - t_begin;
      if (threadIdx.x == 0) p.dbg[(size_t)blockIdx.x * 8 + slot] = c;
      __syncwarp();
    }
  };

  pdl_trigger();   // the next kernel of the chain may start its prologue
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

  mark(0);
  if (warp < kNumProducerWarps) {
    pdl_wait();   // x was written by the previous kernel of the chain (the TMA / MMA warps only touch the static weights)
    // ---------------- LayerNorm producer (affine folded into W1 / b1 by the caller) ------------------------------------
    // All 32 loads of the thread (4 row passes x 1 KB rows) are issued before the first reduction, then ln_rows() advances
    // the four rows together.
    {
      const int j = lane & 7, rsub = lane >> 3;
      float v[kRowsPerThread][4][8];
#pragma unroll
      for (int i = 0; i < kRowsPerThread; ++i) {
        const int mm = m0 + warp * (4 * kRowsPerThread) + i * 4 + rsub;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          if (mm < p.P) {
            const float* src = p.x + (size_t)mm * kM3C + kb * 64 + j * 8;
            const float4 t0 = *reinterpret_cast<const float4*>(src), t1 = *reinterpret_cast<const float4*>(src + 4);
            v[i][kb][0] = t0.x; v[i][kb][1] = t0.y; v[i][kb][2] = t0.z; v[i][kb][3] = t0.w;
            v[i][kb][4] = t1.x; v[i][kb][5] = t1.y; v[i][kb][6] = t1.z; v[i][kb][7] = t1.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[i][kb][e] = 0.f;
          }
        }
      }
      float rstd[kRowsPerThread];
      ln_rows<kRowsPerThread, 4>(v, rstd);
#pragma unroll
      for (int i = 0; i < kRowsPerThread; ++i) {
        const int row = warp * (4 * kRowsPerThread) + i * 4 + rsub;
        const uint32_t dst = s_xn + (uint32_t)row * 128u + (((uint32_t)j ^ (uint32_t)(row & 7)) << 4);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = v[i][kb][e] * rstd[i];
          const uint4 pk = pack8_bf16(o);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + kb * (BM * 128)), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w)
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_xn);
    }
    mark(1);
    // ---------------- GELU epilogues: thread = tile row, 64 hidden columns of the chunk per warp half ------------------
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
#pragma unroll 1
    for (int jl = 0; jl < NL; ++jl) {
      const int jc = rank + CL * jl;            // global hidden chunk
      const uint32_t a = jl & 1, ph = (jl >> 1) & 1;
      // this chunk's 64 bias values travel while the thread waits for the accumulator
      const float* b1p = p.b1 + jc * kM3Chunk + half * 64;
      float4 bia[16];
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) bia[jj] = __ldg(reinterpret_cast<const float4*>(b1p + jj * 4));
      mbar_wait(bar_a1full + 8 * a, ph);
      if (jl == 0) mark(2);
      tcgen05_fence_after();
      uint32_t raw0[32], raw1[32];
      tmem_ld_32x32b_x32(tmem_base + a * 128 + lane_off + (uint32_t)(half * 64), raw0);
      tmem_ld_32x32b_x32(tmem_base + a * 128 + lane_off + (uint32_t)(half * 64 + 32), raw1);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a1empty + 8 * a);   // fc1 of chunk jc + 2 may overwrite the accumulator
      mbar_wait(bar_hempty + 8 * a, ph ^ 1u);            // fc2 of chunk jc - 2 has finished reading H[a]
      const uint32_t hrow = s_h + a * 32768 + (uint32_t)half * (BM * 128) + (uint32_t)row * 128u;   // slab `half` of H[a]
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float4 bA = bia[2 * jj], bB = bia[2 * jj + 1];
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
    mark(3);
    mbar_wait(bar_a2full, 0);     // this CTA's (partial) fc2 accumulator is complete: all its MMAs have retired
    mark(4);
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

  // ---------------- final epilogue: x += (sum over the cluster of the partial fc2 accumulators) + b2 -----------------------
  // CL > 1: the partial sums meet IN x (global memory, i.e. L2) instead of in distributed shared memory: the DSMEM exchange
  // of 64 KB per CTA ran at the ~17 B / clk of the SM-to-SM network (7.2 K cycles, round-2 phase counters).
  // Phase k < CL - 1: CTA r adds its partial of the columns OWNED by CTA (r + 1 + k) % CL into x, then a cluster barrier
  // (release / acquire: the peers' global writes are visible; the loads bypass L1); last phase: its own columns + b2.
  // Every element is updated by exactly one CTA per phase, in a fixed order -> deterministic, no atomics.
  const int q = warp & 3, half = warp >> 2;
  const uint32_t lane_off = (uint32_t)(q * 32) << 16;
  float* stg = reinterpret_cast<float*>(sgen) + warp * 1024;   // per-warp transpose tile in the XN region (this CTA's MMAs have retired)
  const int rq = lane >> 3, cq4 = lane & 7;
  if (CL > 1) m3_cluster_sync();   // nobody writes x before every CTA of the cluster has read its LayerNorm rows
  mark(5);
#pragma unroll 1
  for (int ph = 0; ph < CL; ++ph) {
    const int owner = (rank + 1 + ph) % CL;   // ph == CL - 1: my own columns
    const bool last = ph == CL - 1;
    if (warp < kNumProducerWarps) {
      auto load_rows = [&](int cb, float4 (&dst)[8]) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int mm = m0 + q * 32 + it * 4 + rq;
          dst[it] = mm < p.P ? __ldcg(reinterpret_cast<const float4*>(p.x + (size_t)mm * kM3C + cb + cq4 * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      float4 cur[8];
      load_rows(owner * NC + half * 32, cur);
#pragma unroll 1
      for (int cl = half * 32; cl < NC; cl += 64) {
        const int cb = owner * NC + cl;
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
        float4 nxt[8];
        if (cl + 64 < NC) load_rows(cb + 64, nxt);   // the next pass's rows travel while this pass is added and stored
        const float4 bb = last ? __ldg(reinterpret_cast<const float4*>(p.b2 + cb + cq4 * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + rq;
          const int mm = m0 + q * 32 + r;
          const float4 acc = *reinterpret_cast<const float4*>(stg + r * 32 + ((cq4 ^ (r & 7)) << 2));
          if (mm < p.P) {
            const float4 xnew = make_float4(cur[it].x + acc.x + bb.x, cur[it].y + acc.y + bb.y, cur[it].z + acc.z + bb.z, cur[it].w + acc.w + bb.w);
            *reinterpret_cast<float4*>(p.x + (size_t)mm * kM3C + cb + cq4 * 4) = xnew;
            if (last && p.sum_io != nullptr) mlp_fused_sum(p.sum_io, p.sum_t, (size_t)mm * kM3C + cb + cq4 * 4, xnew);
          }
        }
        if (cl + 64 < NC) {
#pragma unroll
          for (int it = 0; it < 8; ++it) cur[it] = nxt[it];
        }
      }
    }
    if (CL > 1 && !last) {
      m3_cluster_sync();
      mark(6);
    }
  }
  if (warp < kNumProducerWarps) tcgen05_fence_before();
  mark(7);

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
    // (measured on B200, tools/mlp_probe.py, us per launch for 1 / 2 / 4 CTAs per tile: 46 tiles 25.2 / 21.0 / 33.3 (184 CTAs no longer
    // fit one wave), 12 tiles 25.1 / 20.9 / 17.9)
    int cl = tiles * 4 <= n_sm ? 4 : (tiles * 2 <= n_sm ? 2 : 1);
    if (const char* e = getenv("BDE2VID_MLP256_CLUSTER")) {
      const int f = atoi(e);
      if (f == 1 || f == 2 || f == 4) cl = f;
    }
    void (*kern)(const CUtensorMap, const CUtensorMap, const Mlp3Params) =
        cl == 4 ? mlp_fused256_kernel<4> : (cl == 2 ? mlp_fused256_kernel<2> : mlp_fused256_kernel<1>);
    p3.dbg = (g_dbg != nullptr && (size_t)tiles * cl <= g_dbg_ctas) ? g_dbg : nullptr;
    cudaError_t e = cudaSuccess;
    if (first_use_on_device((const void*)kern)) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kM3Smem);
    BDE_REQUIRE(e == cudaSuccess, "bde_mlp_fused: smem attribute: %s", cudaGetErrorString(e));
    e = launch_pdl(kern, (unsigned)(tiles * cl), (unsigned)kM3Threads, (size_t)kM3Smem, (cudaStream_t)stream, cl, t1, t2, p3);
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
  p.dbg = (g_dbg != nullptr && ceil_div(rows, BM) <= g_dbg_ctas) ? g_dbg : nullptr;   // (at most that many CTAs)
  // persistent: at most two CTAs per SM (the resident limit of this kernel), each walking its share of the row tiles
  const unsigned n_tiles = (unsigned)ceil_div(rows, BM), max_ctas = 2u * (unsigned)device_sm_count();
  const cudaError_t le = launch_pdl(mlp_fused_kernel, n_tiles < max_ctas ? n_tiles : max_ctas, (unsigned)kMlpThreads, (size_t)kMlpSmem, (cudaStream_t)stream, 1, t1, t2, p);
  BDE_REQUIRE(le == cudaSuccess, "bde_mlp_fused: launch: %s", cudaGetErrorString(le));
  return check_launch("mlp_fused_kernel");
}
