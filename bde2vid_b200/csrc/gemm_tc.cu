// tcgen05 implicit-GEMM engine (bf16 operands, fp32 accumulation in TMEM) for sm_100a.
//
//   C[m, n] = sum_k A[m, k] * W[n, k]      m = (img, oy, ox),  k = (tap, channel)
//
// One CTA computes a 128 x BN output tile.  Warp roles (320 threads):
//   warps 0-7  A producers.  8 consecutive lanes fetch the 8 16-byte chunks of one 128-byte im2col
//              row (cp.async, zero fill for padding / K tail / M tail) into the 128B-swizzled
//              K-major layout UMMA expects, so a warp-wide copy touches 4 full cache lines.
//              Everything that depends only on the output pixel (base index, per-tap validity
//              mask) is computed once; the per-K-block work is a handful of integer ops per row.
//              After the main loop the same warps run the epilogue (warp w <-> TMEM lanes
//              32*(w&3).., column half w>>2).
//   warp 8     B producer: one thread issues a TMA 2D tiled load (SWIZZLE_128B) of the
//              [BN x 64] weight tile per K block.  (kBTma=false: cp.async gather, for bring-up.)
//   warp 9     MMA issuer: one thread issues 4 x tcgen05.mma (M128 x BN x K16) per K block and
//              commits to the stage's "empty" barrier; owns the TMEM allocation.
// Pipelines: smem full/empty mbarriers per stage, one "accumulator ready" mbarrier.
// Epilogues are fused: bias + activation (+ fp32 residual), ConvLSTM gate math, window scatter.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace bde {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // bf16 elements = 128 bytes = one swizzle row
constexpr int kNumProducerWarps = 8;
constexpr int kNumProducerThreads = kNumProducerWarps * 32;
constexpr int kRowsPerThread = BM / (kNumProducerWarps * 4);  // 4
constexpr int kThreads = kNumProducerThreads + 64;
constexpr int kTmaWarp = kNumProducerWarps, kMmaWarp = kNumProducerWarps + 1;

struct TcParams {
  const __nv_bfloat16* a0;
  const __nv_bfloat16* a1;
  const __nv_bfloat16* w;  // only used when !kBTma
  const float* bias;
  int c0, c1, ctot;
  int n_img, h_in, w_in, h_out, w_out, ksize, stride, pad;
  int M, N, K, w_ld, num_kb;
  int tiles_x, tiles_y;  // > 0: M tiles are 8 x 16 output-pixel patches (L1-friendly halo reuse); 0: 128 consecutive pixels
  uint32_t ypat_all;     // bit (ky * ksize) set for every ky: multiplying by an x-bit mask replicates it per row
  int dense;             // 1x1, stride 1, pad 0: A is a plain [M, K] matrix
  int k_order;           // 0: k = (tap, channel);  1: k = (64-channel chunk, tap, channel in chunk)
  int epi, act, out_f32;
  void* out;
  void* out2;
  const float* residual;
  const float* c_prev;
  float* c_out;
  const int* row_map;
  // LayerNorm-gather A operand (kLn kernels): A[m, :] = (x - mean) / sqrt(var + 1e-5) of the fp32 row that
  // token m = (win, d, tok) maps to (affine folded into the weights by the caller); zero tokens stay 0.
  const float* ln_f[8];
  const int* ln_map;
  int ln_D, ln_ntok;
  long long* dbg;  // optional per-CTA phase timestamps (clock64), 8 slots per CTA, for bring-up profiling
};

#define BDE_DBG(slot)                                                                     \
  do {                                                                                    \
    if (p.dbg != nullptr) p.dbg[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + (slot)] = clock64(); \
  } while (0)

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// arrive on `bar` once all cp.async issued so far by this thread have landed (counted in the
// barrier's expected arrivals: .noinc)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
// (cute::UMMA::SmemDescriptor: start[0,14) LBO[16,30) SBO[32,46) version[46,48)=1 layout[61,64)=2)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;            // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: c=F32 [4,6)=1, a=BF16 [7,10)=1, b=BF16 [10,13)=1,
// a/b K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

constexpr int kTileH = 8, kTileW = 16;  // 2-D M tile (kTileH * kTileW == BM)

// M-tile geometry.  The CTA-uniform part (integer divisions) is computed once; mapping a tile row to
// its output pixel is then a handful of adds.  Three modes:
//   2-D    : the tile is an 8 x 16 patch of output pixels of one image (convolutions)
//   dense  : 1x1 / stride 1 / no padding: the "pixel" index is the row index itself (linear layers)
//   linear : 128 consecutive output pixels in raster order (small images)
struct TileGeom {
  int img, oy0, ox0;  // 2-D mode
  int m0;             // dense / linear mode
  __device__ __forceinline__ void init(const TcParams& p, int tile) {
    img = oy0 = ox0 = 0;
    m0 = tile * BM;
    if (p.tiles_x > 0) {
      const int per_img = p.tiles_x * p.tiles_y;
      img = tile / per_img;
      const int t = tile - img * per_img;
      const int ty = t / p.tiles_x;
      oy0 = ty * kTileH;
      ox0 = (t - ty * p.tiles_x) * kTileW;
    }
  }
  // returns false for rows outside the problem; m = linear output index (epilogue addressing)
  __device__ __forceinline__ bool row_pixel(const TcParams& p, int row, int& im, int& oy, int& ox, int& m) const {
    if (p.tiles_x > 0) {
      im = img;
      oy = oy0 + (row >> 4);
      ox = ox0 + (row & 15);
      m = (im * p.h_out + oy) * p.w_out + ox;
      return oy < p.h_out && ox < p.w_out;
    }
    m = m0 + row;
    if (m >= p.M) return false;
    if (p.dense) {  // no spatial structure needed
      im = 0; oy = m; ox = 0;
      return true;
    }
    const int hw = p.h_out * p.w_out;
    im = m / hw;
    const int rem = m - im * hw;
    oy = rem / p.w_out;
    ox = rem - oy * p.w_out;
    return true;
  }
};

// kDeep = false: few stages so that two CTAs share an SM (mainloop of one overlaps the epilogue of the
//                other) -- used for linear layers, whose A operand has no reuse.
// kDeep = true : one CTA per SM with a deep pipeline; the smaller shared-memory carve-out leaves
//                >= 64 KB of L1 so that the im2col re-reads of a 2-D pixel tile (each input pixel is
//                needed by up to k*k taps) are served by L1 instead of L2.
// kLn = true : the LayerNorm-gather producer fills all K blocks of the tile at once (K = C <= 256), so it
//                needs >= 4 stages.
template <int BN, bool kDeep, int kLn = 0>
struct TileCfg {
  static constexpr int kStages = kLn > 0 ? kLn : kDeep ? (BN >= 256 ? 4 : (BN >= 128 ? 5 : (BN >= 64 ? 6 : 8)))
                                       : (BN >= 256 ? 4 : (BN >= 128 ? 3 : 4));
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + BN * 4 /*bias*/;
  static constexpr int kTmemCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
};

// fast transcendental forms for the bf16 path (relative error ~1e-6, far below bf16 resolution)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }

__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  uint4 pk;
  __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 t1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]);
  __nv_bfloat162 t3 = __floats2bfloat162_rn(v[6], v[7]);
  pk.x = *reinterpret_cast<uint32_t*>(&t0);
  pk.y = *reinterpret_cast<uint32_t*>(&t1);
  pk.z = *reinterpret_cast<uint32_t*>(&t2);
  pk.w = *reinterpret_cast<uint32_t*>(&t3);
  return pk;
}

// ------------------------------------------------------------------------------------------
// Epilogue on 4 consecutive columns [nb, nb+4) of output row m (after the shared-memory transpose every
// lane owns such a quad, 8 lanes cover 32 contiguous columns of one row -> coalesced global traffic).
// For the LSTM epilogue the quad is exactly (in, remember, out, cell) of hidden channel nb / 4.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void epilogue_quad(const TcParams& p, int m, int nb, float4 acc, const float* bias_s4,
                                              int dst_row) {
  const float4 b = *reinterpret_cast<const float4*>(bias_s4);
  float v0 = acc.x + b.x, v1 = acc.y + b.y, v2 = acc.z + b.z, v3 = acc.w + b.w;
  if (p.epi == BDE_EPI_STORE) {
    if (p.act == BDE_ACT_RELU) {
      v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
    } else if (p.act == BDE_ACT_RELU6) {
      v0 = fminf(fmaxf(v0, 0.f), 6.f); v1 = fminf(fmaxf(v1, 0.f), 6.f);
      v2 = fminf(fmaxf(v2, 0.f), 6.f); v3 = fminf(fmaxf(v3, 0.f), 6.f);
    } else if (p.act != BDE_ACT_NONE) {
      v0 = apply_act(v0, p.act); v1 = apply_act(v1, p.act); v2 = apply_act(v2, p.act); v3 = apply_act(v3, p.act);
    }
    const size_t o = (size_t)m * p.N + nb;
    if (p.residual != nullptr) {
      const float4 r = *reinterpret_cast<const float4*>(p.residual + o);
      v0 += r.x; v1 += r.y; v2 += r.z; v3 += r.w;
    }
    if (p.out_f32) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o) = make_float4(v0, v1, v2, v3);
    __nv_bfloat16* dstb = p.out_f32 ? reinterpret_cast<__nv_bfloat16*>(p.out2) : reinterpret_cast<__nv_bfloat16*>(p.out);
    if (dstb != nullptr) {
      const __nv_bfloat162 t0 = __floats2bfloat162_rn(v0, v1), t1 = __floats2bfloat162_rn(v2, v3);
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&t0);
      pk.y = *reinterpret_cast<const uint32_t*>(&t1);
      *reinterpret_cast<uint2*>(dstb + o) = pk;
    }
  } else if (p.epi == BDE_EPI_LSTM) {
    // (submodules.py:320-332)  c = sig(remember) * c_prev + sig(in) * tanh(cell);  h = sig(out) * tanh(c)
    const size_t o = (size_t)m * (p.N >> 2) + (nb >> 2);
    const float cprev = p.c_prev != nullptr ? p.c_prev[o] : 0.f;
    const float c = fast_sigmoid(v1) * cprev + fast_sigmoid(v0) * fast_tanh(v3);
    const float h = fast_sigmoid(v2) * fast_tanh(c);
    p.c_out[o] = c;
    reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(h);
  } else {  // BDE_EPI_SCATTER
    if (dst_row >= 0) {
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (size_t)dst_row * p.N + nb);
      float4 cur = *dst;
      cur.x += v0; cur.y += v1; cur.z += v2; cur.w += v3;
      *dst = cur;
    }
  }
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <int BN, bool kBTma, bool kDeep, int kLn>
__global__ void __launch_bounds__(kThreads) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_b, const TcParams p) {
  using Cfg = TileCfg<BN, kDeep, kLn>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B tiles
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;                       // S x [128 x 128B]
  const uint32_t smem_b = smem_base + S * Cfg::kABytes;    // S x [BN x 128B]
  const uint32_t bar_base = smem_base + S * Cfg::kStageBytes;
  const uint32_t bar_full = bar_base;                      // S x 8B
  const uint32_t bar_empty = bar_base + 8 * S;             // S x 8B
  const uint32_t bar_acc = bar_base + 16 * S;              // 8B
  const uint32_t tmem_slot = bar_base + 16 * S + 8;        // 4B
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic pointer to smem_base
  float* bias_s = reinterpret_cast<float*>(smem_gen + S * Cfg::kStageBytes + 256);  // BN floats

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x, n0 = blockIdx.y * BN;
  const int num_kb = p.num_kb;

  if (threadIdx.x == 0) BDE_DBG(0);
  // bias slice -> smem once (the epilogue then never waits on global memory for it)
  if ((int)threadIdx.x < BN) bias_s[threadIdx.x] = p.bias != nullptr ? __ldg(p.bias + blockIdx.y * BN + threadIdx.x) : 0.0f;
  if (threadIdx.x == 0) {
    const uint32_t full_count = kNumProducerThreads + (kBTma ? 1 : 32);
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, full_count);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
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
  if (threadIdx.x == 0) BDE_DBG(1);
  TileGeom geom;
  geom.init(p, m_tile);

  if (warp < kNumProducerWarps) {
    // =============================== A producers ========================================
    // lane l owns 16-byte chunk j = l & 7 of rows warp*16 + 4*i + (l >> 3), i = 0..3
    const int j = lane & 7;
    const int rsub = lane >> 3;
    const int ntaps = p.ksize * p.ksize;
    if (kLn > 0) {
      // ---- LayerNorm-gather producer: fp32 rows -> normalised bf16 operand tile, all K blocks at once ----
      constexpr int nchunk = kLn > 0 ? kLn : 1;  // K blocks = pipeline stages (C = 64 * nchunk)
      constexpr int C = 64 * nchunk;
      // rows are independent: unroll so that several rows' loads are in flight (all 4 for C <= 128)
#pragma unroll(kLn <= 2 ? 4 : 2)
      for (int i = 0; i < kRowsPerThread; ++i) {
        const int row = warp * (4 * kRowsPerThread) + i * 4 + rsub;
        const uint32_t dst = (uint32_t)row * 128u + (((uint32_t)j ^ (uint32_t)(row & 7)) << 4);
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
        // the 8 lanes of a row group (same rsub) reduce with xor 1, 2, 4
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
          {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = v[kb][e] * rstd;
            const uint4 pk = pack8_bf16(o);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_a + kb * Cfg::kABytes + dst), "r"(pk.x), "r"(pk.y),
                         "r"(pk.z), "r"(pk.w)
                         : "memory");
          }
        }
      }
      // generic-proxy stores -> visible to the tensor core's async proxy, then one arrival per K block
      fence_proxy_async_smem();
      for (int kb = 0; kb < nchunk; ++kb) mbar_arrive(bar_full + 8 * kb);
    } else {
    int pix0[kRowsPerThread];        // linear input pixel index of tap (0,0) (may lie outside the image)
    uint32_t vmask[kRowsPerThread];  // bit t set <=> tap t of this output pixel is inside the image
    uint32_t dsto[kRowsPerThread];   // swizzled byte offset of this lane's chunk inside a stage
    {
#pragma unroll
      for (int i = 0; i < kRowsPerThread; ++i) {
        const int row = warp * (4 * kRowsPerThread) + i * 4 + rsub;
        dsto[i] = (uint32_t)row * 128u + (((uint32_t)j ^ (uint32_t)(row & 7)) << 4);
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
          // valid taps form a rectangle [ky_lo, ky_hi) x [kx_lo, kx_hi): x bits times the pattern of valid rows
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
    }
    if (threadIdx.x == 0) BDE_DBG(2);
    // running (tap, channel) position of this lane's chunk: k = kb*64 + j*8
    int tap, c;
    if (p.k_order == 1) {  // chunk-major: every lane is on the same tap; channel = chunk*64 + j*8
      tap = 0;
      c = j * 8;
    } else {
      tap = (j * 8) / p.ctot;
      c = (j * 8) - tap * p.ctot;
    }
    int ky = tap / p.ksize, kx = tap - ky * p.ksize;
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % S;
      const uint32_t ph = (uint32_t)(kb / S) & 1u;
      const bool from0 = c < p.c0;
      const __nv_bfloat16* sbase = from0 ? p.a0 + c : p.a1 + (c - p.c0);
      const int cs = from0 ? p.c0 : p.c1;
      const int tapoff = ky * p.w_in + kx;
      const uint32_t tapbit = (tap < ntaps && c < p.ctot) ? (1u << tap) : 0u;  // K tail -> zero fill
      mbar_wait(bar_empty + 8 * s, ph ^ 1u);
      const uint32_t stage_a = smem_a + s * Cfg::kABytes;
#pragma unroll
      for (int i = 0; i < kRowsPerThread; ++i) {
        const bool ok = (vmask[i] & tapbit) != 0u;
        const __nv_bfloat16* src = ok ? sbase + (size_t)(pix0[i] + tapoff) * cs : p.a0;
        cp_async_16(stage_a + dsto[i], src, ok ? 16u : 0u);
      }
      // asynchronous arrive: fires when this thread's copies for the stage have landed, so the
      // thread runs ahead by up to S stages without ever blocking on its own loads
      cp_async_mbar_arrive_noinc(bar_full + 8 * s);
      // advance to the next K block
      if (p.k_order == 1) {
        ++tap;
        if (++kx == p.ksize) {
          kx = 0;
          if (++ky == p.ksize) {  // all taps of this 64-channel chunk done -> next chunk
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

    }  // !kLn

    // =============================== epilogue ===========================================
    // thread <-> tile row (TMEM lane) 32*(warp & 3) + lane; the two warps sharing a lane quarter
    // split the columns in alternating 32-wide chunks
    const int q = warp & 3, half = warp >> 2;
    if (threadIdx.x == 0) BDE_DBG(3);
    // TMEM hands every thread one accumulator ROW (32 columns per load).  Writing rows straight to
    // global memory would make every warp-wide store touch 32 different lines, so each warp first
    // transposes its 32 x 32 chunk through shared memory (the pipeline stages are free once the
    // accumulator is complete) and then works on quads: lane -> (row 4*it + lane/8, columns 4*(lane%8)..+3),
    // i.e. 8 lanes cover 128 contiguous bytes of one output row.
    constexpr int kPitch = 36;  // floats; 16-byte aligned rows, conflict-free for both access patterns
    float* stg = reinterpret_cast<float*>(smem_gen) + warp * (32 * kPitch);
    const int rq = lane >> 3, cq = (lane & 7) * 4;
    int m_it[8];            // output row index of (4*it + rq), -1 if outside the problem
    int dst_it[8];          // SCATTER destination
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      int im, oy, ox, mm;
      const bool ok = geom.row_pixel(p, q * 32 + it * 4 + rq, im, oy, ox, mm);
      m_it[it] = ok ? mm : -1;
      dst_it[it] = (ok && p.epi == BDE_EPI_SCATTER) ? __ldg(p.row_map + mm) : -1;
    }
    mbar_wait(bar_acc, 0);
    tcgen05_fence_after();
    if (threadIdx.x == 0) BDE_DBG(5);
    const uint32_t lane_taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int cb = half * 32; cb < BN; cb += 64) {
      uint32_t raw[32];
      tmem_ld_32x32b_x32(lane_taddr + (uint32_t)cb, raw);
      tmem_ld_wait();
      __syncwarp();  // previous chunk fully read back before it is overwritten
#pragma unroll
      for (int jq = 0; jq < 32; jq += 4)
        *reinterpret_cast<float4*>(stg + lane * kPitch + jq) =
            make_float4(__uint_as_float(raw[jq]), __uint_as_float(raw[jq + 1]), __uint_as_float(raw[jq + 2]),
                        __uint_as_float(raw[jq + 3]));
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const float4 acc = *reinterpret_cast<const float4*>(stg + (it * 4 + rq) * kPitch + cq);
        if (m_it[it] >= 0) epilogue_quad(p, m_it[it], n0 + cb + cq, acc, bias_s + cb + cq, dst_it[it]);
      }
    }
    tcgen05_fence_before();
    if (threadIdx.x == 0) BDE_DBG(6);
  } else if (warp == kTmaWarp) {
    // =============================== B producer =========================================
    if (kBTma) {
      if (lane == 0) {
        for (int kb = 0; kb < num_kb; ++kb) {
          const int s = kb % S;
          const uint32_t ph = (uint32_t)(kb / S) & 1u;
          mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          mbar_arrive_expect_tx(bar_full + 8 * s, Cfg::kBBytes);
          tma_load_2d(smem_b + s * Cfg::kBBytes, &tmap_b, bar_full + 8 * s, kb * BK, n0);
        }
      }
    } else {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        for (int row = lane >> 3; row < BN; row += 4) {
          const int jj = lane & 7;
          const __nv_bfloat16* src = p.w + (size_t)(n0 + row) * p.w_ld + kb * BK + jj * 8;
          cp_async_16(smem_b + s * Cfg::kBBytes + (uint32_t)row * 128u + (((uint32_t)jj ^ (uint32_t)(row & 7)) << 4), src, 16u);
        }
        cp_async_mbar_arrive_noinc(bar_full + 8 * s);
      }
    }
  } else {
    // =============================== MMA issuer =========================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        mbar_wait(bar_full + 8 * s, ph);
        if (kb == 0) BDE_DBG(4);
        // operands were written through the generic proxy (cp.async): order them before the
        // tensor core's async-proxy reads
        fence_proxy_async_smem();
        tcgen05_fence_after();
        const uint64_t adesc = make_smem_desc(smem_a + s * Cfg::kABytes);
        const uint64_t bdesc = make_smem_desc(smem_b + s * Cfg::kBBytes);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 bf16 = 32 bytes inside the swizzle row: +2 in the (addr >> 4) field
          umma_bf16(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(bar_empty + 8 * s);  // frees the smem stage once these MMAs have read it
      }
      umma_commit(bar_acc);  // accumulator complete
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_acc, Cfg::kTmemCols);
    if (lane == 0) BDE_DBG(7);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

// weight tensor maps are immutable per (pointer, shape, tile): cache them
int get_weight_tmap(const void* w, int n, int w_ld, int bn, CUtensorMap* out) {
  static std::mutex mu;
  static std::map<std::tuple<const void*, int, int, int>, CUtensorMap> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_tuple(w, n, w_ld, bn);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return 0;
  }
  auto encode = get_encode_fn();
  BDE_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled entry point not available");
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)w_ld, (cuuint64_t)n};
  cuuint64_t gstride[1] = {(cuuint64_t)w_ld * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BDE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  *out = m;
  return 0;
}

template <int BN, bool kBTma, bool kDeep, int kLn = 0>
int launch(const CUtensorMap& tmap, const TcParams& p, cudaStream_t s) {
  using Cfg = TileCfg<BN, kDeep, kLn>;
  auto kern = gemm_tc_kernel<BN, kBTma, kDeep, kLn>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    BDE_REQUIRE(e == cudaSuccess, "bde_gemm(tcgen05): smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const size_t m_tiles = p.tiles_x > 0 ? (size_t)p.n_img * p.tiles_x * p.tiles_y : ceil_div(p.M, BM);
  dim3 grid((unsigned)m_tiles, (unsigned)(p.N / BN));
  kern<<<grid, kThreads, Cfg::kSmemBytes, s>>>(tmap, p);
  return check_launch("gemm_tc_kernel");
}

}  // namespace

// bring-up profiling: bde_tc_debug_enable(n) allocates 8 timestamps per CTA for the next launches
static long long* g_dbg = nullptr;
static size_t g_dbg_ctas = 0;

// tuning switches (read per call; used by tools/tc_phase_probe.py to compare configurations)
static bool env_flag(const char* name, bool dflt) {
  const char* e = getenv(name);
  if (e == nullptr || e[0] == 0) return dflt;
  return e[0] == '1';
}

// BDE2VID_TC_B_CPASYNC=1 loads the weight tile with cp.async instead of TMA (bring-up / bisecting)
static bool b_via_tma() {
  const char* e = getenv("BDE2VID_TC_B_CPASYNC");
  return !(e != nullptr && e[0] == '1');
}

int gemm_tcgen05(const bde_gemm_desc* d, cudaStream_t s) {
  BDE_REQUIRE(d->dtype == BDE_BF16, "bde_gemm(tcgen05): operands must be bf16");
  TcParams p;
  p.a0 = (const __nv_bfloat16*)d->a0;
  p.a1 = (const __nv_bfloat16*)d->a1;
  p.w = (const __nv_bfloat16*)d->w;
  p.bias = d->bias;
  p.c0 = d->c0; p.c1 = d->c1; p.ctot = d->c0 + d->c1;
  p.n_img = d->n_img;
  p.h_in = d->h_in; p.w_in = d->w_in; p.h_out = d->h_out; p.w_out = d->w_out;
  p.ksize = d->ksize; p.stride = d->stride; p.pad = d->pad;
  p.M = d->n_img * d->h_out * d->w_out;
  p.N = d->n;
  p.K = d->ksize * d->ksize * p.ctot;
  p.w_ld = d->w_ld > 0 ? d->w_ld : p.K;
  p.num_kb = (p.K + BK - 1) / BK;
  p.epi = d->epi; p.act = d->act; p.out_f32 = d->out_f32;
  p.out = d->out; p.out2 = d->out2; p.residual = d->residual; p.c_prev = d->c_prev; p.c_out = d->c_out;
  p.row_map = d->row_map;
  p.dbg = nullptr;
  const bool ln = d->ln_mode != 0;
  for (int i = 0; i < 8; ++i) p.ln_f[i] = ln && i < d->ln_D ? d->ln_frames[i] : nullptr;
  p.ln_map = d->ln_tok_map;
  p.ln_D = ln ? d->ln_D : 1;
  p.ln_ntok = ln ? (d->ln_n_tok > 0 ? d->ln_n_tok : 1) : 1;
  if (ln) {
    BDE_REQUIRE(d->ksize == 1 && d->stride == 1 && d->pad == 0 && d->c1 == 0, "bde_gemm(tcgen05): ln_mode needs a 1x1 / dense GEMM");
    BDE_REQUIRE(d->c0 == 64 || d->c0 == 128 || d->c0 == 256, "bde_gemm(tcgen05): ln_mode supports C in {64, 128, 256}");
    BDE_REQUIRE(d->ln_D >= 1 && d->ln_D <= 8, "bde_gemm(tcgen05): ln_D must be in [1, 8]");
  }
  BDE_REQUIRE(p.c0 % 8 == 0 && p.c1 % 8 == 0, "bde_gemm(tcgen05): channel counts must be multiples of 8");
  BDE_REQUIRE(p.ksize * p.ksize <= 32, "bde_gemm(tcgen05): kernel size up to 5x5");
  BDE_REQUIRE(p.w_ld % BK == 0 && p.w_ld >= p.num_kb * BK, "bde_gemm(tcgen05): w_ld must be a zero-padded multiple of 64");
  BDE_REQUIRE(p.N % 32 == 0, "bde_gemm(tcgen05): N must be a multiple of 32");
  BDE_REQUIRE((size_t)d->n_img * d->h_in * d->w_in < ((size_t)1 << 31), "bde_gemm(tcgen05): input pixel count overflows int32");
  BDE_REQUIRE(ln || d->a0 != nullptr, "bde_gemm(tcgen05): null A operand");
  BDE_REQUIRE((((uintptr_t)d->a0) & 15) == 0 && (((uintptr_t)d->a1) & 15) == 0 && (((uintptr_t)d->w) & 127) == 0,
              "bde_gemm(tcgen05): operands must be 16-byte (weights 128-byte) aligned");
  if (p.M == 0) return 0;
  p.k_order = d->k_order;
  p.dense = (p.ksize == 1 && p.stride == 1 && p.pad == 0) ? 1 : 0;
  p.ypat_all = 0u;
  for (int ky = 0; ky < p.ksize; ++ky) p.ypat_all |= 1u << (ky * p.ksize);
  BDE_REQUIRE(p.k_order == 0 || (p.k_order == 1 && p.c0 % BK == 0 && p.c1 % BK == 0),
              "bde_gemm(tcgen05): chunk-major K order needs channel counts that are multiples of 64");
  // convolutions with a real footprint use 2-D pixel tiles + the deep one-CTA-per-SM configuration
  const bool conv = p.ksize > 1 && p.h_out >= kTileH && p.w_out >= kTileW;
  const bool tile2d = conv && env_flag("BDE2VID_TC_TILE2D", true);
  const bool deep = conv && env_flag("BDE2VID_TC_DEEP", false);
  p.tiles_x = tile2d ? (int)ceil_div(p.w_out, kTileW) : 0;
  p.tiles_y = tile2d ? (int)ceil_div(p.h_out, kTileH) : 0;
  // tile width: widest tile that still gives every SM work
  int bn = 32;
  if (p.N % 128 == 0) bn = 128;
  else if (p.N % 64 == 0) bn = 64;
  const size_t m_tiles = ceil_div(p.M, BM);
  if (p.N % 256 == 0 && m_tiles * (p.N / 256) >= 2 * (size_t)kNumSMs) bn = 256;
  if (ln) {
    // the LayerNorm is recomputed by every N tile of a row block: prefer the widest tile
    bn = (p.N % 256 == 0) ? 256 : (p.N % 192 == 0) ? 192 : (p.N % 128 == 0) ? 128 : (p.N % 64 == 0) ? 64 : 32;
  }
  if (g_dbg != nullptr) {
    const size_t mt = p.tiles_x > 0 ? (size_t)p.n_img * p.tiles_x * p.tiles_y : ceil_div(p.M, BM);
    if (mt * (p.N / bn) <= g_dbg_ctas) p.dbg = g_dbg;
  }
  const bool tma = b_via_tma();
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (tma) {
    int rc = get_weight_tmap(d->w, p.N, p.w_ld, bn, &tmap);
    if (rc != 0) return rc;
  }
  if (ln) {
    const int chunks = p.c0 / 64;
#define BDE_TC_LN(BN_, CH_) \
  return tma ? launch<BN_, true, false, CH_>(tmap, p, s) : launch<BN_, false, false, CH_>(tmap, p, s)
#define BDE_TC_LN_BN(CH_)                 \
  switch (bn) {                           \
    case 32: BDE_TC_LN(32, CH_);          \
    case 64: BDE_TC_LN(64, CH_);          \
    case 128: BDE_TC_LN(128, CH_);        \
    case 192: BDE_TC_LN(192, CH_);        \
    default: BDE_TC_LN(256, CH_);         \
  }
    if (chunks == 1) { BDE_TC_LN_BN(1) }
    if (chunks == 2) { BDE_TC_LN_BN(2) }
    BDE_TC_LN_BN(4)
#undef BDE_TC_LN_BN
#undef BDE_TC_LN
  }
#define BDE_TC_LAUNCH(BN_)                                                                                   \
  case BN_:                                                                                                  \
    if (deep) return tma ? launch<BN_, true, true>(tmap, p, s) : launch<BN_, false, true>(tmap, p, s);       \
    return tma ? launch<BN_, true, false>(tmap, p, s) : launch<BN_, false, false>(tmap, p, s);
  switch (bn) {
    BDE_TC_LAUNCH(32)
    BDE_TC_LAUNCH(64)
    BDE_TC_LAUNCH(128)
    BDE_TC_LAUNCH(256)
  }
#undef BDE_TC_LAUNCH
  BDE_REQUIRE(false, "bde_gemm(tcgen05): no tile config");
}

}  // namespace bde

extern "C" int bde_tc_debug_enable(size_t max_ctas) {
  if (bde::g_dbg != nullptr) cudaFree(bde::g_dbg);
  bde::g_dbg = nullptr;
  bde::g_dbg_ctas = 0;
  if (max_ctas == 0) return 0;
  if (cudaMalloc(&bde::g_dbg, max_ctas * 8 * sizeof(long long)) != cudaSuccess) return -1;
  cudaMemset(bde::g_dbg, 0, max_ctas * 8 * sizeof(long long));
  bde::g_dbg_ctas = max_ctas;
  return 0;
}

extern "C" int bde_tc_debug_read(long long* host_out, size_t n_ctas) {
  if (bde::g_dbg == nullptr || n_ctas > bde::g_dbg_ctas) return -1;
  return cudaMemcpy(host_out, bde::g_dbg, n_ctas * 8 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
