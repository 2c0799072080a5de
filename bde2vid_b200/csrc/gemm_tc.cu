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
  int h_in, w_in, h_out, w_out, ksize, stride, pad;
  int M, N, K, w_ld, num_kb;
  int epi, act, out_f32;
  void* out;
  void* out2;
  const float* residual;
  const float* c_prev;
  float* c_out;
  const int* row_map;
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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

template <int BN>
struct TileCfg {
  // 3 stages of 32 KB keep two CTAs resident per SM (mainloop of one overlaps the epilogue of the other)
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 128 ? 3 : 4);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
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
// epilogue for one row (m) and 32 consecutive columns [nb, nb+32)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void epilogue_row32(const TcParams& p, int m, int nb, const uint32_t (&raw)[32]) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + nb + j));
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (p.epi == BDE_EPI_STORE) {
    if (p.act == BDE_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
    } else if (p.act == BDE_ACT_RELU6) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fminf(fmaxf(v[j], 0.0f), 6.0f);
    } else if (p.act != BDE_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
    }
    const size_t o = (size_t)m * p.N + nb;
    if (p.residual != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 r = *reinterpret_cast<const float4*>(p.residual + o + j);
        v[j] += r.x; v[j + 1] += r.y; v[j + 2] += r.z; v[j + 3] += r.w;
      }
    }
    if (p.out_f32) {
      float* dst = reinterpret_cast<float*>(p.out) + o;
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    __nv_bfloat16* dstb = p.out_f32 ? reinterpret_cast<__nv_bfloat16*>(p.out2) : reinterpret_cast<__nv_bfloat16*>(p.out);
    if (dstb != nullptr) {
      dstb += o;
#pragma unroll
      for (int j = 0; j < 32; j += 8) *reinterpret_cast<uint4*>(dstb + j) = pack8_bf16(v + j);
    }
  } else if (p.epi == BDE_EPI_LSTM) {
    // columns = 8 hidden channels x (in, remember, out, cell)   (submodules.py:320-332)
    const int hid = p.N >> 2, ch = nb >> 2;
    const size_t o = (size_t)m * hid + ch;
    float cp[8];
    if (p.c_prev != nullptr) {
      float4 a = *reinterpret_cast<const float4*>(p.c_prev + o), b = *reinterpret_cast<const float4*>(p.c_prev + o + 4);
      cp[0] = a.x; cp[1] = a.y; cp[2] = a.z; cp[3] = a.w; cp[4] = b.x; cp[5] = b.y; cp[6] = b.z; cp[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) cp[j] = 0.f;
    }
    float h[8], c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c[j] = fast_sigmoid(v[4 * j + 1]) * cp[j] + fast_sigmoid(v[4 * j]) * fast_tanh(v[4 * j + 3]);
      h[j] = fast_sigmoid(v[4 * j + 2]) * fast_tanh(c[j]);
    }
    *reinterpret_cast<float4*>(p.c_out + o) = make_float4(c[0], c[1], c[2], c[3]);
    *reinterpret_cast<float4*>(p.c_out + o + 4) = make_float4(c[4], c[5], c[6], c[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o) = pack8_bf16(h);
  } else {  // BDE_EPI_SCATTER
    const int dst_row = p.row_map[m];
    if (dst_row >= 0) {
      float* dst = reinterpret_cast<float*>(p.out) + (size_t)dst_row * p.N + nb;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 cur = *reinterpret_cast<float4*>(dst + j);
        cur.x += v[j]; cur.y += v[j + 1]; cur.z += v[j + 2]; cur.w += v[j + 3];
        *reinterpret_cast<float4*>(dst + j) = cur;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <int BN, bool kBTma>
__global__ void __launch_bounds__(kThreads) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_b, const TcParams p) {
  using Cfg = TileCfg<BN>;
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int num_kb = p.num_kb;

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

  if (warp < kNumProducerWarps) {
    // =============================== A producers ========================================
    // lane l owns 16-byte chunk j = l & 7 of rows warp*16 + 4*i + (l >> 3), i = 0..3
    const int j = lane & 7;
    const int rsub = lane >> 3;
    const int ntaps = p.ksize * p.ksize;
    int pix0[kRowsPerThread];        // linear input pixel index of tap (0,0) (may lie outside the image)
    uint32_t vmask[kRowsPerThread];  // bit t set <=> tap t of this output pixel is inside the image
    uint32_t dsto[kRowsPerThread];   // swizzled byte offset of this lane's chunk inside a stage
    {
      const int hw = p.h_out * p.w_out;
#pragma unroll
      for (int i = 0; i < kRowsPerThread; ++i) {
        const int row = warp * (4 * kRowsPerThread) + i * 4 + rsub;
        dsto[i] = (uint32_t)row * 128u + (((uint32_t)j ^ (uint32_t)(row & 7)) << 4);
        const int mm = m0 + row;
        pix0[i] = 0;
        vmask[i] = 0u;
        if (mm < p.M) {
          const int img = mm / hw;
          const int rem = mm - img * hw;
          const int oy = rem / p.w_out;
          const int ox = rem - oy * p.w_out;
          const int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
          pix0[i] = (img * p.h_in + iy0) * p.w_in + ix0;
          // valid taps form a rectangle [ky_lo, ky_hi) x [kx_lo, kx_hi)
          const int kx_lo = max(0, -ix0), kx_hi = min(p.ksize, p.w_in - ix0);
          const int ky_lo = max(0, -iy0), ky_hi = min(p.ksize, p.h_in - iy0);
          uint32_t msk = 0u;
          if (kx_hi > kx_lo) {
            const uint32_t xm = ((1u << (kx_hi - kx_lo)) - 1u) << kx_lo;
            for (int ky = ky_lo; ky < ky_hi; ++ky) msk |= xm << (ky * p.ksize);
          }
          vmask[i] = msk;
        }
      }
    }
    // running (tap, channel) position of this lane's chunk: k = kb*64 + j*8
    int tap = (j * 8) / p.ctot;
    int c = (j * 8) - tap * p.ctot;
    int ky = tap / p.ksize, kx = tap - ky * p.ksize;
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % S;
      const uint32_t ph = (uint32_t)(kb / S) & 1u;
      const bool from0 = c < p.c0;
      const __nv_bfloat16* sbase = from0 ? p.a0 + c : p.a1 + (c - p.c0);
      const int cs = from0 ? p.c0 : p.c1;
      const int tapoff = ky * p.w_in + kx;
      const uint32_t tapbit = tap < ntaps ? (1u << tap) : 0u;  // K tail -> zero fill
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

    // =============================== epilogue ===========================================
    // thread <-> tile row (TMEM lane) 32*(warp & 3) + lane; the two warps sharing a lane quarter
    // split the columns in alternating 32-wide chunks
    const int q = warp & 3, half = warp >> 2;
    const int m = m0 + q * 32 + lane;
    const bool row_ok = m < p.M;
    mbar_wait(bar_acc, 0);
    tcgen05_fence_after();
    const uint32_t lane_taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int cb = half * 32; cb < BN; cb += 64) {
      uint32_t raw[32];
      tmem_ld_32x32b_x32(lane_taddr + (uint32_t)cb, raw);
      tmem_ld_wait();
      if (row_ok) epilogue_row32(p, m, n0 + cb, raw);
    }
    tcgen05_fence_before();
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

template <int BN, bool kBTma>
int launch(const CUtensorMap& tmap, const TcParams& p, cudaStream_t s) {
  using Cfg = TileCfg<BN>;
  auto kern = gemm_tc_kernel<BN, kBTma>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    BDE_REQUIRE(e == cudaSuccess, "bde_gemm(tcgen05): smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(p.M, BM), (unsigned)(p.N / BN));
  kern<<<grid, kThreads, Cfg::kSmemBytes, s>>>(tmap, p);
  return check_launch("gemm_tc_kernel");
}

}  // namespace

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
  BDE_REQUIRE(p.c0 % 8 == 0 && p.c1 % 8 == 0, "bde_gemm(tcgen05): channel counts must be multiples of 8");
  BDE_REQUIRE(p.ksize * p.ksize <= 32, "bde_gemm(tcgen05): kernel size up to 5x5");
  BDE_REQUIRE(p.w_ld % BK == 0 && p.w_ld >= p.num_kb * BK, "bde_gemm(tcgen05): w_ld must be a zero-padded multiple of 64");
  BDE_REQUIRE(p.N % 32 == 0, "bde_gemm(tcgen05): N must be a multiple of 32");
  BDE_REQUIRE((size_t)d->n_img * d->h_in * d->w_in < ((size_t)1 << 31), "bde_gemm(tcgen05): input pixel count overflows int32");
  BDE_REQUIRE((((uintptr_t)d->a0) & 15) == 0 && (((uintptr_t)d->a1) & 15) == 0 && (((uintptr_t)d->w) & 127) == 0,
              "bde_gemm(tcgen05): operands must be 16-byte (weights 128-byte) aligned");
  if (p.M == 0) return 0;
  // tile width: widest tile that still gives every SM work
  int bn = 32;
  if (p.N % 128 == 0) bn = 128;
  else if (p.N % 64 == 0) bn = 64;
  const size_t m_tiles = ceil_div(p.M, BM);
  if (p.N % 256 == 0 && m_tiles * (p.N / 256) >= 2 * (size_t)kNumSMs) bn = 256;
  const bool tma = b_via_tma();
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (tma) {
    int rc = get_weight_tmap(d->w, p.N, p.w_ld, bn, &tmap);
    if (rc != 0) return rc;
  }
#define BDE_TC_LAUNCH(BN_)                                                   \
  case BN_:                                                                  \
    return tma ? launch<BN_, true>(tmap, p, s) : launch<BN_, false>(tmap, p, s);
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
