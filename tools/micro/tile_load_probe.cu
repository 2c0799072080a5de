// How long does one CTA need to pull a [128 rows x 1 KB] fp32 tile (L2-resident) into registers / shared memory, for the
// access patterns the LayerNorm producers use?  (The round-2 phase counters put the LN phase of the fused MLP kernels at
// 12 K cycles for 128 KB, i.e. ~10 B / clk / SM, far below the ~42 B / clk / SM the L2 can deliver.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tile_load_probe tools/micro/tile_load_probe.cu && /tmp/tile_load_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int ROWS = 128, ROWF = 256;   // floats per row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// MODE 0: lane&7 owns 8 floats of a 64-float slab, 4 rows per warp instruction (the LN producers' pattern), all loads hoisted
// MODE 1: one row per warp instruction pair, lanes contiguous (16 B x 32 = 512 B per instruction)
// MODE 2: MODE 0 with ld.global.cg (bypass L1)
// MODE 3: bulk async copies (cp.async.bulk, 8 KB pieces) into shared memory + mbarrier
// MODE 4: MODE 0 but the loads of one row pass are consumed before the next pass is issued (dependent passes)
__device__ __forceinline__ float4 ld_plain(const float* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_nc(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_nc_na(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// pattern 0 with a chosen load flavour (FL 0 = ld.global, 1 = ld.global.nc, 2 = ld.global.nc.L1::no_allocate), 9 warps: warp 8 optionally
// streams `bulk_kb` KB of other data into shared memory with cp.async.bulk at the same time (the weight ring of the MLP kernels)
template <int FL>
__global__ void __launch_bounds__(288, 1) k2(const float* x, const float* wts, float* out, long long* cyc, int bulk_kb) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* tile = x + (size_t)blockIdx.x * ROWS * ROWF;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  float acc = 0.f;
  if (warp == 8) {
    if (lane == 0 && bulk_kb > 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bulk_kb * 1024) : "memory");
      for (int i = 0; i < bulk_kb / 16; ++i)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem) + i * 16384),
                     "l"(wts + i * 4096), "r"(16384), "r"(smem_u32(&bar))
                     : "memory");
    }
  } else {
    const int j = lane & 7, rsub = lane >> 3;
    float4 v[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float* src = tile + (size_t)(warp * 16 + i * 4 + rsub) * ROWF + j * 8;
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        v[i][2 * kb] = FL == 0 ? ld_plain(src + kb * 64) : (FL == 1 ? ld_nc(src + kb * 64) : ld_nc_na(src + kb * 64));
        v[i][2 * kb + 1] = FL == 0 ? ld_plain(src + kb * 64 + 4) : (FL == 1 ? ld_nc(src + kb * 64 + 4) : ld_nc_na(src + kb * 64 + 4));
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc += v[i][e].x + v[i][e].y + v[i][e].z + v[i][e].w;
  }
  // the timestamp must not be speculated above the uses of the loaded values: make it depend on acc
  long long t1 = clock64();
  if (acc == 123.456f) t1 = 0;
  out[blockIdx.x * 288 + threadIdx.x] = acc;
  if (bulk_kb > 0) {
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
  }
  __syncthreads();
  const long long t2 = clock64();
  if (threadIdx.x == 0) { cyc[blockIdx.x * 2] = t1 - t0; cyc[blockIdx.x * 2 + 1] = t2 - t0; }
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(const float* __restrict__ x, float* out, long long* cyc, int warps_active) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* tile = x + (size_t)blockIdx.x * ROWS * ROWF;
  __shared__ uint64_t bar;
  if (MODE == 3 && threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  float acc = 0.f;
  if (warp < warps_active) {
    if (MODE == 0 || MODE == 2 || MODE == 4) {
      const int j = lane & 7, rsub = lane >> 3;
      float4 v[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float* src = tile + (size_t)(warp * 16 + i * 4 + rsub) * ROWF + j * 8;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          if (MODE == 2) {
            v[i][2 * kb] = __ldcg(reinterpret_cast<const float4*>(src + kb * 64));
            v[i][2 * kb + 1] = __ldcg(reinterpret_cast<const float4*>(src + kb * 64 + 4));
          } else {
            v[i][2 * kb] = *reinterpret_cast<const float4*>(src + kb * 64);
            v[i][2 * kb + 1] = *reinterpret_cast<const float4*>(src + kb * 64 + 4);
          }
        }
        if (MODE == 4) {
#pragma unroll
          for (int e = 0; e < 8; ++e) acc += v[i][e].x + v[i][e].y + v[i][e].z + v[i][e].w;
          acc = __shfl_xor_sync(0xffffffffu, acc, 1) + acc;
          asm volatile("" ::: "memory");
        }
      }
      if (MODE != 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc += v[i][e].x + v[i][e].y + v[i][e].z + v[i][e].w;
      }
    } else if (MODE == 1) {
      float4 v[16][2];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float* src = tile + (size_t)(warp * 16 + i) * ROWF + lane * 4;
        v[i][0] = *reinterpret_cast<const float4*>(src);
        v[i][1] = *reinterpret_cast<const float4*>(src + 128);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) acc += v[i][0].x + v[i][0].y + v[i][0].z + v[i][0].w + v[i][1].x + v[i][1].y + v[i][1].z + v[i][1].w;
    } else if (MODE == 3) {
      if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(ROWS * ROWF * 4) : "memory");
        for (int i = 0; i < 16; ++i)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem) + i * 8192),
                       "l"(tile + i * 2048), "r"(8192), "r"(smem_u32(&bar))
                       : "memory");
      }
      uint32_t done = 0;
      while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
      acc = reinterpret_cast<const float*>(smem)[threadIdx.x];
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * 256 + threadIdx.x] = acc;
  __syncthreads();
  const long long t2 = clock64();
  if (threadIdx.x == 0) { cyc[blockIdx.x * 2] = t1 - t0; cyc[blockIdx.x * 2 + 1] = t2 - t0; }
}

__global__ void fill(float* x, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = (float)(i & 1023) * 1e-3f;
}

template <int MODE>
void run(const char* name, const float* x, float* out, long long* cyc, int ctas, int warps, size_t n) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  long long h[148 * 2];
  double a = 0, b = 0;
  for (int rep = 0; rep < 3; ++rep) {
    fill<<<148 * 4, 256>>>(const_cast<float*>(x), n);   // the previous kernel wrote the tile (dirty lines in L2)
    k<MODE><<<ctas, 256, MODE == 3 ? 132 * 1024 : 0>>>(x, out, cyc, warps);
    cudaMemcpy(h, cyc, ctas * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
    a = b = 0;
    for (int i = 0; i < ctas; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
  }
  printf("%-62s %3d CTAs x %d warps: thread 0 done after %6.0f cycles, CTA after %6.0f  (%s)\n", name, ctas, warps, a / ctas, b / ctas,
         cudaGetErrorString(cudaGetLastError()));
}

template <int FL>
void run2(const char* name, const float* x, const float* wts, float* out, long long* cyc, int ctas, int bulk_kb, size_t n, int smem_kb = 196) {
  cudaFuncSetAttribute(k2<FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  long long h[148 * 2];
  double a = 0, b = 0;
  for (int rep = 0; rep < 3; ++rep) {
    fill<<<148 * 4, 256>>>(const_cast<float*>(x), n);
    k2<FL><<<ctas, 288, smem_kb * 1024>>>(x, wts, out, cyc, bulk_kb);
    cudaMemcpy(h, cyc, ctas * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
    a = b = 0;
    for (int i = 0; i < ctas; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
  }
  printf("%-30s smem %3d KB + %3d KB bulk copies  %3d CTAs: thread 0 has its data after %6.0f cycles, CTA (incl. bulk) after %6.0f  (%s)\n", name, smem_kb, bulk_kb, ctas,
         a / ctas, b / ctas, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const size_t n = (size_t)148 * ROWS * ROWF;
  float *x, *out; long long* cyc;
  cudaMalloc(&x, n * 4); cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 2 * 8);
  for (int ctas : {148, 24}) {
    run<0>("0: 8 lanes x 32 B per 256 B slab, 4 rows / instr, hoisted", x, out, cyc, ctas, 8, n);
    run<4>("4: same pattern, one row pass in flight per warp", x, out, cyc, ctas, 8, n);
    run<2>("2: pattern 0 with ld.global.cg", x, out, cyc, ctas, 8, n);
    run<1>("1: one row per instruction pair, lanes contiguous", x, out, cyc, ctas, 8, n);
    run<3>("3: cp.async.bulk 16 x 8 KB into shared memory", x, out, cyc, ctas, 8, n);
  }
  float* wts; cudaMalloc(&wts, 1 << 20); cudaMemset(wts, 0, 1 << 20);
  float* out2; cudaMalloc(&out2, 148 * 288 * 4);
  for (int smem_kb : {100, 196, 212, 220, 225}) {
    run2<0>("ld.global (pattern 0)", x, wts, out2, cyc, 92, 96, n, smem_kb);
    run2<1>("ld.global.nc", x, wts, out2, cyc, 92, 96, n, smem_kb);
    run2<2>("ld.global.nc.L1::no_allocate", x, wts, out2, cyc, 92, 96, n, smem_kb);
  }
  run<0>("0: single warp (16 rows = 16 KB)", x, out, cyc, 148, 1, n);
  run<1>("1: single warp (16 rows = 16 KB)", x, out, cyc, 148, 1, n);
  printf("done %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
