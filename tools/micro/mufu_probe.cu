// Throughput of MUFU.EX2 vs an FMA-pipe polynomial exp2 (and mixes) on one SM sub-partition set: is the XU pipe of
// sm_100a 16 or 8 lanes / clk / SM, and how much of a softmax's exponentials can move to the FMA pipe for free?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_probe tools/micro/mufu_probe.cu && /tmp/mufu_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ float ex2_mufu(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for x in [-125, 0]: round-to-nearest split + degree-3 minimax on [-0.5, 0.5] + exponent add (FMA / ALU pipes only)
__device__ __forceinline__ float ex2_poly(float x) {
  const float t = x + 12582912.0f;            // 1.5 * 2^23: integer part lands in the low mantissa bits
  const float xi = t - 12582912.0f;
  const float f = x - xi;
  float p = fmaf(f, 0.0555041086f, 0.2402265070f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int MODE>   // 0: all MUFU, 1: all poly, 2: 3 MUFU + 1 poly, 3: 1 MUFU + 1 poly, 4: 7 MUFU + 1 poly
__global__ void k(float* out, int iters, float seed) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed - 0.01f * (threadIdx.x & 7) - 0.1f * i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      bool poly = MODE == 1 || (MODE == 2 && (i & 3) == 3) || (MODE == 3 && (i & 1)) || (MODE == 4 && i == 7);
      const float y = poly ? ex2_poly(v[i]) : ex2_mufu(v[i]);
      v[i] = y - 1.5f;   // keeps the argument in [-1.5, -0.5]: a dependent chain per register, 8 chains per thread
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* out) {
  const int iters = 4096, blocks = 148 * 2, threads = 512;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<blocks, threads>>>(out, 16, -1.0f);
  cudaEventRecord(a);
  k<MODE><<<blocks, threads>>>(out, iters, -1.0f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double n = (double)blocks * threads * iters * 8;
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("%-22s %.3f ms  %.1f G exp2/s  = %.2f per clk per SM at %d MHz (nominal)\n", name, ms, n / ms / 1e6, n / (ms * 1e-3) / 148 / (khz * 1e3), khz / 1000);
}

int main() {
  float* out; cudaMalloc(&out, 148 * 2 * 512 * 4);
  run<0>("all MUFU.EX2", out);
  run<1>("all polynomial", out);
  run<4>("7 MUFU : 1 poly", out);
  run<2>("3 MUFU : 1 poly", out);
  run<3>("1 MUFU : 1 poly", out);
  // accuracy of the polynomial
  printf("done %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
