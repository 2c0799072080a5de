// Shared helpers for libbde2vid_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/bde2vid.h"

namespace bde {

// thread-local last-error string exposed through bde_last_error()
void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: %s", what, cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

#define BDE_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      bde::set_error(__VA_ARGS__);        \
      return -1;                          \
    }                                     \
  } while (0)

constexpr int kNumSMs = 148;  // B200

__host__ __device__ inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

// ---- storage-type helpers: fp32 math everywhere, float or bf16 storage --------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// load 4 consecutive elements as float4 (pointer must be 4-element aligned)
template <typename T> __device__ __forceinline__ float4 load4(const T* p);
template <> __device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T> __device__ __forceinline__ void store4(T* p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

// ---- activations ---------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) { return tanhf(x); }
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case BDE_ACT_RELU: return fmaxf(v, 0.0f);
    case BDE_ACT_RELU6: return fminf(fmaxf(v, 0.0f), 6.0f);
    case BDE_ACT_GELU: return gelu_f(v);
    case BDE_ACT_SIGMOID: return sigmoid_f(v);
    default: return v;
  }
}

// ConvLSTM pointwise update (model/BDE2VID/submodules.py:320-332); gate order in, remember, out, cell
__device__ __forceinline__ void lstm_update(float gi, float gf, float go, float gg, float c_prev,
                                            float& h, float& c) {
  c = sigmoid_f(gf) * c_prev + sigmoid_f(gi) * tanh_f(gg);
  h = sigmoid_f(go) * tanh_f(c);
}

// true exactly once per (current device, key): guards the per-device cudaFuncSetAttribute calls (api.cu, mutex-protected)
bool first_use_on_device(const void* key);
// multiProcessorCount of the current device (cached per ordinal)
int device_sm_count();

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------------------
// The sequential attention chain is 10 small launches per frame (attention + MLP per block), each one wave or less: with
// the attribute below the NEXT kernel's CTAs may start while the current one drains and run their prologue (barrier init,
// TMEM allocation, index tables, weight prefetch -- everything that does not depend on the previous kernel's output), then
// block in pdl_wait() until the previous grid has completed and its writes are visible.  Rules kept in every such kernel:
// nothing produced by an earlier kernel is read, and nothing global is written, by a thread that has not passed pdl_wait().
// pdl_wait() returns at once when the kernel was launched without the attribute.  BDE2VID_PDL=0 disables the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();   // api.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, int cluster, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// implemented per translation unit
int gemm_simt(const bde_gemm_desc* d, cudaStream_t s);
int gemm_tcgen05(const bde_gemm_desc* d, cudaStream_t s);

}  // namespace bde
