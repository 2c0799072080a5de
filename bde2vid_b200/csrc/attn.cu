// Window multi-head attention core: softmax(q k^T + rel_pos_bias) v per (window, head).
// Reference: WindowAttention3D.forward, model/BDE2VID/DTransformer.py:192-203.
// head_dim is 4..16 on this path (C = 64..256 over 16 heads): far too thin for tensor-core tiles,
// so this runs on CUDA cores with K/V of one (window, head) staged in shared memory.
#include "common.cuh"

namespace bde {

template <typename T, int HD>
__global__ void __launch_bounds__(64) window_attention_kernel(
    const T* __restrict__ q, const T* __restrict__ kv, const float* __restrict__ bias /* [heads][n_kv][n_q] */,
    int n_q, int n_kv, int c, T* __restrict__ out) {
  extern __shared__ float smem[];
  float* ks = smem;                 // [n_kv][HD]
  float* vs = smem + n_kv * HD;     // [n_kv][HD]
  const int head = blockIdx.x, win = blockIdx.y;
  const T* kvw = kv + (size_t)win * n_kv * 2 * c;
  for (int i = threadIdx.x; i < n_kv * HD; i += blockDim.x) {
    int n = i / HD, j = i % HD;
    ks[i] = to_f32<T>(kvw[(size_t)n * 2 * c + head * HD + j]);
    vs[i] = to_f32<T>(kvw[(size_t)n * 2 * c + c + head * HD + j]);
  }
  __syncthreads();
  const float* bh = bias + (size_t)head * n_kv * n_q;
  for (int m = threadIdx.x; m < n_q; m += blockDim.x) {
    float qv[HD];
    const T* qp = q + ((size_t)win * n_q + m) * c + head * HD;
#pragma unroll
    for (int j = 0; j < HD; ++j) qv[j] = to_f32<T>(qp[j]);
    // pass 1: row maximum
    float mx = -INFINITY;
    for (int n = 0; n < n_kv; ++n) {
      float s = bh[(size_t)n * n_q + m];
#pragma unroll
      for (int j = 0; j < HD; ++j) s = fmaf(qv[j], ks[n * HD + j], s);
      mx = fmaxf(mx, s);
    }
    // pass 2: exp, sum, weighted values
    float l = 0.f, acc[HD];
#pragma unroll
    for (int j = 0; j < HD; ++j) acc[j] = 0.f;
    for (int n = 0; n < n_kv; ++n) {
      float s = bh[(size_t)n * n_q + m];
#pragma unroll
      for (int j = 0; j < HD; ++j) s = fmaf(qv[j], ks[n * HD + j], s);
      float p = __expf(s - mx);
      l += p;
#pragma unroll
      for (int j = 0; j < HD; ++j) acc[j] = fmaf(p, vs[n * HD + j], acc[j]);
    }
    float inv = 1.0f / l;
    T* op = out + ((size_t)win * n_q + m) * c + head * HD;
#pragma unroll
    for (int j = 0; j < HD; ++j) op[j] = from_f32<T>(acc[j] * inv);
  }
}

template <typename T>
static int launch_attn(const void* q, const void* kv, const float* bias, int n_win, int n_q, int n_kv, int c, int heads,
                       void* out, cudaStream_t s) {
  const int hd = c / heads;
  dim3 grid(heads, n_win);
  size_t smem = (size_t)2 * n_kv * hd * sizeof(float);
  BDE_REQUIRE(smem <= 48 * 1024, "bde_window_attention: n_kv*head_dim too large for shared memory");
#define BDE_ATTN_CASE(HD)                                                                                         \
  case HD:                                                                                                        \
    window_attention_kernel<T, HD><<<grid, 64, smem, s>>>((const T*)q, (const T*)kv, bias, n_q, n_kv, c, (T*)out); \
    break;
  switch (hd) {
    BDE_ATTN_CASE(2)
    BDE_ATTN_CASE(4)
    BDE_ATTN_CASE(8)
    BDE_ATTN_CASE(16)
    BDE_ATTN_CASE(32)
    default:
      BDE_REQUIRE(false, "bde_window_attention: unsupported head_dim %d", hd);
  }
#undef BDE_ATTN_CASE
  return check_launch("window_attention_kernel");
}

}  // namespace bde

using namespace bde;

extern "C" int bde_window_attention(const void* q, const void* kv, const float* bias, int n_win, int n_q, int n_kv,
                                    int c, int heads, void* out, int dtype, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  BDE_REQUIRE(heads > 0 && c % heads == 0, "bde_window_attention: c must be divisible by heads");
  if (n_win == 0) return 0;
  if (dtype == BDE_F32) return launch_attn<float>(q, kv, bias, n_win, n_q, n_kv, c, heads, out, s);
  return launch_attn<__nv_bfloat16>(q, kv, bias, n_win, n_q, n_kv, c, heads, out, s);
}
