// Window multi-head attention core on tensor cores (bf16 mma.sync m16n8k16, fp32 accumulate).
//
//   out[w, m, h*hd:(h+1)*hd] = softmax_n( q[w,m,h] . k[w,n,h] + bias[h,m,n] ) v[w,n,h]
//
// Reference: WindowAttention3D.forward, model/BDE2VID/DTransformer.py:192-203.
// The per-head GEMMs are tiny (49 x 147 x head_dim with head_dim 4..16): a tcgen05 tile (M=128 per
// CTA, accumulator in TMEM) would be mostly padding and the softmax needs the scores in registers
// anyway, so this kernel uses the warp-level mma.sync path: one warp owns 16 query rows of one
// (window, head), keeps the 16 x n_kv score tile in registers (flash-attention style) and feeds it
// straight back as the A operand of the P.V product.  The relative-position bias is the C operand
// of the first MMA, so adding it is free.
//
// CTA = 8 warps = 2 heads x 4 query tiles.  It stages its bias slice in shared memory once and
// then loops over windows with a 2-deep cp.async pipeline on the K / V slices of its two heads
// (V stays in its natural [key][channel] layout; ldmatrix.trans produces the B fragments).
#include "common.cuh"

namespace bde {

namespace {

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t smem_addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(smem_addr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// element strides of the q / k / v operands (they may live in one merged [rows, 3c] qkv buffer)
struct AttnStrides {
  int q_ld, q_rows_per_win, q_row0;   // q[w, m, :]  = q  + ((w * q_rows_per_win + q_row0 + m) * q_ld)
  int kv_ld, kv_rows_per_win, v_off;  // k[w, n, :]  = kv + ((w * kv_rows_per_win + n) * kv_ld),  v = k + v_off
};

template <int NT>
struct AttnSmem {
  static constexpr int kKeys = NT * 8;
  static constexpr int kBiasStride = (kKeys % 32 == 8) ? kKeys : kKeys + ((8 - kKeys % 32 + 32) % 32);  // == 8 (mod 32) floats
  static constexpr int kKSteps = (NT + 1) / 2;  // k16 steps of P.V
  static constexpr int kRows = kKSteps * 16;    // staged key rows (zero beyond n_kv)
};

// 4 query tiles of 16 rows (n_q <= 64); NT key tiles of 8 (n_kv <= 8*NT); HD = head_dim (4, 8, 16)
template <int HD, int NT>
__global__ void __launch_bounds__(256, 2) window_attention_mma_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kv,
    const float* __restrict__ bias /* [heads][64][kBiasStride], -1e30 beyond n_kv */, int n_win, int n_q, int n_kv,
    int c, AttnStrides st, __nv_bfloat16* __restrict__ out) {
  using SM = AttnSmem<NT>;
  constexpr int KS = 2 * HD + 8;           // staged row: both heads' slices + pad (bf16 elements)
  constexpr int CH = (2 * HD) / 8;         // 16-byte chunks per row and tensor
  constexpr int kBufElems = SM::kRows * KS;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* bias_s = reinterpret_cast<float*>(smem_raw);                                       // [2][64][kBiasStride]
  __nv_bfloat16* ks = reinterpret_cast<__nv_bfloat16*>(bias_s + 2 * 64 * SM::kBiasStride);  // [2 bufs][kRows][KS]
  __nv_bfloat16* vs = ks + 2 * kBufElems;                                                   // [2 bufs][kRows][KS]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int hh = warp >> 2, mt = warp & 3;
  const int h0 = blockIdx.x * 2;  // first head of this CTA's pair
  const int head = h0 + hh;

  // bias slice of the two heads -> smem (once per CTA); zero both K/V buffers once so that the
  // rows beyond n_kv (never written again) are finite zeros
  {
    const float* src = bias + (size_t)h0 * 64 * SM::kBiasStride;
    for (int i = tid; i < 2 * 64 * SM::kBiasStride / 4; i += 256)
      reinterpret_cast<float4*>(bias_s)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
    for (int i = tid; i < 4 * kBufElems / 8; i += 256) reinterpret_cast<uint4*>(ks)[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();

  const float* my_bias = bias_s + ((size_t)hh * 64 + mt * 16) * SM::kBiasStride;
  const int row0 = mt * 16 + g, row1 = row0 + 8;
  const uint32_t ks_u32 = (uint32_t)__cvta_generic_to_shared(ks);
  const uint32_t vs_u32 = (uint32_t)__cvta_generic_to_shared(vs);

  auto stage = [&](int w, int buf) {  // asynchronous copy of window w's K / V slices into buffer `buf`
    const __nv_bfloat16* kvw = kv + (size_t)w * st.kv_rows_per_win * st.kv_ld + h0 * HD;
    for (int i = tid; i < n_kv * CH; i += 256) {
      const int n = i / CH, ch = i - n * CH;
      const __nv_bfloat16* rowp = kvw + (size_t)n * st.kv_ld + ch * 8;
      const uint32_t off = (uint32_t)((buf * kBufElems + n * KS + ch * 8) * 2);
      cp_async16(ks_u32 + off, rowp, 16u);
      cp_async16(vs_u32 + off, rowp + st.v_off, 16u);
    }
    cp_async_commit();
  };
  auto load_q = [&](int w, uint32_t (&qa)[4]) {
    // a0:(row0, k 2t..2t+1) a1:(row1, same k) a2:(row0, k 2t+8..) a3:(row1, k 2t+8..)
    const __nv_bfloat16* qw = q + ((size_t)w * st.q_rows_per_win + st.q_row0) * st.q_ld + head * HD;
    qa[0] = qa[1] = qa[2] = qa[3] = 0u;
    if (2 * t < HD) {
      if (row0 < n_q) qa[0] = __ldg(reinterpret_cast<const uint32_t*>(qw + (size_t)row0 * st.q_ld + 2 * t));
      if (row1 < n_q) qa[1] = __ldg(reinterpret_cast<const uint32_t*>(qw + (size_t)row1 * st.q_ld + 2 * t));
    }
    if (2 * t + 8 < HD) {
      if (row0 < n_q) qa[2] = __ldg(reinterpret_cast<const uint32_t*>(qw + (size_t)row0 * st.q_ld + 2 * t + 8));
      if (row1 < n_q) qa[3] = __ldg(reinterpret_cast<const uint32_t*>(qw + (size_t)row1 * st.q_ld + 2 * t + 8));
    }
  };

  int w = blockIdx.y;
  uint32_t qa[4] = {0u, 0u, 0u, 0u};
  if (w < n_win) {
    stage(w, 0);
    load_q(w, qa);
  }
  int buf = 0;
  for (; w < n_win; w += gridDim.y, buf ^= 1) {
    const int wn = w + gridDim.y;
    uint32_t qn[4] = {0u, 0u, 0u, 0u};
    cp_async_wait<0>();   // this thread's copies for window w have landed
    __syncthreads();      // ... everyone's have, and everyone is done reading buffer buf^1 (window w - stride)
    if (wn < n_win) {
      stage(wn, buf ^ 1);  // overlaps with the math below
      load_q(wn, qn);
    }
    const __nv_bfloat16* kb = ks + buf * kBufElems;
    const uint32_t vb_u32 = vs_u32 + (uint32_t)(buf * kBufElems * 2);

    // ---- S = bias + Q K^T ---------------------------------------------------------------------
    float s[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const float2 b0 = *reinterpret_cast<const float2*>(my_bias + (size_t)g * SM::kBiasStride + j * 8 + 2 * t);
      const float2 b1 = *reinterpret_cast<const float2*>(my_bias + (size_t)(g + 8) * SM::kBiasStride + j * 8 + 2 * t);
      s[j][0] = b0.x; s[j][1] = b0.y; s[j][2] = b1.x; s[j][3] = b1.y;
      // B fragment: b0 = K[key 8j+g][2t..2t+1], b1 = K[key 8j+g][2t+8..2t+9]
      const __nv_bfloat16* kr = kb + (j * 8 + g) * KS + hh * HD;
      uint32_t kb0 = 0u, kb1 = 0u;
      if (2 * t < HD) kb0 = *reinterpret_cast<const uint32_t*>(kr + 2 * t);
      if (2 * t + 8 < HD) kb1 = *reinterpret_cast<const uint32_t*>(kr + 2 * t + 8);
      mma_bf16_16816(s[j], qa, kb0, kb1);
    }

    // ---- softmax over keys (rows row0 and row1; quad reduction over t) --------------------------
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    constexpr float kLog2e = 1.4426950408889634f;
    const float m0s = mx0 * kLog2e, m1s = mx1 * kLog2e;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      s[j][0] = exp2f(fmaf(s[j][0], kLog2e, -m0s));
      s[j][1] = exp2f(fmaf(s[j][1], kLog2e, -m0s));
      s[j][2] = exp2f(fmaf(s[j][2], kLog2e, -m1s));
      s[j][3] = exp2f(fmaf(s[j][3], kLog2e, -m1s));
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

    // ---- O = P V : P (bf16) from the score registers, V fragments via ldmatrix.trans ------------
    // HD >= 8: channel tile v covers channels hh*HD + 8v ..+7 of the staged row.
    // HD == 4: the single 8-wide tile holds both heads' channels; this warp uses columns 4*hh..4*hh+3.
    constexpr int NV = (HD + 7) / 8;
    float o[NV][4];
#pragma unroll
    for (int v = 0; v < NV; ++v) o[v][0] = o[v][1] = o[v][2] = o[v][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < SM::kKSteps; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      if (2 * kk + 1 < NT) {
        pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      } else {
        pa[2] = 0u;
        pa[3] = 0u;
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int ch0 = (HD >= 8) ? hh * HD + 8 * v : 0;
        // lanes 0-7: rows (keys) 16kk + lane of the first 8x8 matrix; lanes 8-15: keys 16kk + 8 + (lane - 8)
        const uint32_t addr = vb_u32 + (uint32_t)(((kk * 16 + (lane & 15)) * KS + ch0) * 2);
        uint32_t vb0, vb1;
        ldmatrix_x2_trans(vb0, vb1, addr);
        mma_bf16_16816(o[v], pa, vb0, vb1);
      }
    }
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    __nv_bfloat16* ow = out + (size_t)w * n_q * c + head * HD;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      // accumulator columns 2t, 2t+1 of tile v  ->  channel (within the head)
      int col = 8 * v + 2 * t;
      bool mine = col < HD;
      if (HD == 4) {
        mine = (t >> 1) == hh;   // columns 4*hh .. 4*hh+3 of the shared tile
        col = 2 * (t & 1);
      }
      if (mine) {
        if (row0 < n_q) *reinterpret_cast<uint32_t*>(ow + (size_t)row0 * c + col) = pack_bf16(o[v][0] * inv0, o[v][1] * inv0);
        if (row1 < n_q) *reinterpret_cast<uint32_t*>(ow + (size_t)row1 * c + col) = pack_bf16(o[v][2] * inv1, o[v][3] * inv1);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) qa[i] = qn[i];
  }
  cp_async_wait<0>();
}

template <int HD, int NT>
int launch_mma(const void* q, const void* kv, const float* bias, int n_win, int n_q, int n_kv, int c, int heads,
               const AttnStrides& st, void* out, cudaStream_t s) {
  using SM = AttnSmem<NT>;
  constexpr int KS = 2 * HD + 8;
  const size_t smem = (size_t)2 * 64 * SM::kBiasStride * sizeof(float) + (size_t)4 * SM::kRows * KS * 2;
  auto kern = window_attention_mma_kernel<HD, NT>;
  if (first_use_on_device((const void*)kern)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    BDE_REQUIRE(e == cudaSuccess, "bde_window_attention(mma): smem attribute: %s", cudaGetErrorString(e));
  }
  // two CTAs per SM: (heads/2) head pairs x window groups ~ one wave of 2 x 148 CTAs
  int groups = (2 * device_sm_count()) / (heads / 2);
  if (groups < 1) groups = 1;
  if (groups > n_win) groups = n_win;
  dim3 grid(heads / 2, groups);
  kern<<<grid, 256, smem, s>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)kv, bias, n_win, n_q, n_kv, c, st,
                               (__nv_bfloat16*)out);
  return check_launch("window_attention_mma_kernel");
}

}  // namespace

// row pitch (floats) of the padded bias the mma kernel expects for n_kv keys; 0 if unsupported
int attn_mma_bias_stride(int n_kv) {
  const int nt = (n_kv + 7) / 8;
  if (nt <= 7) return AttnSmem<7>::kBiasStride;
  if (nt <= 13) return AttnSmem<13>::kBiasStride;
  if (nt <= 19) return AttnSmem<19>::kBiasStride;
  return 0;
}

int window_attention_mma(const void* q, const void* kv, const float* bias, int n_win, int n_q, int n_kv, int c, int heads,
                         const AttnStrides& st, void* out, cudaStream_t s) {
  const int hd = c / heads;
  const int nt = (n_kv + 7) / 8;
  BDE_REQUIRE(n_q <= 64 && nt <= 19 && heads % 2 == 0 && (hd == 4 || hd == 8 || hd == 16) && c % 8 == 0,
              "bde_window_attention_mma: unsupported shape (n_q=%d n_kv=%d hd=%d)", n_q, n_kv, hd);
  BDE_REQUIRE(st.q_ld % 2 == 0 && st.kv_ld % 8 == 0 && st.v_off % 8 == 0, "bde_window_attention_mma: misaligned strides");
#define BDE_ATTN_MMA(HD_, NT_) return launch_mma<HD_, NT_>(q, kv, bias, n_win, n_q, n_kv, c, heads, st, out, s)
#define BDE_ATTN_MMA_HD(NT_)              \
  switch (hd) {                           \
    case 4: BDE_ATTN_MMA(4, NT_);         \
    case 8: BDE_ATTN_MMA(8, NT_);         \
    default: BDE_ATTN_MMA(16, NT_);       \
  }
  if (nt <= 7) { BDE_ATTN_MMA_HD(7) }
  if (nt <= 13) { BDE_ATTN_MMA_HD(13) }
  BDE_ATTN_MMA_HD(19)
#undef BDE_ATTN_MMA_HD
#undef BDE_ATTN_MMA
}

}  // namespace bde

using namespace bde;

extern "C" int bde_window_attention_mma_bias_stride(int n_kv) { return attn_mma_bias_stride(n_kv); }

extern "C" int bde_window_attention_mma(const void* q, const void* kv, const float* bias_padded, int n_win, int n_q,
                                        int n_kv, int c, int heads, void* out, void* stream) {
  if (n_win == 0) return 0;
  AttnStrides st = {c, n_q, 0, 2 * c, n_kv, c};
  return window_attention_mma(q, kv, bias_padded, n_win, n_q, n_kv, c, heads, st, out, (cudaStream_t)stream);
}

extern "C" int bde_window_attention_mma_qkv(const void* qkv, const float* bias_padded, int n_win, int n_q, int n_kv,
                                            int q_row0, int c, int heads, void* out, void* stream) {
  if (n_win == 0) return 0;
  // rows of window w: [w * n_kv, (w + 1) * n_kv); columns [0, c) = q, [c, 2c) = k, [2c, 3c) = v
  AttnStrides st = {3 * c, n_kv, q_row0, 3 * c, n_kv, c};
  return window_attention_mma(qkv, (const __nv_bfloat16*)qkv + c, bias_padded, n_win, n_q, n_kv, c, heads, st, out,
                              (cudaStream_t)stream);
}
