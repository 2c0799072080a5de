// CUDA-core implicit-GEMM engine (fp32 accumulate, float or bf16 storage).
//
// This is the exact-arithmetic ("fp32 parity mode") engine behind bde_gemm and the checker the
// tcgen05 engine is validated against on the GPU.  Same descriptor, same epilogues.
//   C[m, n] = sum_k A[m, k] * W[n, k],  m = (img, oy, ox), k = (tap, channel of source 0 | source 1)
#include "common.cuh"

namespace bde {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256;

struct GemmParams {
  const void* a0;
  const void* a1;
  const void* w;
  const float* bias;
  int c0, c1, ctot;
  int n_img, h_in, w_in, h_out, w_out, ksize, stride, pad;
  int M, N, K, w_ld, k_order;
  int epi, act, out_f32;
  void* out;
  void* out2;
  const void* residual;
  int res_mode;
  const float* c_prev;
  float* c_out;
  const int* row_map;
};

template <typename T>
__global__ void __launch_bounds__(THREADS) gemm_simt_kernel(GemmParams p) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // 16 x 16 threads, each 4 x 4 outputs
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // loader role: row lr (0..63), k-quad lq (0..3) -> 4 consecutive k
  const int lr = tid / 4, lq = tid % 4;
  const int am = m0 + lr;
  const bool a_row_ok = am < p.M;
  int img = 0, oy = 0, ox = 0;
  if (a_row_ok) {
    int hw = p.h_out * p.w_out;
    img = am / hw;
    int r = am % hw;
    oy = r / p.w_out;
    ox = r % p.w_out;
  }
  const int bn = n0 + lr;
  const bool b_row_ok = bn < p.N;
  const T* wrow = reinterpret_cast<const T*>(p.w) + (size_t)bn * p.w_ld;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ktiles = (p.K + BK - 1) / BK;

  auto load_tile = [&](int kt, float4& av, float4& bv) {
    const int k = kt * BK + lq * 4;  // K % 4 == 0 and ctot % 4 == 0 are required by the host wrapper
    av = make_float4(0.f, 0.f, 0.f, 0.f);
    bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < p.K) {
      if (b_row_ok) bv = load4<T>(wrow + k);
      if (a_row_ok) {
        int tap, c;
        if (p.k_order == 1) {  // k = (64-channel chunk, tap, channel in chunk)
          const int per_chunk = p.ksize * p.ksize * 64;
          const int chunk = k / per_chunk, rem = k - chunk * per_chunk;
          tap = rem >> 6;
          c = chunk * 64 + (rem & 63);
        } else {
          tap = k / p.ctot;
          c = k % p.ctot;
        }
        int ky = tap / p.ksize, kx = tap % p.ksize;
        int iy = oy * p.stride + ky - p.pad, ix = ox * p.stride + kx - p.pad;
        if (iy >= 0 && iy < p.h_in && ix >= 0 && ix < p.w_in) {
          size_t pix = ((size_t)img * p.h_in + iy) * p.w_in + ix;
          if (c < p.c0)
            av = load4<T>(reinterpret_cast<const T*>(p.a0) + pix * p.c0 + c);
          else
            av = load4<T>(reinterpret_cast<const T*>(p.a1) + pix * p.c1 + (c - p.c0));
        }
      }
    }
  };
  auto store_tile = [&](int buf, const float4& av, const float4& bv) {
    As[buf][lq * 4 + 0][lr] = av.x;
    As[buf][lq * 4 + 1][lr] = av.y;
    As[buf][lq * 4 + 2][lr] = av.z;
    As[buf][lq * 4 + 3][lr] = av.w;
    Bs[buf][lq * 4 + 0][lr] = bv.x;
    Bs[buf][lq * 4 + 1][lr] = bv.y;
    Bs[buf][lq * 4 + 2][lr] = bv.z;
    Bs[buf][lq * 4 + 3][lr] = bv.w;
  };

  float4 av, bv;
  load_tile(0, av, bv);
  store_tile(0, av, bv);
  __syncthreads();
  for (int kt = 0; kt < ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ktiles) load_tile(kt + 1, av, bv);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    if (kt + 1 < ktiles) {
      store_tile(buf ^ 1, av, bv);
      __syncthreads();
    }
  }

  // ---------------------------------- epilogue ----------------------------------------------
  const int nb = n0 + tx * 4;  // N % 4 == 0 (host-checked) so the 4 columns are all valid or all not
  if (nb >= p.N) return;
  float b4[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias != nullptr) {
    float4 t = *reinterpret_cast<const float4*>(p.bias + nb);
    b4[0] = t.x; b4[1] = t.y; b4[2] = t.z; b4[3] = t.w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + b4[j];
    if (p.epi == BDE_EPI_STORE) {
      size_t o = (size_t)m * p.N + nb;
      if (p.res_mode == 1 && p.residual != nullptr) {
        const T* r = reinterpret_cast<const T*>(p.residual) + o;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += to_f32(r[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], p.act);
      if (p.res_mode == 0 && p.residual != nullptr) {
        float4 r = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + o);
        v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
      }
      float4 ov = make_float4(v[0], v[1], v[2], v[3]);
      if (p.out_f32) {
        store4<float>(reinterpret_cast<float*>(p.out) + o, ov);
        if (p.out2 != nullptr) store4<T>(reinterpret_cast<T*>(p.out2) + o, ov);
      } else {
        store4<T>(reinterpret_cast<T*>(p.out) + o, ov);
      }
    } else if (p.epi == BDE_EPI_LSTM) {
      // columns nb..nb+3 = gates (in, remember, out, cell) of hidden channel nb/4
      const int hid = p.N / 4, ch = nb / 4;
      size_t o = (size_t)m * hid + ch;
      float cp = p.c_prev != nullptr ? p.c_prev[o] : 0.f;
      float h, c;
      lstm_update(v[0], v[1], v[2], v[3], cp, h, c);
      p.c_out[o] = c;
      reinterpret_cast<T*>(p.out)[o] = from_f32<T>(h);
    } else if (p.epi == BDE_EPI_GRU_UR) {
      // columns nb..nb+3 = (update, reset) of hidden channels nb/2 and nb/2 + 1 (submodules.py:371-373)
      const int hid = p.N / 2, ch = nb / 2;
      size_t o = (size_t)m * hid + ch;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float hp = p.c_prev != nullptr ? p.c_prev[o + j] : 0.f;
        p.c_out[o + j] = sigmoid_f(v[2 * j]);
        reinterpret_cast<T*>(p.out)[o + j] = from_f32<T>(hp * sigmoid_f(v[2 * j + 1]));
      }
    } else if (p.epi == BDE_EPI_GRU_OUT) {
      // h' = h (1 - u) + tanh(out_gate) u (submodules.py:374-375)
      size_t o = (size_t)m * p.N + nb;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float hp = p.c_prev != nullptr ? p.c_prev[o + j] : 0.f;
        const float u = reinterpret_cast<const float*>(p.residual)[o + j];
        const float hn = hp * (1.0f - u) + tanh_f(v[j]) * u;
        p.c_out[o + j] = hn;
        reinterpret_cast<T*>(p.out)[o + j] = from_f32<T>(hn);
      }
    } else {  // BDE_EPI_SCATTER
      int dst = p.row_map[m];
      if (dst >= 0) {
        float* o = reinterpret_cast<float*>(p.out) + (size_t)dst * p.N + nb;
        float4 cur = *reinterpret_cast<float4*>(o);
        cur.x += v[0]; cur.y += v[1]; cur.z += v[2]; cur.w += v[3];
        *reinterpret_cast<float4*>(o) = cur;
      }
    }
  }
}

}  // namespace

int gemm_simt(const bde_gemm_desc* d, cudaStream_t s) {
  GemmParams p;
  p.a0 = d->a0; p.a1 = d->a1; p.w = d->w; p.bias = d->bias;
  p.c0 = d->c0; p.c1 = d->c1; p.ctot = d->c0 + d->c1;
  p.n_img = d->n_img; p.h_in = d->h_in; p.w_in = d->w_in; p.h_out = d->h_out; p.w_out = d->w_out;
  p.ksize = d->ksize; p.stride = d->stride; p.pad = d->pad;
  p.M = d->n_img * d->h_out * d->w_out;
  p.N = d->n;
  p.K = d->ksize * d->ksize * p.ctot;
  p.w_ld = d->w_ld > 0 ? d->w_ld : p.K;
  BDE_REQUIRE(p.w_ld >= p.K && p.w_ld % 4 == 0, "bde_gemm(simt): bad w_ld");
  p.k_order = d->k_order;
  BDE_REQUIRE(p.k_order == 0 || (p.c0 % 64 == 0 && p.c1 % 64 == 0), "bde_gemm(simt): chunk-major K needs channels %% 64 == 0");
  p.epi = d->epi; p.act = d->act; p.out_f32 = d->out_f32;
  p.out = d->out; p.out2 = d->out2; p.residual = d->residual; p.res_mode = d->res_mode; p.c_prev = d->c_prev; p.c_out = d->c_out;
  p.row_map = d->row_map;
  BDE_REQUIRE(p.c0 % 4 == 0 && p.c1 % 4 == 0 && p.N % 4 == 0, "bde_gemm(simt): channels and N must be multiples of 4");
  if (p.M == 0) return 0;
  dim3 grid((unsigned)ceil_div(p.M, BM), (unsigned)ceil_div(p.N, BN));
  if (d->dtype == BDE_F32)
    gemm_simt_kernel<float><<<grid, THREADS, 0, s>>>(p);
  else
    gemm_simt_kernel<__nv_bfloat16><<<grid, THREADS, 0, s>>>(p);
  return check_launch("gemm_simt_kernel");
}

}  // namespace bde
