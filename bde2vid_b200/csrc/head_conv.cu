// Head convolution of the UNet, straight from the voxeliser's planar grid:
//
//   out[n, y, x, :] = act( conv5x5( vox[n, :, y, x] ) + bias )        (model/BDE2VID/..._V5.py:116, submodules.py:85-114)
//
//   vox : float32 [N, Cin, H, W]  planar (Cin = num_bins <= 6; the layout events_to_voxel_torch produces)
//   w   : float32 [32, Cin, 5, 5] (the checkpoint tensor head.conv2d.weight, no repacking)
//   out : bf16    [N, H, W, 32]   NHWC, the layout every later layer reads
//
// With 5 input channels the implicit-GEMM engines spend their time gathering 16-byte im2col chunks (one per tap);
// here a CTA stages a 16 x 32 pixel tile plus halo once, as bf16 PAIRS (v[x], v[x+1]) so that every A-fragment
// register of mma.sync.m16n8k16 is a single aligned 32-bit shared-memory load:
//   k16 step = (ky, channel pair): columns 0-7 = kx of the even channel, 8-15 = kx of the odd one (kx >= 5 and the
//   missing 6th channel carry zero weights).  15 steps x 4 n-tiles of 8 output channels per 16 pixels.
// Operands are rounded to bf16 exactly as the packed-voxel + bde_gemm path did (fp32 accumulation).
#include "common.cuh"

namespace bde {
namespace {

constexpr int kHcTW = 32, kHcTH = 16;            // output tile
constexpr int kHcK = 5, kHcPad = 2;
constexpr int kHcCout = 32;
constexpr int kHcRows = kHcTH + kHcK - 1;        // 20 halo rows
constexpr int kHcPitch = 44;                      // pair columns per halo row (needs 32 + 4 + 7; 44 keeps rows 16-byte multiples)
constexpr int kHcPairs = 3;                       // channel pairs (Cin <= 6)
constexpr int kHcSteps = kHcK * kHcPairs;         // 15 k16 steps
constexpr int kHcWPitch = 12;                     // 32-bit words per (step, n) weight row: 8 used (16 bf16) + pad, conflict-free
constexpr int kHcThreads = 256;

struct HeadConvParams {
  const float* vox;
  const float* w;
  const float* bias;
  __nv_bfloat16* out;
  int n_img, cin, h, w_px;
  float lo, hi;   // activation clamp
};

__device__ __forceinline__ void mma16816_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

__global__ void __launch_bounds__(kHcThreads, 2) head_conv_kernel(const HeadConvParams p) {
  // pairs[c][row][col] = bf16x2 (v[col], v[col + 1]) of halo column col (image x = x0 - 2 + col)
  __shared__ uint32_t pairs[2 * kHcPairs][kHcRows][kHcPitch];
  // wk[step][n][word]: word j < 4 -> (w[kx = 2j], w[kx = 2j + 1]) of the even channel, 4 <= j < 8 -> odd channel
  __shared__ uint32_t wk[kHcSteps][kHcCout][kHcWPitch];
  __shared__ float bias_s[kHcCout];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int x0 = blockIdx.x * kHcTW, y0 = blockIdx.y * kHcTH;
  const int H = p.h, W = p.w_px;

  // ---- weights -> smem (bf16 pairs; zero for kx >= 5 and channels >= cin), ONCE per CTA: the CTA then walks the images
  // n = blockIdx.z, blockIdx.z + gridDim.z, ... of its tile position (staging them per image was a third of the kernel) ----
  {
    const int j = tid & 7, co = tid >> 3;      // 32 output channels x 8 words = 256 threads
    const int kx = 2 * (j & 3);
    for (int step = 0; step < kHcSteps; ++step) {
      const int ky = step / kHcPairs, c = 2 * (step - ky * kHcPairs) + (j >> 2);
      float w0 = 0.f, w1 = 0.f;
      if (c < p.cin) {
        const float* wr = p.w + ((size_t)(co * p.cin + c) * kHcK + ky) * kHcK;
        if (kx < kHcK) w0 = __ldg(wr + kx);
        if (kx + 1 < kHcK) w1 = __ldg(wr + kx + 1);
      }
      wk[step][co][j] = pack_bf16x2(w0, w1);
    }
  }
  if (tid < kHcCout) bias_s[tid] = p.bias != nullptr ? __ldg(p.bias + tid) : 0.f;
  // 8-byte loads of (v[x], v[x + 1]): x0 - 2 is even, so rows must start 8-byte aligned
  const bool vec_ok = (W % 2 == 0) && ((((uintptr_t)p.vox) & 7) == 0);

  for (int n = blockIdx.z; n < p.n_img; n += gridDim.z) {
  // ---- input tile + halo -> smem as pairs (zero outside the image / beyond cin): one warp per (channel, halo row), lane l
  // loads v[2l], v[2l + 1], takes v[2l + 2] from its neighbour and stores the pair columns 2l and 2l + 1 ----
  for (int r = warp; r < 2 * kHcPairs * kHcRows; r += kHcThreads / 32) {
    const int c = r / kHcRows, row = r - c * kHcRows;
    const int y = y0 - kHcPad + row, x = x0 - kHcPad + 2 * lane;
    float v0 = 0.f, v1 = 0.f;
    if (c < p.cin && y >= 0 && y < H && lane < kHcPitch / 2 + 1) {
      const float* src = p.vox + ((size_t)(n * p.cin + c) * H + y) * W;
      if (vec_ok && x >= 0 && x + 1 < W) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(src + x));
        v0 = v.x; v1 = v.y;
      } else {
        if (x >= 0 && x < W) v0 = __ldg(src + x);
        if (x + 1 >= 0 && x + 1 < W) v1 = __ldg(src + x + 1);
      }
    }
    const float vn = __shfl_down_sync(0xffffffffu, v0, 1);
    if (lane < kHcPitch / 2)
      *reinterpret_cast<uint2*>(&pairs[c][row][2 * lane]) = make_uint2(pack_bf16x2(v0, v1), pack_bf16x2(v1, vn));
  }
  __syncthreads();

  // ---- warp w: output rows 2w, 2w + 1 of the tile = 4 M tiles of 16 pixels ----
  float acc[4][4][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[m][nt][0] = acc[m][nt][1] = acc[m][nt][2] = acc[m][nt][3] = 0.f;
#pragma unroll 1
  for (int ky = 0; ky < kHcK; ++ky) {
#pragma unroll
    for (int cp = 0; cp < kHcPairs; ++cp) {
      const int step = ky * kHcPairs + cp;
      uint32_t b0[4], b1[4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        b0[nt] = wk[step][nt * 8 + g][t];
        b1[nt] = wk[step][nt * 8 + g][4 + t];
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int row = 2 * warp + (m >> 1) + ky;      // halo row of this M tile's output row
        const int col = (m & 1) * 16 + g + 2 * t;      // halo column of (pixel g, kx = 2t)
        const uint32_t a0 = pairs[2 * cp][row][col], a1 = pairs[2 * cp][row][col + 8];
        const uint32_t a2 = pairs[2 * cp + 1][row][col], a3 = pairs[2 * cp + 1][row][col + 8];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma16816_bf16(acc[m][nt], a0, a1, a2, a3, b0[nt], b1[nt]);
      }
    }
  }

  // ---- bias + activation + bf16 NHWC store ----
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int y = y0 + 2 * warp + (m >> 1);
    if (y >= H) continue;
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int x = x0 + (m & 1) * 16 + g + hr * 8;
      if (x >= W) continue;
      __nv_bfloat16* dst = p.out + ((size_t)(n * H + y) * W + x) * kHcCout;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int co = nt * 8 + 2 * t;
        const float v0 = fminf(fmaxf(acc[m][nt][hr * 2 + 0] + bias_s[co], p.lo), p.hi);
        const float v1 = fminf(fmaxf(acc[m][nt][hr * 2 + 1] + bias_s[co + 1], p.lo), p.hi);
        *reinterpret_cast<uint32_t*>(dst + co) = pack_bf16x2(v0, v1);
      }
    }
  }
  __syncthreads();   // every warp is done with this image's halo tile before the next one overwrites it
  }  // image loop
}

}  // namespace
}  // namespace bde

using namespace bde;

extern "C" int bde_head_conv(const float* vox, const float* w, const float* bias, void* out, int n_img, int cin, int h, int w_px,
                             int cout, int ksize, int act, void* stream) {
  BDE_REQUIRE(vox != nullptr && w != nullptr && out != nullptr, "bde_head_conv: null operand");
  BDE_REQUIRE(cout == kHcCout && ksize == kHcK, "bde_head_conv: supports 32 output channels and a 5x5 kernel (got %d, %d)", cout, ksize);
  BDE_REQUIRE(cin >= 1 && cin <= 2 * kHcPairs, "bde_head_conv: 1..6 input channels (got %d)", cin);
  BDE_REQUIRE(act == BDE_ACT_NONE || act == BDE_ACT_RELU || act == BDE_ACT_RELU6, "bde_head_conv: activation none / ReLU / ReLU6");
  BDE_REQUIRE(n_img >= 0 && n_img < 65536 && h > 0 && w_px > 0, "bde_head_conv: bad shape");
  if (n_img == 0) return 0;
  HeadConvParams p;
  p.vox = vox; p.w = w; p.bias = bias; p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.n_img = n_img; p.cin = cin; p.h = h; p.w_px = w_px;
  p.lo = act == BDE_ACT_NONE ? -INFINITY : 0.f;
  p.hi = act == BDE_ACT_RELU6 ? 6.f : INFINITY;
  // CTAs walk the images of their tile position (weights staged once per CTA): enough CTAs in z for ~8 waves of two per SM
  const int tiles = ceil_div(w_px, kHcTW) * ceil_div(h, kHcTH);
  int gz = ceil_div(device_sm_count() * 2 * 8, tiles);
  gz = gz < 1 ? 1 : (gz > n_img ? n_img : gz);
  dim3 grid((unsigned)ceil_div(w_px, kHcTW), (unsigned)ceil_div(h, kHcTH), (unsigned)gz);
  head_conv_kernel<<<grid, kHcThreads, 0, (cudaStream_t)stream>>>(p);
  return check_launch("head_conv_kernel");
}
