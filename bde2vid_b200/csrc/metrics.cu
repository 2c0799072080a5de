// Frame metrics on the device: per-frame MSE and SSIM of the centre-cropped prediction against the ground truth
// (eval_models_seq.py:242-258 -> evaluate/metrics.py:42-65: F.mse_loss and skimage.metrics.structural_similarity with
// its defaults: 7x7 uniform window, K1 = 0.01, K2 = 0.03, sample covariance, mean over the valid interior).
// The reference moves every frame to the host and runs numpy per frame; here one launch handles a whole sequence and
// only 2 doubles per frame leave the GPU (they feed the path's single collective, the metric all-reduce).
//
//   pred : float32 [n, Hp, Wp]  model output on the padded grid; the crop window starts at (y0, x0)  (Croper.crop)
//   gt   : float32 [n, H, W]
//   out  : float64 [n, 2]       (mse, ssim) per frame, accumulated with double atomics (must be zeroed by the caller)
#include "common.cuh"

namespace bde {
namespace {

constexpr int kMtTile = 32;            // output pixels per tile side
constexpr int kMtWin = 7;
constexpr int kMtHalo = kMtTile + kMtWin - 1;   // 38

__global__ void __launch_bounds__(256) frame_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int H, int W,
                                                            int Hp, int Wp, int y0, int x0, double c1, double c2, double* __restrict__ out) {
  __shared__ float sa[kMtHalo][kMtHalo + 1], sb[kMtHalo][kMtHalo + 1];
  __shared__ double red[2][8];
  const int n = blockIdx.z;
  const int ty0 = blockIdx.y * kMtTile, tx0 = blockIdx.x * kMtTile;   // tile origin in the cropped image
  const float* pa = pred + ((size_t)n * Hp + y0) * Wp + x0;
  const float* pb = gt + (size_t)n * H * W;
  for (int i = threadIdx.x; i < kMtHalo * kMtHalo; i += blockDim.x) {
    const int r = i / kMtHalo, c = i - r * kMtHalo;
    const int y = ty0 + r, x = tx0 + c;
    const bool in = y < H && x < W;
    sa[r][c] = in ? pa[(size_t)y * Wp + x] : 0.f;
    sb[r][c] = in ? pb[(size_t)y * W + x] : 0.f;
  }
  __syncthreads();
  double mse = 0.0, ssim = 0.0;
  const int vh = H - (kMtWin - 1), vw = W - (kMtWin - 1);    // valid SSIM positions (window fully inside)
  for (int i = threadIdx.x; i < kMtTile * kMtTile; i += blockDim.x) {
    const int r = i / kMtTile, c = i - r * kMtTile;
    const int y = ty0 + r, x = tx0 + c;
    if (y < H && x < W) {   // squared error: every pixel of the crop, owned by the tile whose origin region holds it
      const double d = (double)sa[r][c] - (double)sb[r][c];
      mse += d * d;
    }
    if (y < vh && x < vw) {
      double sx = 0.0, sy = 0.0, sxx = 0.0, syy = 0.0, sxy = 0.0;
#pragma unroll
      for (int dy = 0; dy < kMtWin; ++dy)
#pragma unroll
        for (int dx = 0; dx < kMtWin; ++dx) {
          const double a = sa[r + dy][c + dx], b = sb[r + dy][c + dx];
          sx += a; sy += b; sxx += a * a; syy += b * b; sxy += a * b;
        }
      constexpr double NP = kMtWin * kMtWin, cov = NP / (NP - 1.0);
      const double ux = sx / NP, uy = sy / NP;
      const double vx = cov * (sxx / NP - ux * ux), vy = cov * (syy / NP - uy * uy), vxy = cov * (sxy / NP - ux * uy);
      ssim += ((2.0 * ux * uy + c1) * (2.0 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2));
    }
  }
  // block reduction
  for (int o = 16; o > 0; o >>= 1) {
    mse += __shfl_xor_sync(0xffffffffu, mse, o);
    ssim += __shfl_xor_sync(0xffffffffu, ssim, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = mse; red[1][warp] = ssim; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0, s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { m += red[0][w]; s += red[1][w]; }
    atomicAdd(out + 2 * n, m / ((double)H * W));
    atomicAdd(out + 2 * n + 1, s / ((double)vh * vw));
  }
}

}  // namespace
}  // namespace bde

using namespace bde;

extern "C" int bde_frame_metrics(const float* pred, const float* gt, int n, int H, int W, int Hp, int Wp, int y0, int x0,
                                 double data_range, double* out, void* stream) {
  BDE_REQUIRE(pred != nullptr && gt != nullptr && out != nullptr, "bde_frame_metrics: null operand");
  BDE_REQUIRE(H >= kMtWin && W >= kMtWin, "bde_frame_metrics: image smaller than the 7x7 SSIM window");
  BDE_REQUIRE(y0 >= 0 && x0 >= 0 && y0 + H <= Hp && x0 + W <= Wp, "bde_frame_metrics: crop window outside the padded frame");
  BDE_REQUIRE(n >= 0 && n < 65536, "bde_frame_metrics: bad frame count");
  if (n == 0) return 0;
  const double c1 = (0.01 * data_range) * (0.01 * data_range), c2 = (0.03 * data_range) * (0.03 * data_range);
  dim3 grid((unsigned)ceil_div(W, kMtTile), (unsigned)ceil_div(H, kMtTile), (unsigned)n);
  frame_metrics_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, gt, H, W, Hp, Wp, y0, x0, c1, c2, out);
  return check_launch("frame_metrics_kernel");
}
