// Loader-side voxel post-processing on the device: LegacyNorm, RobustNorm and the hot-pixel mask.
//
//   LegacyNorm      utils_func/data_augmentation.py:311-330
//   RobustNorm      utils_func/utils.py:7-51 (same class at utils_func/data_augmentation.py:258-308)
//   hot-pixel mask  events_contrast_maximization/utils/event_utils.py:100-116 (get_hot_event_mask),
//                   data_loader/h5_dataset.py:163-172
//
// The reference applies these per window on the CPU DataLoader workers (h5_dataset.py:226 transform_voxel) before
// Croper.pad; here one CTA per window works in place on the sensor area of the padded grid that the voxeliser wrote
// (the grid of a window is L2-resident: 1.8 MB at 346x260), so the padding ring stays zero.
#include "common.cuh"

namespace bde {
namespace {

constexpr int kNormThreads = 1024;

__device__ __forceinline__ double block_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[0] = v;
  }
  __syncthreads();
  const double r = red[0];
  __syncthreads();
  return r;
}

// cell i of the unpadded [bins, H, W] tensor -> pointer into the padded grid
struct SensorView {
  float* g;
  int H, W, Hp, Wp, pt, pl;
  __device__ __forceinline__ float* at(int i) const {
    const int x = i % W, r = i / W, y = r % H, b = r / H;
    return g + ((size_t)b * Hp + (y + pt)) * Wp + (x + pl);
  }
};

// ---- LegacyNorm ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNormThreads) legacy_norm_kernel(float* grids, size_t stride, int bins, int H, int W, int pt, int pl,
                                                                   int Hp, int Wp, float* stats) {
  __shared__ double red[32];
  SensorView v{grids + (size_t)blockIdx.x * stride, H, W, Hp, Wp, pt, pl};
  const int n = bins * H * W;
  double s = 0.0, s2 = 0.0, cnt = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = *v.at(i);
    if (x != 0.0f) {
      s += (double)x;
      s2 += (double)x * (double)x;   // x ** 2 is exact in double for a float x
      cnt += 1.0;
    }
  }
  s = block_sum(s, red);
  s2 = block_sum(s2, red);
  cnt = block_sum(cnt, red);
  // the reference's scalars are float32 tensors: mean = x.sum() / n;  std = sqrt((x ** 2).sum() / n - mean ** 2)
  const float nf = (float)cnt;
  float mean = 0.f, sd = 0.f;
  if (cnt > 0.0) {
    mean = __fdiv_rn((float)s, nf);
    sd = sqrtf(__fsub_rn(__fdiv_rn((float)s2, nf), __fmul_rn(mean, mean)));
  }
  if (stats != nullptr && threadIdx.x == 0) {
    float* o = stats + (size_t)blockIdx.x * 4;
    o[0] = mean; o[1] = sd; o[2] = nf; o[3] = 0.f;
  }
  if (!(cnt > 0.0) || sd == 0.0f) return;   // `if stddev != 0` is also true for NaN: handled below like the reference
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float* p = v.at(i);
    const float x = *p;
    // mask * (x - mean) / stddev, left to right
    *p = __fdiv_rn(__fmul_rn(x != 0.0f ? 1.0f : 0.0f, __fsub_rn(x, mean)), sd);
  }
}

// ---- RobustNorm: exact k-th smallest by a 4-pass radix select on order-preserving keys ----------------------------------
__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// k is 1-based.  hist: 256 counters in shared memory; sel: two words of shared scratch.
__device__ float kth_smallest(const SensorView& v, int n, int k, unsigned* hist, unsigned* sel) {
  uint32_t prefix = 0, mask = 0;
  unsigned rank = (unsigned)k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t key = f2key(*v.at(i));
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned acc = 0;
      int d = 0;
      for (; d < 256; ++d) {
        if (acc + hist[d] >= rank) break;
        acc += hist[d];
      }
      sel[0] = (unsigned)d;
      sel[1] = rank - acc;
    }
    __syncthreads();
    prefix |= sel[0] << shift;
    mask |= 255u << shift;
    rank = sel[1];
    __syncthreads();
  }
  return key2f(prefix);
}

__global__ void __launch_bounds__(kNormThreads) robust_norm_kernel(float* grids, size_t stride, int bins, int H, int W, int pt, int pl,
                                                                   int Hp, int Wp, int k_low, int k_top, float* stats) {
  __shared__ unsigned hist[256];
  __shared__ unsigned sel[2];
  SensorView v{grids + (size_t)blockIdx.x * stride, H, W, Hp, Wp, pt, pl};
  const int n = bins * H * W;
  const float t_max = kth_smallest(v, n, k_top, hist, sel);
  const float t_min = kth_smallest(v, n, k_low, hist, sel);
  if (stats != nullptr && threadIdx.x == 0) {
    float* o = stats + (size_t)blockIdx.x * 4;
    o[0] = t_min; o[1] = t_max; o[2] = 0.f; o[3] = 0.f;
  }
  if (t_max == 0.0f && t_min == 0.0f) return;
  // normed = clamp(x, t_min, t_max); (normed - min(normed)) / (max(normed) + 1e-6): both order statistics are elements
  // of x, so min(normed) = t_min and max(normed) = t_max (for t_min <= t_max)
  const float den = __fadd_rn(t_max, 1e-6f);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float* p = v.at(i);
    const float c = fminf(fmaxf(*p, t_min), t_max);
    *p = __fdiv_rn(__fsub_rn(c, t_min), den);
  }
}

// ---- hot-pixel mask -----------------------------------------------------------------------------------------------
__global__ void hot_accumulate_kernel(const int16_t* xs, const int16_t* ys, const uint8_t* ps, int64_t n, int H, int W, float* img) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int x = xs[e], y = ys[e];
    if (x >= 0 && x < W && y >= 0 && y < H) atomicAdd(img + (size_t)y * W + x, ps[e] ? 1.0f : -1.0f);   // integer-valued: exact
  }
}

// one CTA: num_hot rounds of (argmax -> mask = 0, img = 0); every thread caches the best of its own strided cells and
// only the owner of the cleared cell rescans
__global__ void __launch_bounds__(kNormThreads) hot_select_kernel(float* img, float* mask, int npix, int num_hot) {
  __shared__ float bv[32];
  __shared__ int bi[32];
  __shared__ int win_idx;
  for (int i = threadIdx.x; i < npix; i += blockDim.x) mask[i] = 1.0f;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  auto rescan = [&]() {
    best = -INFINITY;
    best_i = 0x7fffffff;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {
      const float x = img[i];
      if (x > best) { best = x; best_i = i; }     // strided ascending: the first maximum wins ties
    }
  };
  rescan();
  __syncthreads();
  for (int it = 0; it < num_hot; ++it) {
    float v = best;
    int ix = best_i;
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_down_sync(0xffffffffu, v, o);
      const int oi = __shfl_down_sync(0xffffffffu, ix, o);
      if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { bv[warp] = v; bi[warp] = ix; }
    __syncthreads();
    if (warp == 0) {
      v = lane < (blockDim.x >> 5) ? bv[lane] : -INFINITY;
      ix = lane < (blockDim.x >> 5) ? bi[lane] : 0x7fffffff;
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_down_sync(0xffffffffu, v, o);
        const int oi = __shfl_down_sync(0xffffffffu, ix, o);
        if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
      }
      if (lane == 0) win_idx = ix;
    }
    __syncthreads();
    const int w = win_idx;
    if (w >= 0 && w < npix && (w % (int)blockDim.x) == (int)threadIdx.x) {
      mask[w] = 0.0f;
      img[w] = 0.0f;
      rescan();
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace bde

using namespace bde;

extern "C" int bde_voxel_normalize(float* grids, size_t window_stride, int T, int num_bins, int H, int W, int pad_top, int pad_left,
                                   int Hp, int Wp, int mode, float low_perc, float top_perc, float* stats, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  BDE_REQUIRE(grids != nullptr && T >= 0 && num_bins >= 1 && H > 0 && W > 0, "bde_voxel_normalize: bad arguments");
  BDE_REQUIRE(pad_top >= 0 && pad_left >= 0 && Hp >= H + pad_top && Wp >= W + pad_left, "bde_voxel_normalize: bad padding");
  BDE_REQUIRE(window_stride >= (size_t)num_bins * Hp * Wp, "bde_voxel_normalize: window stride smaller than one grid");
  BDE_REQUIRE((size_t)num_bins * H * W < ((size_t)1 << 31), "bde_voxel_normalize: grid too large");
  if (T == 0) return 0;
  if (mode == 1) {
    legacy_norm_kernel<<<T, kNormThreads, 0, s>>>(grids, window_stride, num_bins, H, W, pad_top, pad_left, Hp, Wp, stats);
    return check_launch("legacy_norm_kernel");
  }
  BDE_REQUIRE(mode == 2, "bde_voxel_normalize: mode must be 1 (LegacyNorm) or 2 (RobustNorm)");
  BDE_REQUIRE(low_perc >= 0.f && low_perc <= 100.f && top_perc >= 0.f && top_perc <= 100.f, "bde_voxel_normalize: percentiles");
  const long n = (long)num_bins * H * W;
  // k = 1 + round(.01 * float(q) * (numel - 1)); Python's round() is round-half-to-even = nearbyint in the default mode
  const int k_low = 1 + (int)nearbyint(.01 * (double)low_perc * (double)(n - 1));
  const int k_top = 1 + (int)nearbyint(.01 * (double)top_perc * (double)(n - 1));
  robust_norm_kernel<<<T, kNormThreads, 0, s>>>(grids, window_stride, num_bins, H, W, pad_top, pad_left, Hp, Wp, k_low, k_top, stats);
  return check_launch("robust_norm_kernel");
}

extern "C" int bde_hot_pixel_mask(const int16_t* xs, const int16_t* ys, const uint8_t* ps, int64_t n, int H, int W, int num_hot,
                                  float* mask, float* scratch, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  BDE_REQUIRE(mask != nullptr && scratch != nullptr && H > 0 && W > 0 && n >= 0 && num_hot >= 0, "bde_hot_pixel_mask: bad arguments");
  BDE_REQUIRE((size_t)H * W < ((size_t)1 << 30), "bde_hot_pixel_mask: sensor too large");
  cudaError_t e = cudaMemsetAsync(scratch, 0, (size_t)H * W * sizeof(float), s);
  BDE_REQUIRE(e == cudaSuccess, "bde_hot_pixel_mask: memset: %s", cudaGetErrorString(e));
  if (n > 0) {
    const unsigned blocks = (unsigned)(ceil_div((size_t)n, 256) < 2048 ? ceil_div((size_t)n, 256) : 2048);
    hot_accumulate_kernel<<<blocks, 256, 0, s>>>(xs, ys, ps, n, H, W, scratch);
    const int rc = check_launch("hot_accumulate_kernel");
    if (rc != 0) return rc;
  }
  hot_select_kernel<<<1, kNormThreads, 0, s>>>(scratch, mask, H * W, num_hot);
  return check_launch("hot_select_kernel");
}
